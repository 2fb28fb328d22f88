"""CPU: the MPS reader / writer, the mps2numpy mirror (reference src/data/mps2numpy.py) and the classifier's general-flag
forward for MPS / PLNN items against outputs of the unmodified reference model (tests/golden/s2v_plnn_items.npz)."""
import math
import os

import numpy as np
import pytest
import torch

from deep_dantzig_b200.data import mps
from deep_dantzig_b200.data import mps2numpy as m2n

FIXED = """NAME          TESTLP
ROWS
 N  COST
 L  LIM1
 G  LIM2
 E  MYEQN
COLUMNS
    X         COST         1.0   LIM1         1.0
    X         LIM2         1.0
    MARKER    'MARKER'     'INTORG'
    Y         COST         2.0   LIM1         1.0
    Y         MYEQN       -1.0
    MARKER    'MARKER'     'INTEND'
    Z         COST        -1.0   MYEQN        1.0
RHS
    RHS       COST       -10.0
    RHS       LIM1         4.0   LIM2         1.0
    RHS       MYEQN        7.0
BOUNDS
 UP BND       X            4.0
 LO BND       Y           -1.0
 UP BND       Y            1.0
 MI BND       Z
ENDATA
"""


def test_fixed_format_sections(tmp_path):
    p = tmp_path / 'fixed.mps'
    p.write_text(FIXED)
    model = mps.read_mps(str(p))
    assert model.ModelName == 'TESTLP' and model.ModelSense == 1 and model.ObjCon == 10.0
    assert [v.VarName for v in model.getVars()] == ['X', 'Y', 'Z']
    assert [(c.ConstrName, c.Sense, c.RHS) for c in model.getConstrs()] == [('LIM1', '<', 4.0), ('LIM2', '>', 1.0), ('MYEQN', '=', 7.0)]
    assert model.Obj == [1.0, 2.0, -1.0]
    x, y, z = model.getVars()
    assert (x.LB, x.UB) == (0.0, 4.0) and (y.LB, y.UB) == (-1.0, 1.0) and (z.LB, z.UB) == (-math.inf, math.inf)
    row = model.getRow(model.getConstrs()[0])
    assert row.size() == 2 and row.getVar(1).VarName == 'Y' and row.getCoeff(1) == 1.0
    item = m2n.model2numpy(model, standardize=True)
    #   matrix rows (the '>' row flipped), then x_lb, x_ub, y_lb, y_ub (z is free)
    want_A = np.array([[1, 1, 0], [-1, 0, 0], [0, -1, 1], [-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0]], dtype=float)
    want_b = np.array([4, -1, 7, 0, 4, 1, 1], dtype=float)
    assert (item['A'] == want_A).all() and (item['b'] == want_b).all() and list(item['c']) == [1.0, 2.0, -1.0]
    assert item['obj'] == 'min' and item['in_loss'] == [0, 1]
    assert item['cnames'] == {'LIM1': 0, 'LIM2': 1, 'MYEQN': 2, 'X_lb': 3, 'X_ub': 4, 'Y_lb': 5, 'Y_ub': 6}
    assert item['csenses']['MYEQN'] == '=' and item['csenses']['LIM2'] == '<'
    assert item['bounds']['Y']['lb'] == {'val': 1.0, 'sense': '<', 'name': 'Y_lb'} and item['bounds']['Z'] == {'lb': None, 'ub': None}
    raw = m2n.model2numpy(model, standardize=False)
    assert raw['csenses']['LIM2'] == '>' and (raw['A'][1] == [1, 0, 0]).all() and raw['b'][1] == 1.0
    A, b, c, ops, obj = m2n.mps2numpy(str(p))
    assert ops == ['<', '<', '=', '<', '<', '<', '<'] and obj == 'min' and A.shape == (7, 3)


def test_free_format_round_trip_and_objsense(tmp_path):
    rs = np.random.RandomState(0)
    m, n = 7, 4
    A = rs.randn(m, n); A[rs.rand(m, n) < 0.3] = 0.0
    b = rs.randn(m); c = rs.randn(n)
    ops = ['<', '>', '=', '<', '<', '>', '=']
    lb = [0.0, -math.inf, -2.0, 1.5]
    ub = [math.inf, math.inf, 3.0, 1.5]
    p = str(tmp_path / 'rt.mps')
    mps.write_mps(p, A, b, c, ops, lb=lb, ub=ub, obj='max')
    model = mps.read_mps(p)
    assert model.ModelSense == -1
    for j, v in enumerate(model.getVars()):
        assert v.LB == lb[j] and v.UB == ub[j] and v.Obj == c[j]
    for i, con in enumerate(model.getConstrs()):
        assert con.Sense == ops[i] and con.RHS == b[i]
        row = np.zeros(n)
        for v, coeff in con.terms:
            row[v.index] = coeff
        assert (row == A[i]).all()
    item = m2n.model2numpy(model)
    assert item['obj'] == 'min' and (item['c'] == -c).all()          # a 'max' model is standardised to min -c
    assert item['A'].shape[0] == m + 1 + 0 + 2 + 2                   # bounds: x0 lb; x2 lb, ub; x3 lb, ub (x1 free)


def test_negative_upper_bound_and_ranges(tmp_path):
    p = tmp_path / 'neg.mps'
    p.write_text('NAME N\nROWS\n N obj\n L r1\nCOLUMNS\n x obj 1 r1 1\nRHS\n rhs r1 5\nBOUNDS\n UP bnd x -2\nENDATA\n')
    v = mps.read_mps(str(p)).getVars()[0]
    assert v.UB == -2.0 and v.LB == -math.inf
    q = tmp_path / 'rng.mps'
    q.write_text('NAME N\nROWS\n N obj\n L r1\nCOLUMNS\n x obj 1 r1 1\nRHS\n rhs r1 5\nRANGES\n rng r1 2\nENDATA\n')
    with pytest.raises(NotImplementedError):
        mps.read_mps(str(q))


def test_flagged_items_match_the_reference_model(golden_dir):
    """Bipartite items with equality / bound rows and complete items with 0/1 node features through Model.forward (torch
    restatement with the item's flags) against the unmodified reference model's outputs."""
    from deep_dantzig_b200.ml.models.s2v import Model
    g = np.load(os.path.join(golden_dir, 's2v_plnn_items.npz'))
    for ci in range(int(g['ncases'])):
        pre = 'bip%d_' % ci
        m, n, p, T = [int(v) for v in g[pre + 'dims']]
        A = g[pre + 'A']
        idx = [[i, j] for i in range(m) for j in range(n) if A[i, j] != 0]
        item = {'c_feats': torch.from_numpy(g[pre + 'c_feats'].copy()), 'v_feats': torch.from_numpy(g[pre + 'v_feats']),
                'e_feats': {'i': idx, 'coeffs': [float(A[i, j]) for i, j in idx]}, 'in_loss': [int(q) for q in g[pre + 'in_loss']],
                'dims': {'m': m, 'n': n}}
        model = Model('bipartite', p, T, on_cuda=False, verbose_init=False)
        model.force_torch = True
        model.load_state_dict({k[len(pre) + 6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre + 'param_')})
        got = model.forward(item).detach().numpy()
        assert got.shape == g[pre + 'logp'].shape and np.abs(got - g[pre + 'logp']).max() <= 2e-5
        pre = 'cmp%d_' % ci
        item = {'A': torch.from_numpy(g[pre + 'A']).unsqueeze(0), 'b': torch.from_numpy(g[pre + 'b']).unsqueeze(0),
                'c': torch.from_numpy(g[pre + 'c']).unsqueeze(0), 'node_features': torch.from_numpy(g[pre + 'node_features']).unsqueeze(0),
                'in_loss': [int(q) for q in g[pre + 'in_loss']]}
        model = Model('complete', p, T, on_cuda=False, verbose_init=False)
        model.force_torch = True
        model.load_state_dict({k[len(pre) + 6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre + 'param_')})
        got = model.forward(item).detach().numpy()
        assert got.shape == g[pre + 'logp'].shape and np.abs(got - g[pre + 'logp']).max() <= 2e-5
