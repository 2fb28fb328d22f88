"""GPU: the MPS / PLNN ingestion path end to end -- MPS files with equality rows, '>' rows and variable
bounds are read, solved on the B200 solver (``LinProg.solve_mps`` writes the ``.info`` side files the reference obtained from
Gurobi), checked against HiGHS with the same bounds, turned into ``DatasetPLNN`` items (both graph layouts) and pushed through
the classifier and one reference-style training epoch."""
import json
import math
import os

import numpy as np
import pytest
import torch
from scipy.optimize import linprog

pytestmark = pytest.mark.gpu


def _make_tree(root, ndirs=3, per_dir=6, seed=0):
    """<root>/data/plnn/problem_k/lp_j.mps: bounded-feasible LPs with '<', '>' and '=' rows and mixed variable bounds."""
    from deep_dantzig_b200.data import mps
    rs = np.random.RandomState(seed)
    paths = []
    for d in range(ndirs):
        pdir = os.path.join(root, 'data', 'plnn', 'problem_%d' % d)
        os.makedirs(pdir)
        for j in range(per_dir + (d == 1)):                       # problem_1 is the largest directory
            n = int(rs.randint(4, 9)); m = int(rs.randint(n + 3, 2 * n + 6))
            A = rs.randn(m, n); A[rs.rand(m, n) < 0.25] = 0.0
            x0 = rs.uniform(-1.0, 1.0, n)
            ops = [['<', '>', '='][k] for k in rs.choice(3, m, p=[0.6, 0.25, 0.15])]
            slack = np.abs(rs.randn(m)) + 0.1
            b = A.dot(x0) + np.array([0.0 if op == '=' else (s if op == '<' else -s) for op, s in zip(ops, slack)])
            c = rs.randn(n)
            lb = np.where(rs.rand(n) < 0.7, x0 - rs.uniform(0.2, 2.0, n), -math.inf)
            ub = np.where(rs.rand(n) < 0.7, x0 + rs.uniform(0.2, 2.0, n), math.inf)
            for q in range(n):                                     # keep every variable bounded on at least one side: LP bounded
                if lb[q] == -math.inf and ub[q] == math.inf:
                    lb[q] = x0[q] - 1.0
                if (c[q] > 0 and lb[q] == -math.inf) or (c[q] < 0 and ub[q] == math.inf):
                    c[q] = -c[q]
            path = os.path.join(pdir, 'lp_%d.mps' % j)
            mps.write_mps(path, A, b, c, ops, lb=list(lb), ub=list(ub), obj='min' if j % 3 else 'max' if False else 'min')
            paths.append((path, A, b, c, ops, lb, ub))
    return paths


def test_mps_solve_matches_highs_and_dataset_items(cuda_device, tmp_path):
    from deep_dantzig_b200.data.plnn_dataset import DatasetPLNN
    from deep_dantzig_b200.data.gurobi_lp import LinProg
    from deep_dantzig_b200.ml.models.s2v import Model
    root = str(tmp_path)
    lps = _make_tree(root)
    assert DatasetPLNN.write_infos(dataset='plnn', root=root) == len(lps)
    nopt = 0
    for path, A, b, c, ops, lb, ub in lps:
        info = json.load(open(os.path.splitext(path)[0] + '.info'))
        ubi = [i for i, op in enumerate(ops) if op == '<']; gei = [i for i, op in enumerate(ops) if op == '>']
        eqi = [i for i, op in enumerate(ops) if op == '=']
        Aub = np.vstack([A[ubi], -A[gei]]) if (ubi or gei) else None
        bub = np.concatenate([b[ubi], -b[gei]]) if (ubi or gei) else None
        ref = linprog(c, A_ub=Aub, b_ub=bub, A_eq=A[eqi] if eqi else None, b_eq=b[eqi] if eqi else None,
                      bounds=[(None if l == -math.inf else l, None if u == math.inf else u) for l, u in zip(lb, ub)], method='highs-ds')
        assert (info['sc'] == 2) == (ref.status == 0), path
        if ref.status == 0:
            nopt += 1
            assert abs(info['objval'] - ref.fun) <= 1e-8 * max(1.0, abs(ref.fun)), (path, info['objval'], ref.fun)
            x = np.array([info['x_opt']['x%d' % j] for j in range(len(c))])
            slack = b - A.dot(x)
            want_active = ['c%d' % i for i in range(len(b)) if abs(slack[i]) <= 1e-7]
            assert sorted(info['active']) == sorted(want_active)
            assert all(abs(slack[i]) <= 1e-7 for i in eqi)                         # equalities hold
    assert nopt >= len(lps) * 0.8
    # dataset, both graph layouts, the three element types
    for graph in ('bipartite', 'complete'):
        for elem_type in ('lp', 'property', 'constraint'):
            ds = DatasetPLNN('plnn', graph, None, elem_type, seed=7, test=False, root=root)
            assert len(ds) >= 1 and abs(sum(ds.weight) - 1.0) < 1e-12
            it = ds[0]
            if graph == 'bipartite':
                m, n = it['dims']['m'], it['dims']['n']
                assert it['c_feats'].shape == (m, 3) and it['v_feats'].shape == (n, 1) and len(it['c_labels']) == m
                assert all(it['c_feats'][i, 0] == 1 and it['c_feats'][i, 2] == 0 for i in it['in_loss'])
                assert set(it['c_feats'][:, 0].tolist()) <= {0.0, 1.0} and len(it['e_feats']['i']) == len(it['e_feats']['coeffs'])
            else:
                assert it['lp']['A'].shape[0] == len(it['node_labels']) == len(it['node_features']) - 1
                assert all(it['node_features'][i] == 1 for i in it['in_loss'])
    # the classifier on PLNN items (general-flag path) and one reference-style accumulation step per item
    ds = DatasetPLNN('plnn', 'bipartite', None, 'lp', seed=7, test=False, root=root)
    torch.manual_seed(0)
    model = Model('bipartite', 8, 2, on_cuda=True, verbose_init=False)
    crit = torch.nn.NLLLoss(weight=torch.tensor(ds.weight, dtype=torch.float32).cuda(), reduction='sum')
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)
    opt.zero_grad()
    total = 0.0
    for k in range(len(ds)):
        it = ds[k]
        y = it['c_labels'][it['in_loss']].long().cuda()
        fx = model.forward(it)
        assert fx.shape == (len(it['in_loss']), 2) and torch.isfinite(fx).all()
        loss = crit(fx, y)
        loss.backward()
        total += float(loss)
    opt.step()
    assert math.isfinite(total) and all(torch.isfinite(q).all() for q in model.parameters())
    # has_matrix_inequalities / ineq_num bookkeeping
    d = LinProg.ineq_num(lps[0][0])
    assert d['num_constrs'] == len(lps[0][2]) and d['num_pos'] + d['num_inactive_ineq'] == d['num_ineq']
    # the reference's own experiment driver on the MPS tree (src/benchmark.py:46-95): DatasetPLNN + per-item training loop
    from deep_dantzig_b200 import benchmark
    for graph in ('bipartite', 'complete'):
        res, trained = benchmark.run_experiment_batch('plnn', graph, 'lp', None, 8, 2, epochs=2, batch_size=3, learning_rate=1e-3,
                                                      momentum=0.9, weight_decay=0.0, seed=7, cuda=True, tag='t', root=root)
        assert res['dataset'] == 'plnn' and set(res['out']['results']) == {'train', 'test'}
        assert len(res['out']['results']['train']) == 2 and res['out']['lps']
        for ep in res['out']['results']['train'] + res['out']['results']['test']:
            assert math.isfinite(ep['total_loss']) and 0.0 <= ep['accuracy'] <= 1.0 and 0.0 <= ep['recall'] <= 1.0
        assert res['out']['results']['train'][-1]['recall'] == 1.0          # the threshold is the train set's recall-1 point
        assert all(torch.isfinite(q).all() for q in trained.parameters())
