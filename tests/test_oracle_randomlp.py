"""CPU: pins the oracle (generator, solver stand-in, labelling) against the committed golden vectors."""
import json
import os

import numpy as np
import pytest

from oracle import randomlp as o


@pytest.fixture(scope='module')
def kat(golden_dir):
    with open(os.path.join(golden_dir, 'randomlp_kat.json')) as f:
        return json.load(f)


def test_seed_schedule(kat):
    # seed 0 -> step 685, seed 3231 -> step 578
    assert o.seed_schedule(0, 4) == kat['seed_schedule']['0'] == [0, 685, 1370, 2055]
    assert o.seed_schedule(3231, 4) == kat['seed_schedule']['3231'] == [3231, 3809, 4387, 4965]
    assert o.seed_schedule(7, 0) == []


def test_generator_known_answers(kat):
    A, b, c = o.generate_instance(10, 5, 0)
    assert A[0, 0] == 1.764052345967664          # numpy legacy stream, any (m,n) at seed 0
    assert A.shape == (10, 5) and A.flags['C_CONTIGUOUS']
    for it in kat['instances']:
        A, b, c = o.generate_instance(it['m'], it['n'], it['seed'])
        assert A[0, 0] == it['A00']
        assert abs(b[0] - it['b0']) <= 1e-13 * max(1.0, abs(it['b0']))   # dgemv order may differ across BLAS builds
        assert c[0] == it['c0']
        assert (c >= 0).all()


def test_solver_known_answers(kat):
    for it in kat['instances']:
        if it['m'] * it['n'] > 5000:
            continue
        p = o.create_lp_problem(it['m'], it['n'], it['seed'], with_stats=True)
        assert p['stats']['sc'] == it['status']
        assert [int(i) for i in p['active']] == it['active']
        if it['status'] == 2:
            assert abs(p['stats']['objval'] - it['objval']) <= 1e-9 * abs(it['objval'])
            assert p['stats']['active'] == it['n']     # non-degenerate: exactly n active rows
        else:
            assert p['stats']['objval'] is None and not p['stats']['success']
            assert all(l == 0 for _, l in p['labels'])


def test_survey_known_answers():
    # values recorded in SURVEY.md section 8(c)
    p = o.create_lp_problem(10, 5, 0, with_stats=True)
    assert p['b'][0] == pytest.approx(-4.19474316251259, rel=1e-13)
    assert p['c'][0] == 0.4017809362082619
    assert p['stats']['objval'] == pytest.approx(-2.56554105335413, rel=1e-10)
    assert list(p['active']) == [0, 5, 6, 8, 9]
    assert o.create_lp_problem(10, 5, 1)['active'] == [] and o.create_lp_problem(10, 5, 2)['active'] == []
    p3 = o.create_lp_problem(10, 5, 3, with_stats=True)
    assert p3['stats']['objval'] == pytest.approx(2.368364743974648, rel=1e-10) and list(p3['active']) == [0, 3, 4, 6, 9]
    p = o.create_lp_problem(50, 20, 0, with_stats=True)
    assert p['stats']['objval'] == pytest.approx(-2.088414069103431, rel=1e-10)
    assert list(p['active'][:10]) == [3, 9, 10, 11, 21, 22, 25, 27, 28, 29]


def test_config1_golden(golden_dir):
    """BASELINE.json configs[0] at its stated size: 1 000 instances, seeds 3231 + 578 i."""
    g = np.load(os.path.join(golden_dir, 'randomlp_config1.npz'))
    seeds = o.seed_schedule(3231, 1000)
    assert list(g['seeds']) == seeds and len(g['status']) == 1000
    for i in range(0, 1000, 25):
        p = o.create_lp_problem(50, 20, seeds[i], with_stats=True)
        assert p['stats']['sc'] == g['status'][i]
        lab = np.unpackbits(g['labels_packed'][i])[:50]
        assert [l for _, l in p['labels']] == list(lab)
        if g['status'][i] == 2:
            assert abs(p['stats']['objval'] - g['obj'][i]) <= 1e-12 * abs(g['obj'][i])


@pytest.mark.parametrize('name,stride', [('randomlp_config2.npz', 125), ('randomlp_300x150.npz', 16),
                                         ('randomlp_400x100.npz', 16), ('randomlp_500x250.npz', 32)])
def test_large_fixtures_against_live_oracle(golden_dir, name, stride):
    """A sample of every committed solver fixture is re-derived by the live oracle (HiGHS + certified polish)."""
    g = np.load(os.path.join(golden_dir, name))
    m, n = int(g['m']), int(g['n'])
    xi = {int(k): j for j, k in enumerate(g['x_index'])}
    for i in range(0, len(g['seeds']), stride):
        A, b, c = o.generate_instance(m, n, int(g['seeds'][i]))
        r = o.solve_batch(A[None], b[None], c[None])
        assert r['status'][0] == g['status'][i]
        assert (r['labels'][0] == np.unpackbits(g['labels_packed'][i])[:m]).all()
        if g['status'][i] == 2:
            assert r['certified'][0] and r['n_active'][0] == n
            assert abs(r['obj'][0] - g['obj'][i]) <= 1e-11 * abs(g['obj'][i])
            if i in xi:
                assert np.abs(r['x'][0] - g['x'][xi[i]]).max() <= 1e-11 * np.abs(r['x'][0]).max()


def test_polished_vertex_is_certified_and_sharper_than_highs():
    """The polished x is the optimal vertex to fp64 rounding: active rows have |slack| ~ 1e-13 (HiGHS' raw x leaves up to
    5e-7), it is primal and dual feasible, and it never moves further than HiGHS' own tolerance allows."""
    A, b, c = zip(*[o.generate_instance(60, 30, s) for s in range(40)])
    A, b, c = np.array(A), np.array(b), np.array(c)
    r = o.solve_batch(A, b, c)
    raw = o.solve_batch(A, b, c, polish=False)
    ok = r['status'] == 2
    assert ok.sum() > 10 and r['certified'][ok].all() and not raw['certified'].any()
    assert (r['status'] == raw['status']).all()
    assert r['max_active'][ok].max() <= 1e-11
    assert (r['n_active'][ok] == 30).all()
    rel = np.abs(r['x'] - raw['x'])[ok].max(axis=1) / np.abs(r['x'])[ok].max(axis=1)
    assert rel.max() <= 1e-6
    # a degenerate vertex (three lines through one point) cannot be certified by an n-row active set that is ambiguous:
    # the oracle then keeps the solver's x and says so
    A1 = np.array([[1.0, 0], [0, 1.0], [1.0, 1.0], [-1.0, 0], [0, -1.0]]); b1 = np.array([1.0, 1, 2, 5, 5]); c1 = np.array([-1.0, -1])
    d = o.solve_batch(A1[None], b1[None], c1[None])
    assert d['status'][0] == 2 and list(d['labels'][0]) == [1, 1, 1, 0, 0]


def test_reduced_lp_oracle_leg():
    """row_mask: the oracle solves the kept rows only, labels / violations are over all rows at that optimum."""
    A, b, c = zip(*[o.generate_instance(40, 10, s) for s in range(12)])
    A, b, c = np.array(A), np.array(b), np.array(c)
    full = o.solve_batch(A, b, c)
    rng = np.random.RandomState(1)
    mask = np.maximum(full['labels'], (rng.rand(12, 40) < 0.3).astype(np.uint8))
    red = o.solve_batch(A, b, c, row_mask=mask)
    ok = full['status'] == 2
    assert (red['status'][ok] == 2).all() and (red['labels'][ok] == full['labels'][ok]).all()
    assert (red['violations'][ok] == 0).all()
    assert np.abs(red['obj'][ok] - full['obj'][ok]).max() <= 1e-10
    bad = mask.copy()
    bad[np.arange(12), full['labels'].argmax(axis=1)] = 0          # drop one active row
    r2 = o.solve_batch(A[ok], b[ok], c[ok], row_mask=bad[ok])
    assert ((r2['status'] != 2) | (r2['violations'] > 0)).all()


def test_parallel_batch_equals_serial():
    A, b, c = zip(*[o.generate_instance(30, 12, s) for s in range(24)])
    A, b, c = np.array(A), np.array(b), np.array(c)
    s = o.solve_batch(A, b, c)
    p = o.solve_batch_parallel(A, b, c, procs=2)
    for k in ('status', 'labels', 'x', 'obj', 'certified'):
        assert (s[k] == p[k]).all(), k


def test_three_way_crosscheck():
    """HiGHS dual simplex == HiGHS interior point + crossover == independent tableau simplex (tools/algo_model.py)."""
    from tools import algo_model
    for seed in range(40):
        A, b, c = o.generate_instance(30, 12, seed)
        ds = o.solve_batch(A[None], b[None], c[None])
        ip = o.solve_batch(A[None], b[None], c[None], method='highs-ipm')
        mine = algo_model.solve(A, b, c)
        assert ds['status'][0] == ip['status'][0] == mine['status']
        if mine['status'] == 2:
            assert (ds['labels'][0] == ip['labels'][0]).all()
            assert list(np.flatnonzero(ds['labels'][0])) == list(mine['basis'])
            assert abs(mine['obj'] - ds['obj'][0]) <= 1e-9 * abs(ds['obj'][0])


def test_dataset_and_linprog_surface():
    ds = o.RandomLPDataset(10, 5, num_lps=3, seed=0)
    assert len(ds) == 3
    item = ds[4]                                    # idx % len  (randomlp_dataset.py:49)
    assert set(item) == {'lp', 'labels'} and set(item['lp']) == {'A', 'b', 'c'}
    assert len(item['labels']) == 10 and item['labels'][0][0] == 0
    st = ds.get_lp_params()
    assert [s['id'] for s in st] == [0, 685, 1370] and st[0]['sc'] == 2
    with pytest.raises(ValueError):
        o.LinProg(np.eye(2), np.ones(2), np.ones(2), 'sideways')
    lp = o.LinProg(np.array([[1.0, 0], [0, 1.0], [-1, -1]]), np.array([1.0, 1.0, 1.0]), np.array([-1.0, -1.0]))
    assert lp.get_statuscode() == 1                 # loaded, not yet optimised
    lp.optimize()
    assert lp.get_statuscode() == 2 and list(lp.get_active_constraints()) == [0, 1]
    assert lp.model.objVal == pytest.approx(-2.0)
