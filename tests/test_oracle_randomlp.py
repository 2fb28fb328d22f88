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
    g = np.load(os.path.join(golden_dir, 'randomlp_config1.npz'))
    seeds = o.seed_schedule(3231, 256)
    assert list(g['seeds']) == seeds
    for i in range(0, 256, 8):
        p = o.create_lp_problem(50, 20, seeds[i], with_stats=True)
        assert p['stats']['sc'] == g['status'][i]
        lab = np.unpackbits(g['labels_packed'][i])[:50]
        assert [l for _, l in p['labels']] == list(lab)


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
