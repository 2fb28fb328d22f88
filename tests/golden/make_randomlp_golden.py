"""Regenerates tests/golden/randomlp_kat.json and randomlp_config1.npz from the oracle (numpy legacy stream + HiGHS).

The reference holds no golden vectors for this path and its solver (Gurobi) is absent, so these
are derived: the generator values are numpy's frozen legacy stream (what the reference itself calls), the solver
values come from scipy/HiGHS dual simplex and were cross-checked against HiGHS-ipm when this file was made.
Run from the repo root:  python tests/golden/make_randomlp_golden.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
from oracle import randomlp as o  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    kat = {'seed_schedule': {'0': o.seed_schedule(0, 4), '3231': o.seed_schedule(3231, 4)}, 'instances': []}
    cases = [(10, 5, s) for s in range(8)] + [(50, 20, s) for s in range(6)] + [(200, 100, s) for s in range(4)] + \
            [(30, 20, 0), (25, 20, 1), (40, 10, 2), (64, 32, 3), (33, 17, 4)]
    for m, n, seed in cases:
        p = o.create_lp_problem(m, n, seed, with_stats=True)
        q = o.create_lp_problem(m, n, seed, with_stats=True, method='highs-ipm')
        assert p['stats']['sc'] == q['stats']['sc'], (m, n, seed)
        assert list(p['active']) == list(q['active']), (m, n, seed)
        kat['instances'].append({
            'm': m, 'n': n, 'seed': seed, 'A00': float(p['A'][0, 0]), 'b0': float(p['b'][0]), 'c0': float(p['c'][0]),
            'status': int(p['stats']['sc']), 'objval': p['stats']['objval'],
            'active': [int(i) for i in p['active']]})
    with open(os.path.join(HERE, 'randomlp_kat.json'), 'w') as f:
        json.dump(kat, f, indent=1)
    # BASELINE.json config 1 at its stated size: (50,20), seed 3231, all 1 000 instances (seeds 3231 + 578 i)
    write_fixture('randomlp_config1.npz', 50, 20, o.seed_schedule(3231, 1000), x_rows=1000)
    # BASELINE.json config 2 parity subset: (200,100), numpy seeds 0 + 685 i
    write_fixture('randomlp_config2.npz', 200, 100, o.seed_schedule(0, 2000), x_rows=256)
    # shapes beyond one SM's registers / shared memory (global-memory and cluster kernels)
    write_fixture('randomlp_500x250.npz', 500, 250, o.seed_schedule(0, 64), x_rows=8)
    write_fixture('randomlp_300x150.npz', 300, 150, o.seed_schedule(0, 64), x_rows=16)
    write_fixture('randomlp_400x100.npz', 400, 100, o.seed_schedule(0, 64), x_rows=16)
    print('wrote', len(kat['instances']), 'KAT instances')


def _fixture_job(args):
    os.environ['OMP_NUM_THREADS'] = '1'
    m, n, seeds = args
    A = np.empty((len(seeds), m, n)); b = np.empty((len(seeds), m)); c = np.empty((len(seeds), n))
    for i, s in enumerate(seeds):
        A[i], b[i], c[i] = o.generate_instance(m, n, int(s))
    return o.solve_batch(A, b, c)


def write_fixture(name, m, n, seeds, x_rows):
    """status / bit-packed labels / objective of every instance, x of the first `x_rows` optimal ones -- from the
    POLISHED oracle (HiGHS dual simplex, then the certified extended-precision vertex: oracle.randomlp.polish_vertex).
    oracle_tie marks instances where HiGHS' raw x would have been labelled differently (its 1e-7 tolerance against the
    reference's absolute 1e-7 threshold)."""
    import multiprocessing as mp
    seeds = np.asarray(seeds, np.int64)
    procs = os.cpu_count() or 1
    parts = [p for p in np.array_split(np.arange(len(seeds)), procs * 4) if len(p)]
    with mp.get_context('fork').Pool(procs) as pool:
        outs = pool.map(_fixture_job, [(m, n, seeds[p]) for p in parts])
    r = {k: np.concatenate([q[k] for q in outs]) for k in outs[0]}
    ok = r['status'] == 2
    obj = np.where(ok, r['obj'], np.nan)
    xi = np.flatnonzero(ok)[:x_rows]
    np.savez_compressed(os.path.join(HERE, name), m=m, n=n, seeds=seeds, status=r['status'].astype(np.int32),
                        labels_packed=np.packbits(r['labels'], axis=1), obj=obj, x_index=xi.astype(np.int32), x=r['x'][xi],
                        certified=r['certified'], oracle_tie=r['oracle_tie'], n_active=r['n_active'])
    print('%s: %d instances, %d optimal, %d certified, %d oracle ties, n_active != n on %d' %
          (name, len(seeds), ok.sum(), r['certified'].sum(), r['oracle_tie'].sum(), (r['n_active'][ok] != n).sum()))


if __name__ == '__main__':
    main()
