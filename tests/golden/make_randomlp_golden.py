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
    # BASELINE.json config 1: (50,20), seed 3231, first 256 of the 1k instances
    ds_seeds = o.seed_schedule(3231, 256)
    status, packed, obj = [], [], []
    for s in ds_seeds:
        p = o.create_lp_problem(50, 20, s, with_stats=True)
        status.append(p['stats']['sc'])
        obj.append(p['stats']['objval'] if p['stats']['success'] else np.nan)
        packed.append(np.packbits(np.array([l for _, l in p['labels']], np.uint8)))
    np.savez_compressed(os.path.join(HERE, 'randomlp_config1.npz'), seeds=np.array(ds_seeds), status=np.array(status, np.int32),
                        labels_packed=np.stack(packed), obj=np.array(obj))
    print('wrote', len(kat['instances']), 'KAT instances and', len(ds_seeds), 'config-1 instances')


if __name__ == '__main__':
    main()
