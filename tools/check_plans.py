"""Compare plan 0 (register-resident kernels) with plan 1/2 (generic kernel) on Philox instances: status, labels,
x, objective, pivot counts.  Usage: check_plans.py m n B [key]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deep_dantzig_b200 import solver, _lib

m, n, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
key = int(sys.argv[4]) if len(sys.argv) > 4 else 7
ctx = _lib.context(0)
A, b, c = solver.generate(key, 0, B, m, n)
res = {}
for plan in (0, 1):
    try:
        ctx.set_solve_plan(plan)
        r = solver.solve_label(A, b, c)
        torch.cuda.synchronize()
        res[plan] = {k: v.cpu().numpy() for k, v in r.items()}
    except Exception as e:
        print('plan', plan, 'failed:', e)
    finally:
        ctx.set_solve_plan(-1)
r0, r1 = res[0], res[1]
st_eq = (r0['status'] == r1['status'])
print('(%d,%d) B=%d  status equal %d/%d; status hist p0 %s p1 %s' % (
    m, n, B, st_eq.sum(), B, dict(zip(*np.unique(r0['status'], return_counts=True))),
    dict(zip(*np.unique(r1['status'], return_counts=True)))))
ok = (r0['status'] == 2) & (r1['status'] == 2)
lab_eq = (r0['labels'] == r1['labels']).all(axis=1)
print('labels equal on %d/%d; on both-optimal %d/%d' % (lab_eq.sum(), B, lab_eq[ok].sum(), ok.sum()))
if ok.any():
    relx = np.abs(r0['x'][ok] - r1['x'][ok]).max(axis=1) / np.abs(r1['x'][ok]).max(axis=1)
    relo = np.abs(r0['obj'][ok] - r1['obj'][ok]) / np.abs(r1['obj'][ok])
    print('max rel x diff %.3e, max rel obj diff %.3e' % (relx.max(), relo.max()))
print('mean pivots p0 %s  p1 %s' % (r0['pivots'].mean(axis=0), r1['pivots'].mean(axis=0)))
print('ties p0 %d p1 %d; n_active!=n on optimal: p0 %d' % (r0['ties'].sum(), r1['ties'].sum(),
                                                       (r0['n_active'][r0['status'] == 2] != n).sum()))
bad = np.flatnonzero(~st_eq | ~lab_eq)
for i in bad[:5]:
    print(' LP', i, 'status', r0['status'][i], r1['status'][i], 'pivots', r0['pivots'][i], r1['pivots'][i],
          'nact', r0['n_active'][i], r1['n_active'][i])
