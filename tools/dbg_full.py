import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from deep_dantzig_b200 import solver, _lib
B, m, n = 20000, 200, 100
r = solver.generate_solve_label(11, 0, B, m, n, keep_instances=True)
st = r['status'].cpu().numpy(); na = r['n_active'].cpu().numpy(); piv = r['pivots'].cpu().numpy(); ties = r['ties'].cpu().numpy()
ok = st == 2
bad = np.flatnonzero((ok & (na != n)) | (~ok & (na != 0)))
print('status hist', dict(zip(*np.unique(st, return_counts=True))), 'bad', len(bad))
ctx = _lib.context(0)
for i in bad[:10]:
    A = r['A'][i:i+1].contiguous(); b = r['b'][i:i+1].contiguous(); c = r['c'][i:i+1].contiguous()
    ctx.set_solve_plan(1); r1 = solver.solve_label(A, b, c); ctx.set_solve_plan(-1)
    x0 = r['x'][i].cpu().numpy(); x1 = r1['x'][0].cpu().numpy()
    print('LP', i, 'st', st[i], int(r1['status'][0]), 'nact', na[i], int(r1['n_active'][0]), 'piv', piv[i], r1['pivots'][0].cpu().numpy(), 'ties', ties[i],
          'relx', np.abs(x0 - x1).max() / max(np.abs(x1).max(), 1e-300), 'obj', float(r['obj'][i]), float(r1['obj'][0]))
