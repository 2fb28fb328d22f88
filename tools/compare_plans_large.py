"""Plan 7 against plan 0 on a large Philox batch: how many instances have identical status / labels / pivot counts / x bits.
   python tools/compare_plans_large.py [B] [key]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
key = int(sys.argv[2]) if len(sys.argv) > 2 else 2026
ctx = _lib.context(0)
tot = {'instances': 0, 'status_equal': 0, 'labels_equal': 0, 'pivots_equal': 0, 'x_bits_equal': 0, 'optimal': 0}
mx = 0.0
for lo in range(0, B, 32768):
    nb = min(32768, B - lo)
    A, b, c = solver.generate(key, lo, nb, 200, 100)
    ctx.set_solve_plan(7); r7 = {k: v.clone() for k, v in solver.solve_label(A, b, c).items()}
    ctx.set_solve_plan(0); r0 = solver.solve_label(A, b, c)
    ctx.set_solve_plan(-1)
    ok = r0['status'] == 2
    tot['instances'] += nb
    tot['optimal'] += int(ok.sum())
    tot['status_equal'] += int((r7['status'] == r0['status']).sum())
    tot['labels_equal'] += int((r7['labels'] == r0['labels']).all(dim=1).sum())
    tot['pivots_equal'] += int((r7['pivots'] == r0['pivots']).all(dim=1).sum())
    tot['x_bits_equal'] += int((r7['x'] == r0['x']).all(dim=1).sum())
    if ok.any():
        mx = max(mx, float(((r7['x'][ok] - r0['x'][ok]).abs().max() / r0['x'][ok].abs().max()).item()))
tot['max_rel_x_diff'] = mx
tot['shape'] = [200, 100]; tot['key'] = key
print(json.dumps(tot))
