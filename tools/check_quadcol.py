"""Plan 7 (column block per warp) against plan 0 (row per thread) on the same Philox instances: statuses, labels, x / objective,
pivot counts; then device-timed throughput of both.  tools/check_quadcol.py [m n B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deep_dantzig_b200 import solver, _lib

shapes = [(200, 100, 32768), (50, 20, 2048), (150, 100, 1024), (228, 100, 512), (125, 100, 512), (100, 50, 1024), (300, 100 - 0, 0)]
if len(sys.argv) > 3:
    shapes = [(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))]
ctx = _lib.context(0)
for (m, n, B) in shapes:
    if B == 0 or m - n > 128:
        continue
    A, b, c = solver.generate(77, 0, B, m, n)
    res = {}
    for plan in (0, 7):
        ctx.set_solve_plan(plan)
        r = solver.solve_label(A, b, c)
        torch.cuda.synchronize()
        res[plan] = {k: v.cpu().numpy() for k, v in r.items()}
    ctx.set_solve_plan(-1)
    r0, r7 = res[0], res[7]
    ok = r0['status'] == 2
    st = (r0['status'] == r7['status']).mean()
    lb = (r0['labels'] == r7['labels']).all(axis=1).mean()
    dx = np.abs(r0['x'][ok] - r7['x'][ok]).max() / np.abs(r0['x'][ok]).max() if ok.any() else 0.0
    do = np.abs(r0['obj'][ok] - r7['obj'][ok]).max() / np.abs(r0['obj'][ok]).max() if ok.any() else 0.0
    pv = (r0['pivots'] == r7['pivots']).all(axis=1).mean()
    print('(%d,%d) B=%d: status equal %.4f, labels equal %.4f, pivots equal %.4f, max rel dx %.2e dobj %.2e, optimal %.3f, mean pivots %s vs %s'
          % (m, n, B, st, lb, pv, dx, do, ok.mean(), r0['pivots'].mean(0), r7['pivots'].mean(0)), flush=True)
m, n, B = (200, 100, 32768) if len(sys.argv) <= 3 else shapes[0]
A, b, c = solver.generate(42, 0, B, m, n)
out = solver._alloc_outputs(B, m, n, A.device)
for plan in (0, 7, 0, 7):
    ctx.set_solve_plan(plan)
    solver.solve_label(A, b, c, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        solver.solve_label(A, b, c, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('plan %d: %d LPs (%d,%d) %.3f ms = %.0f LP/s; flagged-for-fixup share n/a' % (plan, B, m, n, ms, B / ms * 1e3), flush=True)
ctx.set_solve_plan(-1)
