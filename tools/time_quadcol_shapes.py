"""Plan 0 against plan 7 over the shapes plan 7 covers (device-timed, resident Philox batches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver, _lib
ctx = _lib.context(0)
for (m, n) in [(200, 100), (150, 100), (125, 100), (228, 100), (180, 90), (160, 80), (150, 75), (144, 72), (120, 60), (100, 50), (300, 100 - 28), (50, 20)]:
    if m - n > 128:
        continue
    B = 16384
    A, b, c = solver.generate(42, 0, B, m, n)
    out = solver._alloc_outputs(B, m, n, A.device)
    r = {}
    for plan in (0, 7):
        ctx.set_solve_plan(plan)
        solver.solve_label(A, b, c, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            solver.solve_label(A, b, c, out=out)
        e1.record(); torch.cuda.synchronize()
        r[plan] = B / (e0.elapsed_time(e1) / 2) * 1e3
    ctx.set_solve_plan(-1)
    print('(%d,%d): plan 0 %.0f LP/s, plan 7 %.0f LP/s (%.2fx)' % (m, n, r[0], r[7], r[7] / r[0]), flush=True)
