"""Device-timed solve throughput at several shapes / plans (resident Philox batches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver, _lib

cases = [(500, 250, 592, -1), (300, 150, 1184, -1), (400, 100, 2368, -1), (50, 20, 65536, -1), (200, 100, 4736, 1), (200, 100, 4736, 2), (100, 50, 16384, -1)]
ctx = _lib.context(0)
for m, n, B, plan in cases:
    ctx.set_solve_plan(plan)
    A, b, c = solver.generate(42, 0, B, m, n)
    out = solver._alloc_outputs(B, m, n, A.device)
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); solver.solve_label(A, b, c, out=out); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    pv = out['pivots'].float().mean(0).tolist()
    print('(%d,%d) B=%d plan %d (auto -> %d): %.3f ms, %.0f LP/s, optimal %.3f, pivots crash %.1f p1 %.1f p2 %.1f' % (
        m, n, B, plan, ctx.solve_plan(m, n), ms, B / ms * 1e3, (out['status'] == 2).float().mean().item(), pv[0], pv[1], pv[2]))
    del A, b, c, out
ctx.set_solve_plan(-1)
