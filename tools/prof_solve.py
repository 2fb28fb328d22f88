"""Small driver for ncu captures: one warm-up + `reps` solve launches over a resident Philox batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver

m = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1184
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
if len(sys.argv) > 5:
    from deep_dantzig_b200 import _lib
    _lib.context(0).set_solve_plan(int(sys.argv[5]))
A, b, c = solver.generate(42, 0, B, m, n)
out = solver._alloc_outputs(B, m, n, A.device)
for _ in range(1 + reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); solver.solve_label(A, b, c, out=out); e1.record()
    torch.cuda.synchronize()
    print('solve %d LPs (%d,%d): %.3f ms, %.0f LP/s, optimal %.3f, mean pivots %.1f' % (
        B, m, n, e0.elapsed_time(e1), B / e0.elapsed_time(e1) * 1e3, (out['status'] == 2).float().mean().item(),
        out['pivots'][:, 3].float().mean().item()))
