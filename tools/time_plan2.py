"""Plan 2 (global-memory tableau) throughput at (500,250) / (300,150) for the L2 budget given by DDB_PLAN2_L2_MB."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver, _lib
for m, n, B in [(500, 250, 592), (300, 150, 1184)]:
    A, b, c = solver.generate(42, 0, B, m, n)
    out = solver._alloc_outputs(B, m, n, A.device)
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); solver.solve_label(A, b, c, out=out); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('L2_MB=%s (%d,%d) B=%d: %.1f ms, %.0f LP/s' % (os.environ.get('DDB_PLAN2_L2_MB', 'default'), m, n, B, ms, B / ms * 1e3))
