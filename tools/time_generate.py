import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver
for m, n, B in [(200, 100, 32768), (50, 20, 262144), (500, 250, 2048), (201, 99, 8192)]:
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); A, b, c = solver.generate(7, 0, B, m, n); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        del A, b, c
    print('generate %d x (%d,%d): %.3f ms, %.2f M instances/s, %.0f GB/s written' % (B, m, n, ms, B / ms / 1e3, B * (m * n + m + n) * 8 / ms / 1e6))
