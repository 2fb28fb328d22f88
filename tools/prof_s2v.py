"""Timing driver for the classifier forward kernels (CUDA events, resident Philox batch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver
from deep_dantzig_b200.ml.models.s2v import Model

graph = sys.argv[1] if len(sys.argv) > 1 else 'bipartite'
m, n, p, T, B = [int(v) for v in (sys.argv[2:7] if len(sys.argv) > 6 else (200, 100, 40, 3, 8192))]
model = Model(graph, p, T, on_cuda=True, verbose_init=False)
A, b, c = solver.generate(42, 0, B, m, n)
with torch.no_grad():
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lp = model.forward_batch(A, b, c); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print('%s forward %d x (%d,%d) p=%d T=%d: %.3f ms, %.0f inst/s, %.1f GB/s of fp64 A' % (
            graph, B, m, n, p, T, ms, B / ms * 1e3, B * (m * n + m + n) * 8 / ms / 1e6))
