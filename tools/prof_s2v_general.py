"""Timing / ncu driver for the general-adjacency loss+gradient path (instances with zero coefficients): tools/prof_s2v_general.py [m n p T B density]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver
from deep_dantzig_b200.ml.models.s2v import Model

m, n, p, T, B = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (200, 100, 40, 3, 1184))]
density = float(sys.argv[6]) if len(sys.argv) > 6 else 0.5
model = Model('bipartite', p, T, on_cuda=True, verbose_init=False)
A, b, c = solver.generate(49, 0, B, m, n, density=density)
y = solver.solve_label(A, b, c)['labels']
for it in range(4):
    model.zero_grad()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); l = model.loss_and_grad_batch(A, b, c, y, [0.25, 0.75]); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('general-adjacency loss+grad %d x (%d,%d) density %.2f p=%d T=%d: %.3f ms, %.0f inst/s; loss %.6g' % (B, m, n, density, p, T, ms, B / ms * 1e3, float(l)))
