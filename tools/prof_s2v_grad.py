"""Timing driver for the classifier loss+gradient kernel against autograd through the batched torch restatement."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver
from deep_dantzig_b200.ml.models.s2v import Model

m, n, p, T, B = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (200, 100, 40, 3, 8192))]
model = Model('bipartite', p, T, on_cuda=True, verbose_init=False)
A, b, c = solver.generate(42, 0, B, m, n)
y = solver.solve_label(A, b, c)['labels']
w = [0.3, 0.7]
crit = torch.nn.NLLLoss(weight=torch.tensor(w, device='cuda'), reduction='sum')
for it in range(4):
    model.zero_grad()
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e0.record(); l = model.loss_and_grad_batch(A, b, c, y, w); e1.record()
    model.zero_grad()
    l2 = crit(model.forward_batch_torch(A, b, c).reshape(-1, 2), y.long().reshape(-1)); l2.backward(); e2.record()
    torch.cuda.synchronize()
    ms, ms2 = e0.elapsed_time(e1), e1.elapsed_time(e2)
    print('loss+grad %d x (%d,%d) p=%d T=%d: kernel %.3f ms (%.0f inst/s, %.1f GB/s of fp64 A), torch autograd %.3f ms; loss %.6g vs %.6g' % (
        B, m, n, p, T, ms, B / ms * 1e3, B * (m * n + m + n) * 8 / ms / 1e6, ms2, float(l), float(l2)))
