import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1: os.environ['DDB_S2V_NO_DENSE'] = sys.argv[1]
import numpy as np, torch
from deep_dantzig_b200 import solver
from deep_dantzig_b200.ml.models.s2v import Model
from oracle import classifier as oc
for (m, n, p, T) in [(120, 128, 64, 2), (120, 128, 12, 2), (128, 120, 64, 2), (120, 128, 64, 1)]:
    torch.manual_seed(21)
    model = Model('bipartite', p, T, on_cuda=True, verbose_init=False)
    B = 700
    A, b, c = solver.generate(13, 0, B, m, n)
    sparse_ids = [3, 150, 151, B - 1]
    for k in sparse_ids:
        A[k, (k * 7) % m, (k * 3) % n] = 0.0
        A[k, 0, :n // 2] = 0.0
    with torch.no_grad():
        lp = model.forward_batch(A, b, c)
        lpt = model.forward_batch_torch(A, b, c)
    P = {k: v.detach().cpu() for k, v in model.named_parameters()}
    for k in sparse_ids + [0]:
        ref, _ = oc.forward('bipartite', P, A[k].cpu().numpy(), b[k].cpu().numpy(), c[k].cpu().numpy(), T)
        print((m, n, p, T), k, 'kernel-oracle %.2e  torch-oracle %.2e  kernel-torch %.2e  max|ref| %.1f' % (
            (lp[k].cpu() - ref).abs().max().item(), (lpt[k].cpu() - ref).abs().max().item(), (lp[k] - lpt[k]).abs().max().item(), ref.abs().max().item()))
