"""How well does the reference's classifier learn on streamed random LPs?  (exploration for the config-4 driver)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deep_dantzig_b200 import solver
from deep_dantzig_b200.ml.models.s2v import Model
from deep_dantzig_b200.ml import train as tr
m, n = 200, 100
A, b, c = solver.generate(999, 0, 2048, m, n)
res = solver.solve_label(A, b, c)
keep = res['status'] == 2
y = res['labels'].long()
val = [{'A': A, 'b': b, 'c': c, 'y': y}]
for lr, mom, w, opt_name in [(1e-6, 0.9, (0.25, 0.75), 'sgd'), (5e-6, 0.9, (0.25, 0.75), 'sgd'), (3e-3, 0.0, (0.25, 0.75), 'adam'), (1e-2, 0.0, (0.5, 0.5), 'adam')]:
    torch.manual_seed(0)
    model = Model('bipartite', 40, 3, on_cuda=True, verbose_init=False)
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=mom) if opt_name == 'sgd' else torch.optim.Adam(model.parameters(), lr=lr)
    crit = torch.nn.NLLLoss(weight=torch.tensor(w, device='cuda'), reduction='sum')
    for rnd in range(3):
        h = tr.train_on_device_stream(model, opt, m, n, steps=300, batch_per_rank=1024, key=100 + rnd, weight=w)
        perf = tr.performance(val, model, crit, 0.5)
        with torch.no_grad():
            model.forward_batch(A[keep], b[keep], c[keep])
        pr = model.probs[..., 1]
        yy = y[keep]
        pos = pr[yy == 1]
        q = [float(torch.quantile(pos, t)) for t in (0.0, 0.001, 0.01)]
        kept = [float((pr >= t).float().mean()) for t in q]
        print('%s lr=%g w=%s round %d: loss/node %.4f acc %.3f prec %.3f rec %.3f | optimal LPs: thresholds for recall 1/0.999/0.99 keep %.3f / %.3f / %.3f of the rows'
              % (opt_name, lr, w, rnd, h['loss'][-20:].mean(), perf['accuracy'], perf['precision'], perf['recall'], kept[0], kept[1], kept[2]), flush=True)
