"""Small end-to-end workload that launches every kernel family once on tiny batches (a quick does-everything-run check; compute-sanitizer is closed on this GPU pool, so it was run plainly)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver, _lib, reduced
from deep_dantzig_b200.ml.models.s2v import Model
from deep_dantzig_b200.ml import train as tr
ctx = _lib.context(0)
for (m, n, B) in [(50, 20, 96), (200, 100, 40)]:
    A, b, c = solver.generate(1, 0, B, m, n)
    for plan in (0, 1, 5):
        ctx.set_solve_plan(plan)
        r = solver.solve_label(A, b, c)
    ctx.set_solve_plan(-1)
    for graph in ('bipartite', 'complete'):
        model = Model(graph, 12, 2, on_cuda=True, verbose_init=False)
        with torch.no_grad():
            model.forward_batch(A, b, c)
    model = Model('bipartite', 12, 2, on_cuda=True, verbose_init=False)
    model.zero_grad()
    model.loss_and_grad_batch(A, b, c, r['labels'], [0.3, 0.7])
    As = A.clone(); As[0, 0, :3] = 0.0
    with torch.no_grad():
        model.forward_batch(As, b, c)
    thr = tr.recall_one_threshold([{'A': A, 'b': b, 'c': c, 'y': r['labels'].long()}], model)
    reduced.timing_forward_pass(model, A, b, c, thr)
A, b, c = solver.generate(2, 0, 8, 300, 150)
solver.solve_label(A, b, c)
solver.generate_solve_label(3, 0, 64, 50, 20)
torch.cuda.synchronize()
print('all kernel families ran')
