"""Times the complete-variant forward (tcgen05 Gram kernel + rounds kernel) and loss+grad: tools/time_gram.py [m n p T B]"""
import sys, torch
sys.path.insert(0, '.')
from deep_dantzig_b200 import solver
from deep_dantzig_b200.ml.models.s2v import Model
m, n, p, T, B = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (200, 100, 40, 3, 8192))]
A, b, c = solver.generate(3, 0, B, m, n)
y = solver.solve_label(A, b, c)['labels']
model = Model('complete', p, T, on_cuda=True, verbose_init=False)
def timed(fn, reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
with torch.no_grad():
    f = timed(lambda: model.forward_batch(A, b, c))
g = timed(lambda: (model.zero_grad(), model.loss_and_grad_batch(A, b, c, y, [0.25, 0.75])))
print('complete (%d,%d) p=%d T=%d B=%d: forward %.3f ms = %.2f M inst/s; loss+grad %.3f ms = %.2f M inst/s' % (m, n, p, T, B, f, B / f / 1e3, g, B / g / 1e3))
