"""Times plan 7 of several builds of the library (DDB200_LIBRARY) on the same batch: tools/time_ab.py lib1.so lib2.so ..."""
import os, subprocess, sys
code = '''
import torch, sys
sys.path.insert(0, '.')
from deep_dantzig_b200 import solver, _lib
ctx = _lib.context(0)
A, b, c = solver.generate(42, 0, 32768, 200, 100)
out = solver._alloc_outputs(32768, 200, 100, A.device)
ctx.set_solve_plan(7)
solver.solve_label(A, b, c, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4): solver.solve_label(A, b, c, out=out)
e1.record(); torch.cuda.synchronize()
print('%.3f ms = %.0f LP/s' % (e0.elapsed_time(e1) / 4, 32768 / (e0.elapsed_time(e1) / 4) * 1e3))
'''
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != 'default':
        env['DDB200_LIBRARY'] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    print(lib, r.stdout.strip(), r.stderr.strip()[-200:], flush=True)
