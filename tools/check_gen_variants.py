"""In-solver generator (fused mode 1) on the wide row-per-thread variants: same bits as solving the materialised instances.
Usage: check_gen_variants.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver, _lib

ctx = _lib.context(0)
bad = 0
for (m, n, B, dens) in [(400, 100, 300, 1.0), (250, 100, 400, 1.0), (300, 150, 300, 1.0), (400, 150, 200, 0.1), (484, 100, 200, 0.5)]:
    A, b, c = solver.generate(61, 7, B, m, n, density=dens)
    want = solver.solve_label(A, b, c)
    ctx.set_fused_mode(1)
    try:
        got = solver.generate_solve_label(61, 7, B, m, n, density=dens)
        keep = solver.generate_solve_label(61, 7, B, m, n, density=dens, keep_instances=True)
    finally:
        ctx.set_fused_mode(0)
    torch.cuda.synchronize()
    okA = bool((keep['A'] == A).all() and (keep['b'] == b).all() and (keep['c'] == c).all())
    same = {k: bool((got[k] == want[k]).all() and (keep[k] == want[k]).all()) for k in ('status', 'labels', 'pivots', 'n_active', 'ties', 'x')}
    print((m, n, dens), 'instances equal', okA, same, flush=True)
    bad += (not okA) + sum(not v for v in same.values())
print('MISMATCHES', bad)
