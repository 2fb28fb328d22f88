"""Small runs of the column-block kernel (plain, row mask, in-solver generator) on the same instances: python tools/smoke_quadcol.py
(written for compute-sanitizer, which is closed on this pool)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver, _lib
ctx = _lib.context(0)
for (m, n, B, dens) in [(200, 100, 24, 1.0), (150, 100, 12, 0.3), (228, 100, 8, 1.0)]:
    A, b, c = solver.generate(5, 0, B, m, n, density=dens)
    ctx.set_solve_plan(7)
    r = solver.solve_label(A, b, c)
    mask = (torch.rand(B, m, device='cuda') < 0.85).to(torch.uint8)
    q = solver.solve_label(A, b, c, row_mask=mask)
    ctx.set_solve_plan(-1)
    ctx.set_fused_mode(1)
    g = solver.generate_solve_label(5, 0, B, m, n, density=dens)
    ctx.set_fused_mode(0)
    torch.cuda.synchronize()
    print((m, n), 'optimal', int((r['status'] == 2).sum()), int((q['status'] == 2).sum()), int((g['status'] == 2).sum()),
          'same as fused:', bool((r['status'] == g['status']).all() and (r['labels'] == g['labels']).all()), flush=True)
