"""Per-cell timing of the configs[2] sweep (m/n x density at n = 100): one fused generate -> solve -> label call per cell.
Usage: time_sweep_cells.py [instances_per_cell]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = 100
for r in (1.25, 1.5, 2.0, 3.0, 4.0):
    for d in (1.0, 0.5, 0.1):
        m = int(round(r * n))
        solver.generate_solve_label(1, 0, 256, m, n, density=d)          # warm-up (kernels loaded, scratch sized)
        best = None
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res = solver.generate_solve_label(2 + rep, 0, N, m, n, density=d); e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        st = res['status']
        print('(%d,%d) density %.1f: %.2f ms, %.0f LP/s, optimal %.3f, other status %d, mean pivots %.1f, ties %d' % (
            m, n, d, best, N / best * 1e3, (st == 2).float().mean().item(), int(((st != 2) & (st != 5)).sum()),
            res['pivots'][:, 3].float().mean().item(), int(res['ties'].sum())))
