"""Throughput of the thread-block-cluster kernel (plan 6) against the global-memory plan (plan 2) at the shapes beyond one SM."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_dantzig_b200 import solver, _lib
ctx = _lib.context(0)
for (m, n, B) in [(500, 250, 1184), (300, 150, 4736), (400, 100, 4736), (600, 300, 592)]:
    A, b, c = solver.generate(7, 0, B, m, n)
    out = solver._alloc_outputs(B, m, n, A.device)
    rec = {'shape': [m, n], 'B': B}
    for plan in (6, 2):
        try:
            ctx.set_solve_plan(plan)
            solver.solve_label(A, b, c, out=out); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); solver.solve_label(A, b, c, out=out); e1.record(); torch.cuda.synchronize()
            rec['plan%d_lps' % plan] = B / e0.elapsed_time(e1) * 1e3
            rec['plan%d_optimal' % plan] = int((out['status'] == 2).sum()); rec['plan%d_flagged_or_other' % plan] = int(((out['status'] != 2) & (out['status'] != 5)).sum())
            rec['plan%d_mean_pivots' % plan] = float(out['pivots'][:, 3].float().mean())
        finally:
            ctx.set_solve_plan(-1)
    print(json.dumps(rec), flush=True)
