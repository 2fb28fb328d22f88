import os, sys, json, time, torch
sys.path.insert(0, '.')
from deep_dantzig_b200 import solver
B = 262144
out = solver._alloc_outputs(B, 200, 100, torch.device('cuda', 0))
solver.generate_solve_label(1, 0, 8192, 200, 100, out=None); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(2):
    e0.record(); solver.generate_solve_label(1, 0, B, 200, 100, out=out); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({'inkernel': os.environ.get('DDB_FUSED_INKERNEL', '1'), 'chunk_mb': os.environ.get('DDB_FUSED_CHUNK_MB', '512'), 'B': B, 'ms': best, 'lps_per_sec': B / best * 1e3,
                  'optimal': int((out['status'] == 2).sum())}))
