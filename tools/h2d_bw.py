"""Raw pinned host -> device copy bandwidth of the box (the ceiling of the host-buffer end-to-end figure): tools/h2d_bw.py"""
import torch, time
for mb in (64, 192, 512, 2048, 5075):
    n = mb * (1 << 20) // 8
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    d = torch.empty(n, dtype=torch.float64, device='cuda')
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print('%5d MB: %.1f GB/s' % (mb, 3 * n * 8 / e0.elapsed_time(e1) / 1e6), flush=True)
