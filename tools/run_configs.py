"""BASELINE.json configs 3 and 4 at their stated sizes (bounded), one JSON line each.
  config 3: m/n-ratio x density sweep at n = 100 (phase_transitions.sweep_ratio_density), instances sharded over the ranks
  config 4: classifier inference + certified reduced-LP solve at (500,250) (reduced.timing_forward_pass)
Run on one GPU:  python tools/run_configs.py        (or under torchrun for config 3 on several GPUs)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from deep_dantzig_b200 import solver, reduced
from deep_dantzig_b200.ml.models.s2v import Model
from deep_dantzig_b200.ml import train as tr
from deep_dantzig_b200.phase_transitions import sweep_ratio_density

world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))

# ---- config 3 ----------------------------------------------------------------------------------------------------------
per_cell = int(os.environ.get('DDB_PER_CELL', '10000'))
from deep_dantzig_b200.phase_transitions import warm_up_sweep
warm_up_sweep(n=100, device=local)                                                  # warm-up: every cell's kernels loaded, scratch allocated, on every rank
torch.cuda.synchronize(); t0 = time.perf_counter()
sw = sweep_ratio_density(n=100, per_cell=per_cell, chunk=2048, key=300, device=local)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
if rank == 0:
    cells = {'%g/%g' % k: v for k, v in sw.items()}
    print(json.dumps({'config': 'BASELINE.json configs[2]: m/n ratio x density sweep at n=100', 'n_gpus': world, 'instances_per_cell': per_cell,
                      'cells': len(cells), 'seconds': dt, 'lps_per_sec': per_cell * len(cells) / dt, 'results (ratio/density)': cells}))

# ---- config 4 (rank 0 only) ----------------------------------------------------------------------------------------------
if rank == 0:
    torch.manual_seed(0)
    model = Model('bipartite', 40, 3, on_cuda=True, verbose_init=False)
    opt = torch.optim.SGD(model.parameters(), lr=2e-6, momentum=0.9)
    hist = tr.train_on_device_stream(model, opt, 100, 50, steps=400, batch_per_rank=2048, key=41, weight=(0.25, 0.75)) if world == 1 else None
    m, n, B = 500, 250, int(os.environ.get('DDB_C4_BATCH', '1184'))
    A, b, c = solver.generate(42, 0, B, m, n, device=local)
    y = solver.solve_label(A, b, c)['labels'].long()
    thr = tr.recall_one_threshold([{'A': A, 'b': b, 'c': c, 'y': y}], model)           # calibration batch
    A2, b2, c2 = solver.generate(43, 0, B, m, n, device=local)
    reduced.timing_forward_pass(model, A2[:64], b2[:64], c2[:64], thr)                 # warm-up
    t = reduced.timing_forward_pass(model, A2, b2, c2, thr)
    t['config'] = 'BASELINE.json configs[3]: classifier inference + certified reduced-LP solve at (500,250)'
    t['model'] = 'bipartite p=40 T=3, %s' % ('trained 400 steps x 2048 LPs at (100,50) by train_on_device_stream' if hist else 'random init')
    if hist:
        t['train_loss_first_last'] = [float(hist['loss'][:5].mean()), float(hist['loss'][-5:].mean())]
    print(json.dumps(t))
    # what the reduced solve can give independently of the classifier's quality: an IDEAL mask -- the true active rows plus 20 %
    # random rows -- through the same reduced solve + certificate
    full = solver.solve_label(A2, b2, c2)
    rng = torch.Generator(device=A2.device).manual_seed(5)
    ideal = torch.maximum(full['labels'], (torch.rand(B, m, device=A2.device, generator=rng) < 0.2).to(torch.uint8)).contiguous()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    solver.solve_label(A2, b2, c2, row_mask=ideal); torch.cuda.synchronize()
    ev[0].record(); full = solver.solve_label(A2, b2, c2); ev[1].record()
    ev[2].record(); red = solver.solve_label(A2, b2, c2, row_mask=ideal); ev[3].record(); torch.cuda.synchronize()
    okf = full['status'] == 2
    cert = (red['status'] == 2) & (red['violations'] == 0)
    print(json.dumps({'config': 'BASELINE.json configs[3], ideal mask (true active rows + 20 % random rows) at (500,250)', 'instances': B,
                      'rows_kept_fraction': float(ideal.float().mean()), 'full_solve_ms': ev[0].elapsed_time(ev[1]), 'reduced_solve_ms': ev[2].elapsed_time(ev[3]),
                      'speedup': ev[0].elapsed_time(ev[1]) / ev[2].elapsed_time(ev[3]), 'full_lps': B / ev[0].elapsed_time(ev[1]) * 1e3,
                      'reduced_lps': B / ev[2].elapsed_time(ev[3]) * 1e3, 'certified_of_optimal': int((cert & okf).sum()), 'optimal': int(okf.sum()),
                      'labels_equal_on_certified': bool((red['labels'][cert & okf] == full['labels'][cert & okf]).all()),
                      'mean_pivots_full': float(full['pivots'][:, 3].float().mean()), 'mean_pivots_reduced': float(red['pivots'][:, 3].float().mean())}))
if world > 1:
    dist.destroy_process_group()
