"""BASELINE.json config 5 driver: data-parallel classifier training fed by on-GPU LP generation, one process per GPU
(torchrun), NCCL all-reduce of the flat gradient.  Prints one JSON line on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_stream_dp.py [m n B steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from deep_dantzig_b200.ml.models.s2v import Model
from deep_dantzig_b200.ml import train as tr

m, n, B, steps = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (200, 100, 1024, 40))]
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
torch.manual_seed(0)
model = Model('bipartite', 40, 3, on_cuda=True, verbose_init=False)
# the criterion is SUM-reduced as in the reference (benchmark.py:75): keep the step per instance constant across world sizes
lr = float(os.environ.get("DDB_LR", "1e-6")) / world
opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=0.9)
res = {}
for overlap in (False, True):
    tr.train_on_device_stream(model, opt, m, n, 5, B, key=5, weight=(0.25, 0.75), overlap=overlap)      # warm-up
    h = tr.train_on_device_stream(model, opt, m, n, steps, B, key=6 + overlap, weight=(0.25, 0.75), overlap=overlap)
    res['overlap' if overlap else 'serial'] = {'lps_per_sec': h['lps_per_sec'], 'seconds': h['seconds'],
                                               'loss_first': float(h['loss'][:3].mean()), 'loss_last': float(h['loss'][-3:].mean())}
flat = torch.cat([q.detach().reshape(-1) for q in model.parameters()])
chk = flat.double().sum().reshape(1).clone()
if world > 1:
    lo = chk.clone(); hi = chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    in_sync = bool((lo == hi).item())
else:
    in_sync = True
if rank == 0:
    print(json.dumps({'config': 'DP training fed by on-GPU generation (BASELINE.json configs[4])', 'm': m, 'n': n, 'n_gpus': world,
                      'batch_per_rank': B, 'steps': steps, 'model': 'bipartite p=40 T=3', 'replicas_in_sync': in_sync, 'params_finite': bool(torch.isfinite(flat).all().item()), 'lr': lr, **res}))
if world > 1:
    dist.destroy_process_group()
