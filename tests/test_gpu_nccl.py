"""Two-rank NCCL tests of the multi-GPU path (skipped on a one-GPU box): instances sharded by global Philox index with a
label/status gather, the sweep's all-reduce, and data-parallel classifier training with one flat gradient all-reduce per
step -- results must not depend on the number of ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from deep_dantzig_b200 import parallel, solver, phase_transitions
        from deep_dantzig_b200.ml.models.s2v import Model
        from deep_dantzig_b200.ml import train as tr
        # --- solve path: contiguous shard of the global instance range, gather of labels / status over NCCL ---------------
        N, m, n = 3001, 50, 20
        lo, hi = parallel.shard_range(N, rank, world)
        r = solver.generate_solve_label(77, lo, hi - lo, m, n, device=rank)
        labels = parallel.gather_to_rank0(r['labels'])
        status = parallel.gather_to_rank0(r['status'])
        # --- config 3: (cell, chunk) work items dealt round-robin, statistics summed with one all-reduce ---------------------
        cells = phase_transitions.sweep_ratio_density(n=20, ratios=(1.5, 2.0, 3.0), densities=(1.0, 0.5), per_cell=1500, chunk=256,
                                                      key=9, device=rank)
        # --- config 5: data-parallel training fed by on-GPU generation, one flat all-reduce per step -------------------------
        torch.manual_seed(0)
        model = Model('bipartite', 12, 3, on_cuda=True, verbose_init=False)
        opt = torch.optim.SGD(model.parameters(), lr=1e-6, momentum=0.9)      # same GLOBAL batch for every world size: same step
        h = tr.train_on_device_stream(model, opt, m, n, 12, 512 // world, key=5, weight=(0.25, 0.75))
        flat = torch.cat([q.detach().reshape(-1) for q in model.parameters()])
        lo_, hi_ = flat.clone(), flat.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        in_sync = bool((lo_ == hi_).all().item())
        if rank == 0:
            np.savez(os.path.join(out_dir, 'nccl_w%d.npz' % world), labels=labels.cpu().numpy(), status=status.cpu().numpy(),
                     cells=np.array([[v['instances'], v['optimal'], v['unbounded'], v['mean_pivots']] for v in cells.values()]),
                     loss=h['loss'], params=flat.cpu().numpy(), in_sync=in_sync)
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_nccl_match_one_rank(cuda_device, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    mp.spawn(_worker, args=(1, _free_port(), str(tmp_path)), nprocs=1, join=True)
    two, one = np.load(tmp_path / 'nccl_w2.npz'), np.load(tmp_path / 'nccl_w1.npz')
    assert bool(two['in_sync'])
    assert (two['labels'] == one['labels']).all() and (two['status'] == one['status']).all()      # rank-count independent results
    assert (two['cells'][:, :3] == one['cells'][:, :3]).all()
    assert np.allclose(two['cells'][:, 3], one['cells'][:, 3], rtol=1e-12)
    # same global batch per step (2 x 256 = 1 x 512 instances, same Philox indices): same loss trajectory and parameters up to
    # the fp32 summation order of the gradient
    assert np.allclose(two['loss'], one['loss'], rtol=1e-5)
    assert np.abs(two['params'] - one['params']).max() <= 1e-5 * max(1.0, np.abs(one['params']).max())
