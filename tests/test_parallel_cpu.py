"""CPU: host logic of the multi-GPU path under gloo (world_size 2) -- sharding, label gather, data-parallel gradient
all-reduce -- and the training loop mirror on the differentiable path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deep_dantzig_b200 import parallel


def test_shard_range_partitions_the_index_range():
    for N in (0, 1, 7, 8, 1000003):
        for W in (1, 2, 3, 8):
            blocks = [parallel.shard_range(N, r, W) for r in range(W)]
            assert blocks[0][0] == 0 and blocks[-1][1] == N
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(W - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    items = list(range(11))
    got = sorted(sum((parallel.shard_round_robin(items, r, 4) for r in range(4)), []))
    assert got == items


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from deep_dantzig_b200.ml.models.s2v import Model
        from oracle import randomlp as orl
        # --- gather of ragged label blocks -------------------------------------------------------------------------
        N = 7
        lo, hi = parallel.shard_range(N, rank, world)
        labels = torch.arange(N * 3, dtype=torch.uint8).reshape(N, 3)[lo:hi].contiguous()
        got = parallel.gather_to_rank0(labels)
        if rank == 0:
            assert torch.equal(got, torch.arange(N * 3, dtype=torch.uint8).reshape(N, 3))
        else:
            assert got is None
        # --- data-parallel gradients == single-process gradients on the whole batch ----------------------------------
        torch.manual_seed(0)
        model = Model('bipartite', 6, 2, verbose_init=False)
        parallel.broadcast_parameters(model)
        insts = [orl.generate_instance(12, 5, s) for s in range(6)]
        A = torch.from_numpy(np.stack([i[0] for i in insts])); b = torch.from_numpy(np.stack([i[1] for i in insts]))
        c = torch.from_numpy(np.stack([i[2] for i in insts]))
        y = (torch.arange(6 * 12).reshape(6, 12) % 3 == 0).long()
        crit = torch.nn.NLLLoss(weight=torch.tensor([0.4, 0.6]), reduction='sum')
        lo, hi = parallel.shard_range(6, rank, world)
        model.zero_grad()
        crit(model.forward_batch(A[lo:hi], b[lo:hi], c[lo:hi]).reshape(-1, 2), y[lo:hi].reshape(-1)).backward()
        parallel.allreduce_gradients(model)
        g_dp = torch.cat([q.grad.reshape(-1) for q in model.parameters()]).clone()
        model.zero_grad()
        crit(model.forward_batch(A, b, c).reshape(-1, 2), y.reshape(-1)).backward()
        g_full = torch.cat([q.grad.reshape(-1) for q in model.parameters()])
        assert torch.allclose(g_dp, g_full, rtol=1e-5, atol=1e-6)
        with open(os.path.join(out_dir, 'ok%d' % rank), 'w') as f:
            f.write('ok')
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / 'ok0').exists() and (tmp_path / 'ok1').exists()


def test_train_net_mirror_learns_on_cpu():
    """train_net keeps the reference's return schema and metric names; the loss goes down on a tiny dataset."""
    from torch.utils.data import DataLoader
    from deep_dantzig_b200.ml.models.s2v import Model
    from deep_dantzig_b200.ml.train import train_net, performance
    from deep_dantzig_b200.ml.utils import collate_randomlp, class_weights
    from oracle import randomlp as orl
    ds = orl.RandomLPDataset(10, 5, num_lps=24, seed=0)          # oracle-labelled (test infrastructure)
    loader = DataLoader(ds, batch_size=8, shuffle=False, collate_fn=collate_randomlp)
    torch.manual_seed(1)
    model = Model('bipartite', 8, 2, verbose_init=False)
    model.force_torch = True                                     # CPU test of the host logic: explicit torch path
    w = torch.tensor(class_weights(ds), dtype=torch.float32)
    crit = torch.nn.NLLLoss(weight=w, reduction='sum')
    before = performance(loader, model, crit, 0.5)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
    hist = train_net(model, crit, opt, loader, loader, epochs=6, batch_size=8, verbose=False)
    assert set(hist) == {'train', 'test'} and len(hist['train']) == 6
    assert set(hist['train'][0]) == {'total_loss', 'accuracy', 'precision', 'recall', 'y_pos', 'y_neg', 'pred_pos', 'pred_neg'}
    assert hist['train'][-1]['total_loss'] < before['total_loss']
    assert hist['train'][-1]['recall'] == 1.0                     # metrics are taken at the recall-1 threshold


def test_evaluation_entry_points_on_cpu():
    """get_accuracy (ml/test.py:10-54), plot_roc's return value (train.py:118-172) and get_prob_recall_one (:102-116)
    keep the reference's signatures; on the CPU path the threshold comes from sklearn's ROC exactly as in the reference."""
    from deep_dantzig_b200.ml.models.s2v import Model
    from deep_dantzig_b200.ml.train import plot_roc, get_prob_recall_one, performance
    from deep_dantzig_b200.ml.test import get_accuracy
    from deep_dantzig_b200.ml.utils import collate_randomlp
    from torch.utils.data import DataLoader
    from oracle import randomlp as orl
    ds = orl.RandomLPDataset(10, 5, num_lps=16, seed=3)
    loader = DataLoader(ds, batch_size=8, shuffle=False, collate_fn=collate_randomlp)
    torch.manual_seed(2)
    model = Model('bipartite', 6, 2, verbose_init=False)
    model.force_torch = True
    p_train, p_test = plot_roc(model, 0, trainloader=loader, testloader=None)
    assert p_test is None and 0.0 < p_train < 1.0
    assert p_train == get_prob_recall_one(loader, model)           # first ROC point with TPR 1 == min prob of a positive
    acc = get_accuracy(loader, model, p_train)
    assert set(acc) == {'accuracy', 'precision', 'recall', 'y_pos', 'y_neg', 'pred_pos', 'pred_neg'} and acc['recall'] == 1.0
    crit = torch.nn.NLLLoss(weight=torch.tensor([0.5, 0.5]), reduction='sum')
    perf = performance(loader, model, crit, p_train)
    assert all(abs(perf[k] - acc[k]) < 1e-12 for k in acc)


def test_numa_binding_is_best_effort():
    """Without a GPU (or without a readable topology) the helper changes nothing and reports None."""
    before = os.sched_getaffinity(0)
    node = parallel.bind_to_gpu_numa_node(0)
    assert node is None or isinstance(node, int)
    if not torch.cuda.is_available():
        assert node is None and os.sched_getaffinity(0) == before


def test_train_net_reference_item_batches_on_cpu():
    """A DataLoader that yields LISTS of reference-format items (what DatasetPLNN + collate_items produce: one shape per
    item) takes the reference's per-item accumulation loop (train.py:57-66); return schema and metrics as for tensors."""
    from torch.utils.data import DataLoader
    from deep_dantzig_b200.ml.models.s2v import Model
    from deep_dantzig_b200.ml.train import train_net, performance, collate_items
    from oracle import randomlp as orl
    from oracle import classifier as oc
    for graph in ('bipartite', 'complete'):
        items = []
        for k, (m, n) in enumerate([(10, 5), (12, 4), (9, 6), (10, 5), (14, 5), (8, 3)]):        # shapes differ per item
            A, b, c = orl.generate_instance(m, n, 100 + k)
            ref = orl.solve_batch(A[None], b[None], c[None])
            labels = ref['labels'][0].astype(int).tolist()
            it = oc.item_bipartite(A, b, c, labels) if graph == 'bipartite' else oc.item_complete(A, b, c, labels)
            if graph == 'bipartite':
                it['c_feats'][0, 0] = 0.0                   # one equality row, kept out of the loss as gurobi_lp.py:164-177 does
                it['in_loss'] = list(range(1, m))
            items.append(it)
        loader = DataLoader(items, batch_size=4, shuffle=False, collate_fn=collate_items)
        torch.manual_seed(2)
        model = Model(graph, 6, 2, verbose_init=False)
        model.force_torch = True
        crit = torch.nn.NLLLoss(weight=torch.tensor([0.3, 0.7]), reduction='sum')
        before = performance(loader, model, crit, 0.5)
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
        hist = train_net(model, crit, opt, loader, loader, epochs=5, batch_size=4, verbose=False)
        assert set(hist) == {'train', 'test'} and len(hist['train']) == 5
        assert hist['train'][-1]['total_loss'] < before['total_loss']
        assert hist['train'][-1]['recall'] == 1.0
        n_rows = sum(len(it['in_loss']) for it in items)
        last = hist['train'][-1]
        assert abs((last['y_pos'] + last['y_neg']) - 1.0) < 1e-12 and abs(last['pred_pos'] + last['pred_neg'] - 1.0) < 1e-12
        assert n_rows > 0
