"""Drop-in for the reference's ``ml.train`` (src/ml/train.py:12-246) without visdom: same ``train_net`` signature and
return value ({'train': [metrics...], 'test': [...]}), same metric names (train.py:228-236), same recall-1 threshold
rule (first ROC threshold with TPR == 1.0 on the train set, :138-140).

What changes: a DataLoader batch is ONE batched forward/backward (``Model.forward_batch``) instead of a Python loop of
single-instance calls; the summed loss gives the same gradient as the reference's accumulation (:60-66).  Under
torch.distributed every rank trains on its shard of the batch and gradients are summed with one flat all-reduce.
Batches are dicts as produced by ``ml.utils.collate_randomlp``; a batch that is a LIST of reference-format items (``DatasetPLNN`` with
``collate_items``: MPS / PLNN LPs of different shapes) takes the reference's own per-item loop."""
import ctypes as C
import time

import numpy as np
import torch
from sklearn.metrics import roc_curve

from .. import _lib, parallel


def _to_device(batch, dev):
    return batch['A'].to(dev), batch['b'].to(dev), batch['c'].to(dev), batch['y'].to(dev)


def collate_items(items):
    """DataLoader ``collate_fn`` for reference-format items of different shapes (``DatasetPLNN``): the batch is the list."""
    return list(items)


def _is_item_batch(data):
    """A batch of reference-format items (MPS / PLNN LPs: one shape per item, equality / bound flags) rather than the
    collated random-LP tensors of ``ml.utils.collate_randomlp``."""
    return isinstance(data, (list, tuple))


def _item_xy(item, graph, dev):
    """(x, y) of one reference-format item as ``batched`` yields them (src/ml/utils.py:3-25): y = labels of the in_loss rows."""
    in_loss = [int(q) for q in item['in_loss']]
    lab = item['node_labels'] if graph == 'complete' else item['c_labels']
    lab = torch.as_tensor(np.asarray(lab)).reshape(-1)
    return item, lab[in_loss].long().to(dev)


def _model_device(model):
    return next(model.parameters()).device


def train_net(model, criterion, optimizer, trainloader, testloader, epochs, batch_size, cuda=False, verbose=True):
    model.train()
    dev = _model_device(model)
    metrics = {'test': [], 'train': []}                        # the reference's return schema (train.py:51, 100)
    train_net.last_thresholds = []                             # recall-1 threshold of every epoch (train.py:138-140), for the parity tests
    train_start = time.time()
    for epoch in range(epochs):
        epoch_start = time.time()
        running_loss = 0.0
        for data in trainloader:
            if _is_item_batch(data):
                # the reference's own loop (train.py:57-66): one forward / backward per item, gradients accumulate
                optimizer.zero_grad()
                for item in data:
                    loss = None
                    if _device_backward_ok_item(model, criterion):
                        # hand-written loss + gradient kernels with the item's row flags (ddb_s2v_loss_grad_flags_dev)
                        try:
                            loss = model.loss_and_grad_item(item, [float(criterion.weight[0]), float(criterion.weight[1])])
                        except _lib.DdbError as exc:
                            if 'do not fit' not in str(exc):
                                raise
                    if loss is None:
                        x, y = _item_xy(item, model.graph, dev)
                        loss = criterion(model(x), y)
                        loss.backward()
                parallel.allreduce_gradients(model)
                optimizer.step()
                running_loss += float(loss.detach())                  # the last item's loss, as train.py:68-69 records it
                continue
            A, b, c, y = _to_device(data, dev)
            optimizer.zero_grad()
            loss = None
            if _device_backward_ok(model, criterion, A):
                # hand-written loss + gradient kernels (csrc/s2v_backward.cu and its general-adjacency / complete-graph
                # siblings): dense instances in one streaming launch, instances with zero coefficients in a second one
                try:
                    loss = model.loss_and_grad_batch(A, b, c, y, [float(criterion.weight[0]), float(criterion.weight[1])])
                except _lib.DdbError as exc:
                    if 'do not fit' not in str(exc):
                        raise
                    model._no_device_backward = True          # shape / embedding size beyond the kernel: autograd from now on
            if loss is None:
                fx = model.forward_batch(A, b, c)                         # [B,m,2] log-probs
                loss = criterion(fx.reshape(-1, 2), y.reshape(-1))        # summed over the batch (benchmark.py:75)
                loss.backward()
            parallel.allreduce_gradients(model)                       # no-op on one rank
            optimizer.step()
            running_loss += float(loss.detach())
        if verbose and parallel.world()[0] == 0:
            print('%d: running_loss %g (epoch %g secs, elapsed %g secs)'
                  % (epoch, running_loss, time.time() - epoch_start, time.time() - train_start))
        p_train = recall_one_threshold(trainloader, model)
        train_net.last_thresholds.append(p_train)
        metrics['train'].append(performance(trainloader, model, criterion, p_train))
        metrics['test'].append(performance(testloader, model, criterion, p_train))
    return metrics


def _device_backward_ok(model, criterion, A):
    """The device backward implements the reference's criterion (weighted NLL, summed: benchmark.py:70-75) for the
    bipartite model on dense instances and for the complete model; anything else goes through autograd."""
    return (hasattr(model, 'device_backward_supported') and not getattr(model, '_no_device_backward', False)
            and model.device_backward_supported(A)
            and isinstance(criterion, torch.nn.NLLLoss) and criterion.reduction == 'sum'
            and criterion.weight is not None and criterion.ignore_index < 0)


def _device_backward_ok_item(model, criterion):
    """The same for one reference-format item (MPS / PLNN): model on the GPU, the reference's criterion."""
    return (hasattr(model, 'loss_and_grad_item') and not getattr(model, 'force_torch', False) and model.p <= 64
            and next(model.parameters()).is_cuda
            and isinstance(criterion, torch.nn.NLLLoss) and criterion.reduction == 'sum'
            and criterion.weight is not None and criterion.ignore_index < 0)


def _probs_and_labels(loader, model):
    dev = _model_device(model)
    ys, ps = [], []
    was_training = model.training
    model.eval()
    with torch.no_grad():
        for data in loader:
            if _is_item_batch(data):
                for item in data:
                    x, y = _item_xy(item, model.graph, dev)
                    model(x)
                    ps.append(model.probs[..., 1].reshape(-1).float().cpu())
                    ys.append(y.reshape(-1).cpu())
                continue
            A, b, c, y = _to_device(data, dev)
            model.forward_batch(A, b, c)
            ps.append(model.probs[..., 1].reshape(-1).float().cpu())
            ys.append(y.reshape(-1).cpu())
    if was_training:
        model.train()
    return torch.cat(ys).numpy(), torch.cat(ps).numpy()


def _loader_has_items(loader):
    """True if the loader yields reference-format item lists (peeks at the dataset, not at a batch)."""
    ds = getattr(loader, 'dataset', None)
    return ds is not None and len(ds) > 0 and isinstance(ds[0], dict) and ('c_feats' in ds[0] or 'node_features' in ds[0])


def _on_device(model):
    return _model_device(model).type == 'cuda' and not getattr(model, 'force_torch', False)


def device_metrics(model, loader, criterion=None, prob_thresh=0.5):
    """One evaluation pass entirely on the device: batched forward kernel + ``ddb_s2v_metrics_dev`` per batch, a single
    8-double read-back at the end.  Returns [tp, fp, tn, fn, min positive prob, weighted NLL sum, #pos, #neg]."""
    dev = _model_device(model)
    ctx = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
    w = criterion.weight if (criterion is not None and getattr(criterion, 'weight', None) is not None) else None
    w0, w1 = (float(w[0]), float(w[1])) if w is not None else (1.0, 1.0)
    acc = torch.zeros(8, dtype=torch.float64, device=dev)
    acc[4] = float('inf')
    out = torch.empty(8, dtype=torch.float64, device=dev)
    was_training = model.training
    model.eval()
    with torch.no_grad():
        for data in loader:
            A, b, c, y = _to_device(data, dev)
            logp = model.forward_batch_cuda(A, b, c)
            y8 = y.to(torch.uint8).contiguous()
            rc = ctx.lib.ddb_s2v_metrics_dev(ctx.handle, y8.numel(), C.c_void_p(logp.data_ptr()),
                                             C.c_void_p(model.probs.data_ptr()), C.c_void_p(y8.data_ptr()),
                                             float(prob_thresh), w0, w1, C.c_void_p(out.data_ptr()),
                                             C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            _lib.check(rc, 'ddb_s2v_metrics_dev')
            mn = torch.minimum(acc[4], out[4])
            acc += out
            acc[4] = mn
    if was_training:
        model.train()
    return acc.cpu().numpy()


def recall_one_threshold(loader, model):
    """train.py:118-150: threshold of the first ROC point whose TPR is 1.0 (keeps every active constraint).  sklearn's
    ROC thresholds are the scores themselves, so that point is the smallest predicted probability of a positive -- which
    is what the device pass returns."""
    if _on_device(model) and not _loader_has_items(loader):
        r = device_metrics(model, loader)
        return float(r[4]) if (r[6] > 0 and r[7] > 0) else 0.5
    y_true, y_prob = _probs_and_labels(loader, model)
    if (y_true == 1).sum() == 0 or (y_true == 0).sum() == 0:
        return 0.5
    fpr, tpr, thresholds = roc_curve(y_true, y_prob, pos_label=1)
    idx = np.where(tpr == 1.0)[0]
    return float(thresholds[idx[0]])


def plot_roc(model, epoch, trainloader=None, testloader=None):
    """Reference signature and return value (train.py:118-172: ``(p_train, p_test)``, the recall-1 thresholds of the two
    loaders); the visdom plot itself is out of scope."""
    p_train = recall_one_threshold(trainloader, model) if trainloader is not None else None
    p_test = recall_one_threshold(testloader, model) if testloader is not None else None
    return p_train, p_test


def get_prob_recall_one(loader, model):
    """train.py:102-116: smallest predicted probability of a positive."""
    if _on_device(model) and not _loader_has_items(loader):
        r = device_metrics(model, loader)
        return float(r[4]) if r[6] > 0 else 0.5
    y_true, y_prob = _probs_and_labels(loader, model)
    return float(y_prob[y_true == 1].min()) if (y_true == 1).any() else 0.5


def performance(loader, model, criterion, prob_thresh=0.5):
    """train.py:174-246 (metric names and formulas unchanged)."""
    if _on_device(model) and isinstance(criterion, torch.nn.NLLLoss) and criterion.reduction == 'sum' and not _loader_has_items(loader):
        r = device_metrics(model, loader, criterion, prob_thresh)
        return _metrics_dict(r[5], int(r[0]), int(r[1]), int(r[2]), int(r[3]))
    dev = _model_device(model)
    was_training = model.training
    model.eval()
    total_loss = 0.0
    tps = fps = tns = fns = 0
    with torch.no_grad():
        for data in loader:
            if _is_item_batch(data):
                for item in data:
                    x, y = _item_xy(item, model.graph, dev)
                    fx = model(x)
                    total_loss += float(criterion(fx, y))
                    pred = model.probs[..., 1] >= prob_thresh
                    tps += int(((y == 1) & pred).sum()); fps += int(((y == 0) & pred).sum())
                    tns += int(((y == 0) & ~pred).sum()); fns += int(((y == 1) & ~pred).sum())
                continue
            A, b, c, y = _to_device(data, dev)
            fx = model.forward_batch(A, b, c)
            total_loss += float(criterion(fx.reshape(-1, 2), y.reshape(-1)))
            pred = model.probs[..., 1] >= prob_thresh
            tps += int(((y == 1) & pred).sum()); fps += int(((y == 0) & pred).sum())
            tns += int(((y == 0) & ~pred).sum()); fns += int(((y == 1) & ~pred).sum())
    if was_training:
        model.train()
    return _metrics_dict(total_loss, tps, fps, tns, fns)


def _metrics_dict(total_loss, tps, fps, tns, fns):
    tot = max(tps + fps + tns + fns, 1)
    return {'total_loss': float(total_loss), 'accuracy': (tps + tns) / tot, 'precision': tps / max(tps + fps, 1),
            'recall': tps / max(tps + fns, 1), 'y_pos': (tps + fns) / tot, 'y_neg': (fps + tns) / tot,
            'pred_pos': (tps + fps) / tot, 'pred_neg': (tns + fns) / tot}


def train_on_device_stream(model, optimizer, m, n, steps, batch_per_rank, key=0, weight=(0.5, 0.5), density=1.0,
                           threshold=None, overlap=True):
    """BASELINE.json config 5: data-parallel classifier training fed by on-GPU LP generation.

    Every step, every rank (one process per GPU) runs the whole hot path on its own shard, nothing touches the host:
      1. fused Philox generate -> solve -> label for instances [(step * W + rank) * B, ... + B) of stream `key`
         (``ddb_generate_solve_label_dev``; the Philox counter is the global instance index, so the data seen by the job
         does not depend on the number of ranks);
      2. loss + gradient of the batch in one launch (``ddb_s2v_loss_grad_dev``: the reference's accumulation loop
         train.py:59-65 with the criterion of benchmark.py:70-75; non-optimal instances carry all-zero labels exactly as
         randomlp_dataset.py:96-102 stores them);
      3. ONE flat all-reduce of the gradient (NCCL over NVLink; 4.6-47 KB, latency-bound) and the optimizer step.
    With ``overlap`` the next step's generate+solve is enqueued on a side stream before the all-reduce, so the collective
    and the classifier kernels hide behind the solver (double-buffered instance tensors).
    Returns {'loss': per-step global mean loss per constraint node, 'lps': LPs consumed by the whole job,
             'seconds': device-timed duration (max over ranks), 'lps_per_sec': ...}."""
    from .. import solver
    dev = _model_device(model)
    if dev.type != 'cuda':
        raise _lib.DdbError('train_on_device_stream runs on the GPU only; there is no CPU fallback')
    rank, W = parallel.world()
    B = int(batch_per_rank)
    thr = _lib.DEFAULT_THRESHOLD if threshold is None else threshold
    parallel.broadcast_parameters(model)
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev) if overlap else main
    bufs = [solver._alloc_outputs(B, m, n, dev) for _ in range(2)]
    for bf in bufs:
        bf['A'] = torch.empty(B, m, n, dtype=torch.float64, device=dev)
        bf['b'] = torch.empty(B, m, dtype=torch.float64, device=dev)
        bf['c'] = torch.empty(B, n, dtype=torch.float64, device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def produce(step):
        bf = bufs[step % 2]
        with torch.cuda.stream(side):
            if step >= 2:
                side.wait_event(consumed[step % 2])
            solver.generate_solve_label(key, (step * W + rank) * B, B, m, n, density=density, threshold=thr, device=dev,
                                        out=bf, instances=(bf['A'], bf['b'], bf['c']))
            ready[step % 2].record(side)

    losses = torch.zeros(steps, dtype=torch.float64, device=dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if W > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    t0.record(main)
    produce(0)
    for step in range(steps):
        bf = bufs[step % 2]
        main.wait_event(ready[step % 2])
        if overlap and step + 1 < steps:
            produce(step + 1)
        optimizer.zero_grad()
        loss = model.loss_and_grad_batch(bf['A'], bf['b'], bf['c'], bf['labels'], weight)
        consumed[step % 2].record(main)
        if W > 1:
            flat = torch.cat([q.grad.reshape(-1) for q in model.parameters()] + [loss.reshape(1).float()])
            torch.distributed.all_reduce(flat)
            off = 0
            for q in model.parameters():
                q.grad.copy_(flat[off:off + q.numel()].view_as(q.grad))
                off += q.numel()
            loss = flat[off].double()
        optimizer.step()
        losses[step] = loss / float(W * B * m)
        if not overlap and step + 1 < steps:
            produce(step + 1)
    t1.record(main)
    torch.cuda.synchronize(dev)
    secs = torch.tensor([t0.elapsed_time(t1) / 1e3], dtype=torch.float64, device=dev)
    if W > 1:
        torch.distributed.all_reduce(secs, op=torch.distributed.ReduceOp.MAX)
    secs = float(secs.item())
    total = steps * W * B
    return {'loss': losses.cpu().numpy(), 'lps': total, 'seconds': secs, 'lps_per_sec': total / secs}
