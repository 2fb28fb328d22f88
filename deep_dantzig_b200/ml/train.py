"""Drop-in for the reference's ``ml.train`` (src/ml/train.py:12-246) without visdom: same ``train_net`` signature and
return value ({'train': [metrics...], 'test': [...]}), same metric names (train.py:228-236), same recall-1 threshold
rule (first ROC threshold with TPR == 1.0 on the train set, :138-140).

What changes: a DataLoader batch is ONE batched forward/backward (``Model.forward_batch``) instead of a Python loop of
single-instance calls; the summed loss gives the same gradient as the reference's accumulation (:60-66).  Under
torch.distributed every rank trains on its shard of the batch and gradients are summed with one flat all-reduce.
Batches are dicts as produced by ``ml.utils.collate_randomlp``."""
import time

import numpy as np
import torch
from sklearn.metrics import roc_curve

from .. import parallel


def _to_device(batch, dev):
    return batch['A'].to(dev), batch['b'].to(dev), batch['c'].to(dev), batch['y'].to(dev)


def _model_device(model):
    return next(model.parameters()).device


def train_net(model, criterion, optimizer, trainloader, testloader, epochs, batch_size, cuda=False, verbose=True):
    model.train()
    dev = _model_device(model)
    metrics = {'test': [], 'train': []}
    train_start = time.time()
    for epoch in range(epochs):
        epoch_start = time.time()
        running_loss = 0.0
        for data in trainloader:
            A, b, c, y = _to_device(data, dev)
            optimizer.zero_grad()
            if _device_backward_ok(model, criterion, A):
                # hand-written loss + gradient kernel (csrc/s2v_backward.cu): one launch per batch
                loss = model.loss_and_grad_batch(A, b, c, y, [float(criterion.weight[0]), float(criterion.weight[1])])
            else:
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
        metrics['train'].append(performance(trainloader, model, criterion, p_train))
        metrics['test'].append(performance(testloader, model, criterion, p_train))
    return metrics


def _device_backward_ok(model, criterion, A):
    """The device backward implements the reference's criterion (weighted NLL, summed: benchmark.py:70-75) for the
    bipartite model on dense instances; anything else goes through autograd."""
    return (hasattr(model, 'device_backward_supported') and model.device_backward_supported(A)
            and isinstance(criterion, torch.nn.NLLLoss) and criterion.reduction == 'sum'
            and criterion.weight is not None and criterion.ignore_index < 0 and bool((A != 0).all()))


def _probs_and_labels(loader, model):
    dev = _model_device(model)
    ys, ps = [], []
    was_training = model.training
    model.eval()
    with torch.no_grad():
        for data in loader:
            A, b, c, y = _to_device(data, dev)
            model.forward_batch(A, b, c)
            ps.append(model.probs[..., 1].reshape(-1).float().cpu())
            ys.append(y.reshape(-1).cpu())
    if was_training:
        model.train()
    return torch.cat(ys).numpy(), torch.cat(ps).numpy()


def recall_one_threshold(loader, model):
    """train.py:118-150: threshold of the first ROC point whose TPR is 1.0 (keeps every active constraint)."""
    y_true, y_prob = _probs_and_labels(loader, model)
    if (y_true == 1).sum() == 0 or (y_true == 0).sum() == 0:
        return 0.5
    fpr, tpr, thresholds = roc_curve(y_true, y_prob, pos_label=1)
    idx = np.where(tpr == 1.0)[0]
    return float(thresholds[idx[0]])


def get_prob_recall_one(loader, model):
    """train.py:102-116: smallest predicted probability of a positive."""
    y_true, y_prob = _probs_and_labels(loader, model)
    return float(y_prob[y_true == 1].min()) if (y_true == 1).any() else 0.5


def performance(loader, model, criterion, prob_thresh=0.5):
    """train.py:174-246 (metric names and formulas unchanged)."""
    dev = _model_device(model)
    was_training = model.training
    model.eval()
    total_loss = 0.0
    tps = fps = tns = fns = 0
    with torch.no_grad():
        for data in loader:
            A, b, c, y = _to_device(data, dev)
            fx = model.forward_batch(A, b, c)
            total_loss += float(criterion(fx.reshape(-1, 2), y.reshape(-1)))
            pred = model.probs[..., 1] >= prob_thresh
            tps += int(((y == 1) & pred).sum()); fps += int(((y == 0) & pred).sum())
            tns += int(((y == 0) & ~pred).sum()); fns += int(((y == 1) & ~pred).sum())
    if was_training:
        model.train()
    tot = max(tps + fps + tns + fns, 1)
    return {'total_loss': total_loss, 'accuracy': (tps + tns) / tot, 'precision': tps / max(tps + fps, 1),
            'recall': tps / max(tps + fns, 1), 'y_pos': (tps + fns) / tot, 'y_neg': (fps + tns) / tot,
            'pred_pos': (tps + fps) / tot, 'pred_neg': (tns + fns) / tot}
