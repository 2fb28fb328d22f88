"""Drop-in for the reference's ``ml.test`` (src/ml/test.py:10-54): ``get_accuracy`` -- the confusion-matrix metrics of a
loader at a probability threshold -- on batched loaders; on a CUDA model the counts come from ``ddb_s2v_metrics_dev``
(one streaming pass per batch, one 8-double read-back)."""
import torch

from . import train as _train


def get_accuracy(testloader, model, prob_thresh=0.5):
    """Same keys and formulas as the reference: accuracy, precision, recall, y_pos, y_neg, pred_pos, pred_neg."""
    if _train._on_device(model):
        r = _train.device_metrics(model, testloader, None, prob_thresh)
        tps, fps, tns, fns = int(r[0]), int(r[1]), int(r[2]), int(r[3])
    else:
        dev = _train._model_device(model)
        tps = fps = tns = fns = 0
        was_training = model.training
        model.eval()
        with torch.no_grad():
            for data in testloader:
                A, b, c, y = _train._to_device(data, dev)
                model.forward_batch(A, b, c)
                pred = model.probs[..., 1] >= prob_thresh
                tps += int(((y == 1) & pred).sum()); fps += int(((y == 0) & pred).sum())
                tns += int(((y == 0) & ~pred).sum()); fns += int(((y == 1) & ~pred).sum())
        if was_training:
            model.train()
    res = _train._metrics_dict(0.0, tps, fps, tns, fns)
    del res['total_loss']
    return res
