"""Drop-in for the reference's ``ml.utils.batched`` (src/ml/utils.py:3-25) plus the random-LP adapter the reference
lacks: ``collate_randomlp`` turns ``RandomLPDataset`` items into the batch tensors the batched forward
consumes; ``batched`` still yields one ``(x, y)`` per instance for reference-style loops (B8 fixed: the batch index is
honoured for the bipartite layout too)."""
import numpy as np
import torch


def batched(data, batch_size, graph_structure):
    if graph_structure == 'complete':
        for batch in range(batch_size):
            x = {}
            x['A'] = data['lp']['A'][batch, :, :].unsqueeze(0)
            x['b'] = data['lp']['b'][batch, :].unsqueeze(0)
            x['c'] = data['lp']['c'][batch, :].unsqueeze(0)
            x['node_features'] = data['node_features']
            x['in_loss'] = [int(p) for p in data['in_loss']]
            labels = data['node_labels']
            labels = labels[batch] if labels.dim() > 1 and labels.shape[0] > 1 else labels.squeeze(0)
            yield x, labels[x['in_loss']]
    elif graph_structure == 'bipartite':
        for batch in range(batch_size):
            pick = (lambda t: t[batch] if t.dim() > 2 and t.shape[0] > 1 else t.squeeze(0))
            x = {}
            x['c_feats'] = pick(data['c_feats'])
            x['v_feats'] = pick(data['v_feats'])
            x['e_feats'] = data['e_feats']
            x['dims'] = data['dims']
            x['in_loss'] = [int(p) for p in data['in_loss']]
            labels = data['c_labels']
            labels = labels[batch] if labels.dim() > 1 and labels.shape[0] > 1 else labels.squeeze(0)
            yield x, labels[x['in_loss']]
    else:
        raise(ValueError('graph_structure not recognised'))


def collate_randomlp(items):
    """list of RandomLPDataset items ({'lp': {'A','b','c'}, 'labels': [(i, label)...]}, randomlp_dataset.py:48-50)
    -> {'A': [B,m,n] f64, 'b': [B,m], 'c': [B,n], 'y': [B,m] int64}."""
    A = torch.from_numpy(np.stack([it['lp']['A'] for it in items]))
    b = torch.from_numpy(np.stack([it['lp']['b'] for it in items]))
    c = torch.from_numpy(np.stack([it['lp']['c'] for it in items]))
    y = torch.tensor([[lab for _, lab in it['labels']] for it in items], dtype=torch.long)
    return {'A': A, 'b': b, 'c': c, 'y': y}


def class_weights(dataset):
    """[n_pos/n_tot, n_neg/n_tot] as the reference weights its NLLLoss (benchmark.py:62-75)."""
    pos = tot = 0
    for k in range(len(dataset)):
        labs = [lab for _, lab in dataset[k]['labels']]
        pos += sum(labs)
        tot += len(labs)
    tot = max(tot, 1)
    return [pos / tot, (tot - pos) / tot]
