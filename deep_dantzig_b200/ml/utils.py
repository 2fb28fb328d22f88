"""Drop-in for the reference's ``ml.utils.batched`` (src/ml/utils.py:3-25) plus the random-LP adapter the reference
lacks: ``collate_randomlp`` turns ``RandomLPDataset`` items into the batch tensors the batched forward
consumes; ``batched`` still yields one ``(x, y)`` per instance for reference-style loops (B8 fixed: the batch index is
honoured for the bipartite layout too)."""
import numpy as np
import torch


def _pick(t, k):
    """k-th instance of a collated tensor; a tensor without a real batch axis is shared by every instance."""
    if t.dim() > 1 and t.shape[0] > 1:
        return t[k]
    return t.squeeze(0) if t.dim() > 1 else t


def batched(data, batch_size, graph_structure):
    """Yield one ``(x, y)`` per instance of a collated DataLoader item, in the item layouts ``Model.forward`` takes
    (reference src/ml/utils.py:3-25; the bipartite branch honours the batch index)."""
    if graph_structure not in ('complete', 'bipartite'):
        raise ValueError('graph_structure not recognised')
    in_loss = [int(q) for q in data['in_loss']]
    for k in range(batch_size):
        if graph_structure == 'complete':
            lp = data['lp']
            x = {'A': lp['A'][k].unsqueeze(0), 'b': lp['b'][k].unsqueeze(0), 'c': lp['c'][k].unsqueeze(0),
                 'node_features': data['node_features'], 'in_loss': in_loss}
            labels = _pick(data['node_labels'], k)
        else:
            cf, vf = data['c_feats'], data['v_feats']
            x = {'c_feats': cf[k] if cf.dim() > 2 and cf.shape[0] > 1 else (cf.squeeze(0) if cf.dim() > 2 else cf),
                 'v_feats': vf[k] if vf.dim() > 2 and vf.shape[0] > 1 else (vf.squeeze(0) if vf.dim() > 2 else vf),
                 'e_feats': data['e_feats'], 'dims': data['dims'], 'in_loss': in_loss}
            labels = _pick(data['c_labels'], k)
        yield x, labels[in_loss]


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
