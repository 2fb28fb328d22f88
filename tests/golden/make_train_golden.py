"""Generates tests/golden/train_trajectory_{bipartite,complete}.npz: the training trajectory of the UNMODIFIED reference
classifier (/root/reference/src/ml/models/s2v.py) on BASELINE.json configs[0] data under the reference's own training
loop, restated here because ml/train.py cannot be imported (visdom):

  * loop          -- src/ml/train.py:49-71: per DataLoader batch  zero_grad; for every instance of the batch
                     loss = criterion(model(x), y); loss.backward();  then ONE optimizer.step()
  * criterion     -- src/benchmark.py:70-77: NLLLoss(weight=[n_pos/n_tot, n_neg/n_tot], size_average=False),
                     SGD(lr, momentum, weight_decay)
  * hyper-params  -- src/run.py:58-72: p = 12, T = 4, lr = 0.001, momentum = 0.9, weight_decay = 0, 4 epochs
  * evaluation    -- src/ml/train.py:118-150 (recall-1 threshold = first ROC threshold with TPR == 1 on the TRAIN set) and
                     :174-246 (performance at that threshold: total_loss + confusion counts), after every epoch
  * data          -- configs[0]: (50,20), seed 3231 -> seeds 3231 + 578 i; labels from tests/golden/randomlp_config1.npz;
                     instances 0..639 train, 640..799 test; DataLoader order without shuffling, batch_size 16.

The fixture holds the initial parameters, the per-epoch metrics and thresholds and the final parameters; the gated GPU
test trains the mirror (deep_dantzig_b200.ml.train.train_net) from the same parameters on the same data in the same order
and must land on the same trajectory.   Run from the repo root:  python tests/golden/make_train_golden.py
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import torch
from sklearn.metrics import roc_curve

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
sys.path.insert(0, '/root/reference/src')
from ml.models.s2v import Model          # noqa: E402  (the reference class, unmodified)
from oracle import classifier as oc      # noqa: E402
from oracle import randomlp as orl       # noqa: E402

M, N, P, T = 50, 20, 12, 4
LR, MOMENTUM, WD, EPOCHS, BATCH = 0.001, 0.9, 0.0, 4, 16
NTRAIN, NTEST = 640, 160


def load_data():
    g = np.load(os.path.join(HERE, 'randomlp_config1.npz'))
    seeds = [int(s) for s in g['seeds'][:NTRAIN + NTEST]]
    labels = np.unpackbits(g['labels_packed'], axis=1)[:NTRAIN + NTEST, :M].astype(np.int64)
    inst = [orl.generate_instance(M, N, s) for s in seeds]
    return seeds, inst, labels


def evaluate(model, items, ys, crit, thresh):
    """train.py:174-246 at a given threshold (instance by instance, as the reference does)."""
    total, tp, fp, tn, fn = 0.0, 0, 0, 0, 0
    with torch.no_grad():
        for item, y in zip(items, ys):
            fx = model(item)
            total += float(crit(fx, y))
            pred = (model.probs[:, 1] >= thresh).numpy()
            yy = y.numpy()
            tp += int(((yy == 1) & pred).sum()); fp += int(((yy == 0) & pred).sum())
            tn += int(((yy == 0) & ~pred).sum()); fn += int(((yy == 1) & ~pred).sum())
    return total, tp, fp, tn, fn


def recall_one_threshold(model, items, ys):
    """train.py:118-150: roc_curve over all train nodes, first threshold with tpr == 1.0."""
    probs, true = [], []
    with torch.no_grad():
        for item, y in zip(items, ys):
            model(item)
            probs.append(model.probs[:, 1].numpy().copy()); true.append(y.numpy())
    fpr, tpr, thr = roc_curve(np.concatenate(true), np.concatenate(probs), pos_label=1)
    return float(thr[np.where(tpr == 1.0)][0])


def main():
    torch.set_num_threads(1)              # one thread: the float32 reductions of the reference run in one fixed order
    seeds, inst, labels = load_data()
    for graph in ('bipartite', 'complete'):
        mk = oc.item_complete if graph == 'complete' else oc.item_bipartite

        class Items(object):
            """A FRESH item per access, as a DataLoader hands the reference: Model._forward_bipartite overwrites the rhs
            feature of the item it is given in place (s2v.py:294), so an item object must not be fed twice."""
            def __init__(self, lo, hi): self.lo, self.hi = lo, hi
            def __len__(self): return self.hi - self.lo
            def __getitem__(self, i): return mk(*inst[self.lo + i])
            def __iter__(self): return (self[i] for i in range(len(self)))
        ys = [torch.from_numpy(labels[i]) for i in range(len(inst))]
        tr_items, tr_y, te_items, te_y = Items(0, NTRAIN), ys[:NTRAIN], Items(NTRAIN, NTRAIN + NTEST), ys[NTRAIN:]
        npos = int(labels[:NTRAIN].sum()); ntot = NTRAIN * M
        weight = [npos / ntot, (ntot - npos) / ntot]                     # benchmark.py:66-68 (trainset.weight)
        torch.manual_seed(7)
        with redirect_stdout(io.StringIO()):
            model = Model(graph, P, T, on_cuda=False)
        out = {'dims': np.array([M, N, P, T, EPOCHS, BATCH, NTRAIN, NTEST]), 'hyper': np.array([LR, MOMENTUM, WD]),
               'weight': np.array(weight, dtype=np.float32), 'seeds': np.array(seeds)}
        for k, v in model.named_parameters():
            out['init_' + k] = v.detach().numpy().copy()
        crit = torch.nn.NLLLoss(weight=torch.tensor(weight, dtype=torch.float32), reduction='sum')
        opt = torch.optim.SGD(model.parameters(), lr=LR, momentum=MOMENTUM, weight_decay=WD)
        model.train()
        hist = []
        for epoch in range(EPOCHS):
            running = 0.0
            for lo in range(0, NTRAIN, BATCH):
                opt.zero_grad()
                batch_loss = 0.0
                for i in range(lo, min(lo + BATCH, NTRAIN)):
                    loss = crit(model(tr_items[i]), tr_y[i])
                    loss.backward()
                    batch_loss += float(loss)
                opt.step()
                running += batch_loss
            thr = recall_one_threshold(model, tr_items, tr_y)
            tr = evaluate(model, tr_items, tr_y, crit, thr)
            te = evaluate(model, te_items, te_y, crit, thr)
            hist.append([running, thr] + list(tr) + list(te))
            print(graph, 'epoch', epoch, 'running %.4f thr %.6g train loss %.4f (tp fp tn fn %s) test loss %.4f (%s)'
                  % (running, thr, tr[0], tr[1:], te[0], te[1:]))
        out['history'] = np.array(hist, dtype=np.float64)    # [epoch][running, thr, train(loss,tp,fp,tn,fn), test(...)]
        for k, v in model.named_parameters():
            out['final_' + k] = v.detach().numpy().copy()
        np.savez_compressed(os.path.join(HERE, 'train_trajectory_%s.npz' % graph), **out)


if __name__ == '__main__':
    main()
