"""Generates tests/golden/s2v_grad_{bipartite,complete}.npz: loss and parameter gradients of the UNMODIFIED reference
classifier (/root/reference/src/ml/models/s2v.py) under the reference's criterion
NLLLoss(weight=[w0, w1], size_average=False) (src/benchmark.py:70-75), accumulated over a small batch of instances the
way train_net does (src/ml/train.py:60-65).  Run from the repo root:  python tests/golden/make_s2v_grad_golden.py
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
sys.path.insert(0, '/root/reference/src')
from ml.models.s2v import Model          # noqa: E402  (the reference class, unmodified)
from oracle import classifier as oc      # noqa: E402
from oracle import randomlp as orl       # noqa: E402

CASES = [  # (m, n, p, T, seed, batch)
    (10, 5, 12, 4, 0, 3), (50, 20, 12, 3, 1, 4), (50, 20, 13, 1, 2, 2), (24, 24, 8, 2, 3, 3), (200, 100, 40, 3, 4, 2),
    (30, 12, 16, 2, 5, 5)]
WEIGHT = [0.4, 0.6]


def main():
    for graph in ('bipartite', 'complete'):
        out = {'weight': np.array(WEIGHT, dtype=np.float32)}
        for ci, (m, n, p, T, seed, batch) in enumerate(CASES):
            torch.manual_seed(200 + seed)
            with redirect_stdout(io.StringIO()):
                model = Model(graph, p, T, on_cuda=False)
            crit = torch.nn.NLLLoss(weight=torch.tensor(WEIGHT), reduction='sum')
            rs = np.random.RandomState(1000 + seed)
            As, bs, cs, ys = [], [], [], []
            total = 0.0
            model.zero_grad()
            for k in range(batch):
                A, b, c = orl.generate_instance(m, n, 50 * seed + k)
                y = (rs.rand(m) < 0.5).astype(np.int64)
                item = oc.item_complete(A, b, c) if graph == 'complete' else oc.item_bipartite(A, b, c)
                loss = crit(model(item), torch.from_numpy(y))
                loss.backward()
                total += float(loss)
                As.append(A); bs.append(b); cs.append(c); ys.append(y)
            pre = 'case%d_' % ci
            out[pre + 'dims'] = np.array([m, n, p, T, seed, batch])
            out[pre + 'A'], out[pre + 'b'], out[pre + 'c'] = np.stack(As), np.stack(bs), np.stack(cs)
            out[pre + 'y'] = np.stack(ys).astype(np.uint8)
            out[pre + 'loss'] = np.array(total)
            for k, v in model.named_parameters():
                out[pre + 'param_' + k] = v.detach().numpy()
                out[pre + 'grad_' + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
        np.savez_compressed(os.path.join(HERE, 's2v_grad_%s.npz' % graph), **out)
        print(graph, 'written:', len(CASES), 'cases')


if __name__ == '__main__':
    main()
