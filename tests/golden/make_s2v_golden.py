"""Generates tests/golden/s2v_{complete,bipartite}.npz by importing the UNMODIFIED reference classifier
(/root/reference/src/ml/models/s2v.py) in the build container.  The reference tree does not exist on the GPU box, so
the fixtures (parameters, inputs, reference outputs) are committed; this script is how they were made.
Run from the repo root:  python tests/golden/make_s2v_golden.py
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

CASES = [  # (m, n, p, T, seed)
    (10, 5, 12, 4, 0), (50, 20, 12, 3, 1), (50, 20, 13, 1, 2), (24, 24, 8, 2, 3), (200, 100, 40, 3, 4), (30, 12, 16, 2, 5)]


def main():
    for graph in ('complete', 'bipartite'):
        out = {}
        for ci, (m, n, p, T, seed) in enumerate(CASES):
            torch.manual_seed(100 + seed)
            with redirect_stdout(io.StringIO()):
                model = Model(graph, p, T, on_cuda=False)
            A, b, c = orl.generate_instance(m, n, seed)
            if graph == 'bipartite' and ci == 3:
                A = A * (np.random.RandomState(7).rand(m, n) < 0.4)      # a sparse instance: exercises the adjacency
            item = oc.item_complete(A, b, c) if graph == 'complete' else oc.item_bipartite(A, b, c)
            with torch.no_grad():
                logp = model(item)
            pre = 'case%d_' % ci
            out[pre + 'dims'] = np.array([m, n, p, T, seed])
            out[pre + 'A'], out[pre + 'b'], out[pre + 'c'] = A, b, c
            out[pre + 'logp'] = logp.numpy()
            out[pre + 'probs'] = model.probs.numpy()
            for k, v in model.state_dict().items():
                out[pre + 'param_' + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, 's2v_%s.npz' % graph), **out)
        print(graph, 'written:', len(CASES), 'cases')


if __name__ == '__main__':
    main()
