"""Generates tests/golden/s2v_plnn_items.npz: outputs of the UNMODIFIED reference classifier (/root/reference/src/ml/models/
s2v.py) on PLNN-style items -- bipartite items with equality rows and bound rows (c_feats flags), complete items with 0/1 node
features -- which is what the MPS ingestion path (deep_dantzig_b200/data/{mps,mps2numpy,plnn_dataset}.py) feeds the model.
Run from the repo root:  python tests/golden/make_plnn_golden.py"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference/src')
from ml.models.s2v import Model          # noqa: E402  (the reference class, unmodified)


def main():
    out = {}
    rs = np.random.RandomState(11)
    cases = [(9, 5, 6, 3), (24, 10, 12, 4), (16, 16, 8, 2)]
    for ci, (m, n, p, T) in enumerate(cases):
        A = rs.randn(m, n)
        A[rs.rand(m, n) < 0.5] = 0.0
        A[:, 0] = 1.0
        A[0, :] = 0.5
        cf = np.stack([(rs.rand(m) < 0.7), rs.randn(m), (rs.rand(m) < 0.3)], 1).astype(np.float32)
        vf = np.abs(rs.randn(n, 1)).astype(np.float32)
        idx = [[i, j] for i in range(m) for j in range(n) if A[i, j] != 0]
        in_loss = [i for i in range(m) if cf[i, 0] == 1 and cf[i, 2] == 0]
        item = {'c_feats': torch.from_numpy(cf.copy()), 'v_feats': torch.from_numpy(vf),
                'e_feats': {'i': idx, 'coeffs': [float(A[i, j]) for i, j in idx]}, 'in_loss': in_loss, 'dims': {'m': m, 'n': n}}
        torch.manual_seed(40 + ci)
        with redirect_stdout(io.StringIO()):
            model = Model('bipartite', p, T, on_cuda=False)
        logp = model(item).detach().numpy()
        pre = 'bip%d_' % ci
        out[pre + 'dims'] = np.array([m, n, p, T]); out[pre + 'A'] = A; out[pre + 'c_feats'] = cf; out[pre + 'v_feats'] = vf
        out[pre + 'in_loss'] = np.array(in_loss); out[pre + 'logp'] = logp
        for k, v in model.named_parameters():
            out[pre + 'param_' + k] = v.detach().numpy()
        # complete graph: 0/1 node features (1 = inequality row), trailing 0 for the cost node
        b = rs.randn(m); c = np.abs(rs.randn(n))
        nf = np.concatenate(((rs.rand(m) < 0.6).astype(np.float32), [0.0]))
        in_loss_c = [i for i in range(m) if nf[i] == 1]
        item = {'A': torch.from_numpy(A).unsqueeze(0), 'b': torch.from_numpy(b).unsqueeze(0), 'c': torch.from_numpy(c).unsqueeze(0),
                'node_features': torch.from_numpy(nf).unsqueeze(0), 'in_loss': in_loss_c}
        torch.manual_seed(60 + ci)
        with redirect_stdout(io.StringIO()):
            model = Model('complete', p, T, on_cuda=False)
        logp = model(item).detach().numpy()
        pre = 'cmp%d_' % ci
        out[pre + 'dims'] = np.array([m, n, p, T]); out[pre + 'A'] = A; out[pre + 'b'] = b; out[pre + 'c'] = c
        out[pre + 'node_features'] = nf; out[pre + 'in_loss'] = np.array(in_loss_c); out[pre + 'logp'] = logp
        for k, v in model.named_parameters():
            out[pre + 'param_' + k] = v.detach().numpy()
    out['ncases'] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, 's2v_plnn_items.npz'), **out)
    print('written', len(cases), 'bipartite and complete PLNN-style items')


if __name__ == '__main__':
    main()
