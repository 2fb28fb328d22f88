"""CPU restatement of the reference classifier forward (structure2vec, both graph variants) and of the feature
adapter the reference lacks for random LPs.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates, in plain torch fp32 on the CPU, one instance at a time as the reference runs it:
  * ``Model._forward_complete`` + ``_s2v_complete``   -- src/ml/models/s2v.py:124-187, 91-122 (incl. quirk B10: the
    scalar ``t4rc . sum_j relu(t4rc W_mj)`` is what is added to every row embedding; ``t3rc`` is unused);
  * ``Model._forward_bipartite`` + ``_s2v_bipartite`` -- s2v.py:253-323, 218-251 (incl. quirk B9: ``term2`` is laid out
    variables-first while ``term1``/``term3`` are constraints-first);
  * parameter shapes of ``_init_complete`` / ``_init_bipartite`` -- s2v.py:60-89, 189-216.
Pinned by tests/golden/s2v_*.npz, produced by the unmodified reference class (tests/golden/make_s2v_golden.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

COMPLETE_PARAMS = ['t0', 't1', 't2rr', 't2rc', 't2cr', 't3rr', 't3rc', 't3cr', 't4rr', 't4rc', 't4cr', 't6r', 't6c', 't7', 't8']
BIPARTITE_PARAMS = ['t0', 't1c', 't1v', 't2c', 't2v', 't3c', 't3v', 't4c', 't4v', 't6c', 't6v', 't7', 't8']


def param_shapes(graph, p):
    """s2v.py:60-89 (complete) and :189-216 (bipartite)."""
    if graph == 'complete':
        return {'t0': (p, 1), 't1': (p, 1), 't2rr': (p, p), 't2rc': (p, p), 't2cr': (p, p), 't3rr': (p, p),
                't3rc': (p, p), 't3cr': (p, p), 't4rr': (p, 1), 't4rc': (p,), 't4cr': (p,), 't6r': (p, p),
                't6c': (p, p), 't7': (p, p), 't8': (2, 2 * p)}
    if graph == 'bipartite':
        return {'t0': (p, 1), 't1c': (p, 4), 't1v': (p, 1), 't2c': (p, p), 't2v': (p, p), 't3c': (1, p, p),
                't3v': (1, p, p), 't4c': (1, p, 1), 't4v': (1, p, 1), 't6c': (p, p), 't6v': (p, p), 't7': (p, p),
                't8': (2, 2 * p + 4)}
    raise ValueError('Graph not recognised')


def init_params(graph, p, seed=0):
    """Random parameters with the reference's scales (sqrt(1/p), sqrt(1/4), sqrt(1/(2p+4)))."""
    g = torch.Generator().manual_seed(seed)
    C = math.sqrt(1.0 / p)
    out = {}
    for name, shape in param_shapes(graph, p).items():
        w = torch.randn(*shape, generator=g)
        if name in ('t0', 't1', 't1v'):
            scale = 1.0
        elif name == 't1c':
            scale = math.sqrt(1.0 / 4)
        elif name == 't8' and graph == 'bipartite':
            scale = math.sqrt(1.0 / (2 * p + 4))
        else:
            scale = C
        out[name] = (scale * w).float()
    return out


# ---------------------------------------------------------------------------------------------------------------
# feature adapter (layouts from gurobi_lp.py:127-187 (bipartite) and :326-366 (complete))
# ---------------------------------------------------------------------------------------------------------------
def item_complete(A, b, c, labels=None):
    m, n = A.shape
    item = {'A': torch.from_numpy(np.ascontiguousarray(A)).unsqueeze(0), 'b': torch.from_numpy(b).unsqueeze(0),
            'c': torch.from_numpy(c).unsqueeze(0),
            'node_features': torch.cat((torch.ones(1, m), torch.zeros(1, 1)), 1), 'in_loss': list(range(m))}
    if labels is not None:
        item['node_labels'] = torch.tensor(list(labels) + [0]).long()
    return item


def item_bipartite(A, b, c, labels=None):
    m, n = A.shape
    c_feats = torch.zeros(m, 3)
    c_feats[:, 0] = 1.0                                   # is_inequality
    c_feats[:, 1] = torch.from_numpy(b).float()           # rhs
    v_feats = torch.from_numpy(c).float().unsqueeze(1)    # objective coefficient
    rows, cols = np.nonzero(A)
    item = {'c_feats': c_feats, 'v_feats': v_feats,
            'e_feats': {'i': [[int(r), int(q)] for r, q in zip(rows, cols)], 'coeffs': [float(A[r, q]) for r, q in zip(rows, cols)]},
            'in_loss': list(range(m)), 'dims': {'m': m, 'n': n}}
    if labels is not None:
        item['c_labels'] = torch.tensor(list(labels)).float()
    return item


# ---------------------------------------------------------------------------------------------------------------
# forward passes
# ---------------------------------------------------------------------------------------------------------------
def _relu_outer_sum(t, w):
    """sum_j relu(t_k * w_j) for every k  (the bmm -> relu -> sum pattern of s2v.py:112, 236, 239)."""
    return F.relu(t.reshape(-1, 1) * w.reshape(1, -1)).sum(dim=1)


def forward_complete(P, A, b, c, T, in_loss=None):
    """Returns (log_probs, probs), each (len(in_loss), 2) fp32.  A (m,n), b (m,), c (n,) float64 numpy."""
    m, n = A.shape
    in_loss = list(range(m)) if in_loss is None else in_loss
    p = P['t0'].shape[0]
    Ab = F.normalize(torch.from_numpy(np.concatenate([A, b[:, None]], 1)), p=2, dim=1)   # fp64 normalise (s2v.py:145)
    G = torch.cat((Ab.float(), torch.from_numpy(np.concatenate([c, [0.0]])[None]).float()), 0)
    W = G @ G.t()
    W.fill_diagonal_(0.0)                                                                   # s2v.py:158-162
    feat = torch.cat((torch.ones(1, m), torch.zeros(1, 1)), 1)
    mu = torch.zeros(p, m + 1)
    inv_m = 1.0 / float(m)
    for _ in range(T):
        u1 = P['t0'] + P['t1'] @ feat
        u2r = P['t2rr'] @ mu[:, :m] + (P['t2rc'] @ mu[:, [m]]).repeat(1, m)
        u2c = (P['t2cr'] @ (inv_m * mu[:, :m].sum(dim=1))).unsqueeze(1)
        rr = torch.stack([_relu_outer_sum(P['t4rr'], W[i, :m]) for i in range(m)], 1)      # (p, m)
        u3rr = P['t3rr'] @ rr
        rc = _relu_outer_sum(P['t4rc'], W[m, :m])
        u3r = u3rr + (P['t4rc'] @ rc)                                                       # scalar broadcast (B10)
        cr = _relu_outer_sum(P['t4cr'], W[:m, m])
        u3c = (P['t3cr'] @ cr).unsqueeze(1)
        mu = F.relu(u1 + torch.cat((u2r, u2c), 1) + torch.cat((u3r, u3c), 1))
    u6 = P['t6r'] @ (inv_m * mu[:, :m].sum(dim=1)) + P['t6c'] @ mu[:, m]
    u7 = P['t7'] @ mu[:, in_loss]
    feats = F.relu(torch.cat((u6.unsqueeze(1).repeat(1, len(in_loss)), u7), 0))
    scores = (P['t8'] @ feats).t()
    return F.log_softmax(scores, dim=1), F.softmax(scores, dim=1)


def forward_bipartite(P, A, b, c, T, in_loss=None):
    """Returns (log_probs, probs).  Inputs are cast to fp32 first, as the reference's item tensors are (A1)."""
    m, n = A.shape
    in_loss = list(range(m)) if in_loss is None else in_loss
    p = P['t0'].shape[0]
    A32 = torch.from_numpy(np.ascontiguousarray(A)).float()
    adj = (A32 != 0).float()
    rhs = torch.from_numpy(b).float()
    v_feats = torch.from_numpy(c).float().unsqueeze(1)
    Ab = F.normalize(torch.cat((A32, -rhs.unsqueeze(1)), 1), p=2, dim=1)                    # s2v.py:292
    An = Ab[:, :n]
    c_feats = torch.stack((torch.ones(m), -Ab[:, n], torch.zeros(m), An @ v_feats[:, 0]), 1)   # :293-298
    mu = torch.zeros(p, m + n)
    cadj = F.normalize(adj, p=1, dim=0)
    radj = F.normalize(adj.t(), p=1, dim=0)
    t3c, t3v, t4c, t4v = P['t3c'][0], P['t3v'][0], P['t4c'][0, :, 0], P['t4v'][0, :, 0]
    for _ in range(T):
        term1 = P['t0'] + torch.cat((P['t1c'] @ c_feats.t(), P['t1v'] @ v_feats.t()), 1)
        term2 = torch.cat((P['t2c'] @ (mu[:, :m] @ cadj), P['t2v'] @ (mu[:, m:] @ radj)), 1)   # variables first (B9)
        rc = torch.stack([_relu_outer_sum(t4c, An[i]) for i in range(m)], 1)
        rv = torch.stack([_relu_outer_sum(t4v, An[:, j]) for j in range(n)], 1)
        term3 = torch.cat((t3c @ rc, t3v @ rv), 1)
        mu = F.relu(term1 + term2 + term3)
    u6 = P['t6c'] @ (mu[:, :m].sum(dim=1) / float(m)) + P['t6v'] @ (mu[:, m:].sum(dim=1) / float(n))
    embed = F.relu(torch.cat((u6.unsqueeze(1).expand(-1, len(in_loss)), P['t7'] @ mu[:, in_loss]), 0))
    embed = torch.cat((embed, c_feats[in_loss, :].t()), 0)
    scores = (P['t8'] @ embed).t()
    return F.log_softmax(scores, dim=1), F.softmax(scores, dim=1)


def forward(graph, P, A, b, c, T, in_loss=None):
    if graph == 'complete':
        return forward_complete(P, A, b, c, T, in_loss)
    if graph == 'bipartite':
        return forward_bipartite(P, A, b, c, T, in_loss)
    raise ValueError('Graph not recognised')
