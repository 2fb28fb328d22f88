"""CPU: classifier oracle and the host-side Model mirror against fixtures produced by the UNMODIFIED reference class
(tests/golden/make_s2v_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import classifier as oc

TOL = 1e-5      # fp32 re-association only (reference log-probs are O(1..7))


def _cases(golden_dir, graph):
    g = np.load(os.path.join(golden_dir, 's2v_%s.npz' % graph))
    for ci in range(6):
        pre = 'case%d_' % ci
        m, n, p, T, seed = [int(v) for v in g[pre + 'dims']]
        P = {k[len(pre + 'param_'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre + 'param_')}
        yield (m, n, p, T), P, g[pre + 'A'], g[pre + 'b'], g[pre + 'c'], g[pre + 'logp'], g[pre + 'probs']


@pytest.mark.parametrize('graph', ['complete', 'bipartite'])
def test_oracle_restatement_matches_reference(golden_dir, graph):
    for dims, P, A, b, c, logp, probs in _cases(golden_dir, graph):
        assert {k: tuple(v.shape) for k, v in P.items()} == oc.param_shapes(graph, dims[2])
        lp, pr = oc.forward(graph, P, A, b, c, dims[3])
        assert np.abs(lp.numpy() - logp).max() <= TOL, dims
        assert np.abs(pr.numpy() - probs).max() <= TOL, dims
        assert lp.shape == (dims[0], 2)


@pytest.mark.parametrize('graph', ['complete', 'bipartite'])
def test_model_mirror_loads_reference_state_and_matches(golden_dir, graph):
    from deep_dantzig_b200.ml.models.s2v import Model
    for dims, P, A, b, c, logp, probs in _cases(golden_dir, graph):
        model = Model(graph, dims[2], dims[3], verbose_init=False)
        model.load_state_dict(P)                       # reference state_dicts load unchanged (names + shapes)
        with torch.no_grad():
            lp = model.forward_batch_torch(torch.from_numpy(A)[None], torch.from_numpy(b)[None], torch.from_numpy(c)[None])
        assert np.abs(lp[0].numpy() - logp).max() <= TOL, dims
        assert np.abs(model.probs[0].numpy() - probs).max() <= TOL
        # parameter block layout of the C ABI
        flat = model.flat_params()
        assert flat.numel() == sum(v.numel() for v in P.values())
        off = 0
        for k in model._names:
            assert torch.equal(flat[off:off + P[k].numel()], P[k].reshape(-1))
            off += P[k].numel()


def test_param_counts_and_errors():
    from deep_dantzig_b200.ml.models.s2v import Model
    assert sum(q.numel() for q in Model('complete', 12, 3, verbose_init=False).parameters()) == 1404
    assert sum(q.numel() for q in Model('bipartite', 12, 3, verbose_init=False).parameters()) == 1160
    with pytest.raises(ValueError):
        Model('hypergraph', 4, 1, verbose_init=False)
    with pytest.raises(ValueError):
        oc.forward('hypergraph', {}, None, None, None, 1)


def test_gradients_equal_reference_accumulation():
    """Sum-reduced weighted NLL over a batch == the reference's per-instance accumulation (train.py:60-66)."""
    from deep_dantzig_b200.ml.models.s2v import Model
    from oracle import randomlp as orl
    torch.manual_seed(0)
    model = Model('bipartite', 8, 2, verbose_init=False)
    crit = torch.nn.NLLLoss(weight=torch.tensor([0.5, 0.5]), reduction='sum')
    insts = [orl.generate_instance(12, 5, s) for s in range(4)]
    A = torch.from_numpy(np.stack([i[0] for i in insts])); b = torch.from_numpy(np.stack([i[1] for i in insts]))
    c = torch.from_numpy(np.stack([i[2] for i in insts]))
    y = torch.randint(0, 2, (4, 12))
    model.zero_grad()
    crit(model.forward_batch(A, b, c).reshape(-1, 2), y.reshape(-1)).backward()
    g_batch = torch.cat([q.grad.reshape(-1) for q in model.parameters()])
    model.zero_grad()
    for k in range(4):
        P = {n_: q for n_, q in model.named_parameters()}
        lp, _ = oc.forward('bipartite', P, insts[k][0], insts[k][1], insts[k][2], 2)
        crit(lp, y[k]).backward()
    g_ref = torch.cat([q.grad.reshape(-1) for q in model.parameters()])
    assert torch.allclose(g_batch, g_ref, rtol=1e-4, atol=1e-5)


def _grad_cases(golden_dir, graph):
    g = np.load(os.path.join(golden_dir, 's2v_grad_%s.npz' % graph))
    for ci in range(6):
        pre = 'case%d_' % ci
        dims = [int(v) for v in g[pre + 'dims']]
        P = {k[len(pre + 'param_'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre + 'param_')}
        G = {k[len(pre + 'grad_'):]: g[k] for k in g.files if k.startswith(pre + 'grad_')}
        yield dims, P, G, g[pre + 'A'], g[pre + 'b'], g[pre + 'c'], g[pre + 'y'], float(g[pre + 'loss']), g['weight']


def grads_close(got, want, scale):
    """fp32 gradients summed over m * batch nodes: 2e-4 of the largest reference entry, absolute."""
    return np.abs(got - want).max() <= 2e-4 * scale + 1e-6


@pytest.mark.parametrize('graph', ['complete', 'bipartite'])
def test_batched_autograd_matches_reference_gradients(golden_dir, graph):
    """The batched differentiable path reproduces loss and gradients of the UNMODIFIED reference model trained the
    reference's way (tests/golden/make_s2v_grad_golden.py)."""
    from deep_dantzig_b200.ml.models.s2v import Model
    for dims, P, G, A, b, c, y, loss, w in _grad_cases(golden_dir, graph):
        model = Model(graph, dims[2], dims[3], verbose_init=False)
        model.load_state_dict(P)
        crit = torch.nn.NLLLoss(weight=torch.from_numpy(w), reduction='sum')
        lp = model.forward_batch_torch(torch.from_numpy(A), torch.from_numpy(b), torch.from_numpy(c))
        l = crit(lp.reshape(-1, 2), torch.from_numpy(y.astype(np.int64)).reshape(-1))
        l.backward()
        assert abs(float(l) - loss) <= 1e-4 * abs(loss), dims
        scale = max(np.abs(v).max() for v in G.values())
        for k, q in model.named_parameters():
            if k == 't3rc':                  # unused by the reference (quirk B10): no gradient on either side
                assert q.grad is None or float(q.grad.abs().max()) == 0.0
                continue
            assert grads_close(q.grad.numpy(), G[k], scale), (graph, dims, k)


@pytest.mark.parametrize('graph', ['complete', 'bipartite'])
def test_batched_adapter_yields_reference_items(graph):
    """``ml.utils.batched`` (reference src/ml/utils.py:3-25) turns a collated DataLoader item into per-instance items that
    ``Model.forward`` accepts; the result equals the batched forward on the same instances."""
    from deep_dantzig_b200.ml.models.s2v import Model
    from deep_dantzig_b200.ml.utils import batched
    from oracle import randomlp as orl
    torch.manual_seed(3)
    m, n, B = 12, 5, 3
    insts = [orl.generate_instance(m, n, 40 + k) for k in range(B)]
    labels = torch.randint(0, 2, (B, m))
    model = Model(graph, 6, 2, verbose_init=False)
    model.force_torch = True
    A = torch.from_numpy(np.stack([i[0] for i in insts])); b = torch.from_numpy(np.stack([i[1] for i in insts]))
    c = torch.from_numpy(np.stack([i[2] for i in insts]))
    with torch.no_grad():
        want = model.forward_batch(A, b, c)
    if graph == 'complete':
        data = {'lp': {'A': A, 'b': b, 'c': c}, 'node_features': torch.cat((torch.ones(1, m), torch.zeros(1, 1)), 1),
                'in_loss': list(range(m)), 'node_labels': torch.cat((labels, torch.zeros(B, 1, dtype=torch.long)), 1)}
        got = list(batched(data, B, 'complete'))
        assert len(got) == B
        for k, (x, y) in enumerate(got):
            with torch.no_grad():
                lp = model(x)
            assert torch.allclose(lp, want[k], atol=1e-6) and torch.equal(y, labels[k])
    else:
        # the reference's bipartite item carries one instance per DataLoader item (batch_size 1)
        for k in range(B):
            it = oc.item_bipartite(*insts[k], labels=labels[k].tolist())
            data = {'c_feats': it['c_feats'].unsqueeze(0), 'v_feats': it['v_feats'].unsqueeze(0), 'e_feats': it['e_feats'],
                    'dims': it['dims'], 'in_loss': it['in_loss'], 'c_labels': it['c_labels'].unsqueeze(0)}
            (x, y), = list(batched(data, 1, 'bipartite'))
            with torch.no_grad():
                lp = model(x)
            assert torch.allclose(lp, want[k], atol=1e-5) and torch.equal(y.long(), labels[k])
    with pytest.raises(ValueError):
        list(batched({'in_loss': []}, 1, 'hypergraph'))
