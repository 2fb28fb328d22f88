"""GPU: the batched classifier forward kernels (through the C ABI) against the reference-generated fixtures and the
CPU oracle.  Tolerance: fp32, 5e-5 absolute on log-probabilities (re-association of the relu-sum identity)."""
import os

import numpy as np
import pytest
import torch

from oracle import classifier as oc
from oracle import randomlp as orl

pytestmark = pytest.mark.gpu
TOL = 5e-5


def _close(got, want):
    """fp32: 5e-5 absolute + 1e-5 relative (random-init scores reach |log p| ~ 30 at p = 40)."""
    return bool((np.abs(got - want) <= TOL + 1e-5 * np.abs(want)).all())


def _load(golden_dir, graph):
    g = np.load(os.path.join(golden_dir, 's2v_%s.npz' % graph))
    for ci in range(6):
        pre = 'case%d_' % ci
        m, n, p, T, seed = [int(v) for v in g[pre + 'dims']]
        P = {k[len(pre + 'param_'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre + 'param_')}
        yield (m, n, p, T), P, g[pre + 'A'], g[pre + 'b'], g[pre + 'c'], g[pre + 'logp'], g[pre + 'probs']


@pytest.mark.parametrize('graph', ['bipartite', 'complete'])
def test_kernel_matches_reference_fixtures(cuda_device, golden_dir, graph):
    from deep_dantzig_b200.ml.models.s2v import Model
    for dims, P, A, b, c, logp, probs in _load(golden_dir, graph):
        model = Model(graph, dims[2], dims[3], on_cuda=True, verbose_init=False)
        model.load_state_dict(P)
        with torch.no_grad():
            lp = model.forward_batch(torch.from_numpy(A)[None].cuda(), torch.from_numpy(b)[None].cuda(), torch.from_numpy(c)[None].cuda())
        assert lp.shape == (1, dims[0], 2)
        assert _close(lp[0].cpu().numpy(), logp), (graph, dims)
        assert _close(model.probs[0].cpu().numpy(), probs), (graph, dims)


@pytest.mark.parametrize('graph,m,n,p,T', [('bipartite', 50, 20, 12, 3), ('bipartite', 200, 100, 40, 3), ('bipartite', 500, 250, 40, 3),
                                           ('complete', 50, 20, 12, 4), ('complete', 200, 100, 40, 3), ('bipartite', 37, 19, 13, 2),
                                           ('complete', 37, 19, 5, 1)])
def test_batched_kernel_vs_oracle(cuda_device, graph, m, n, p, T):
    from deep_dantzig_b200.ml.models.s2v import Model
    from deep_dantzig_b200 import solver
    torch.manual_seed(3)
    model = Model(graph, p, T, on_cuda=True, verbose_init=False)
    B = 40
    A, b, c = solver.generate(5, 0, B, m, n)
    with torch.no_grad():
        lp = model.forward_batch(A, b, c).cpu().numpy()
        lpt = model.forward_batch_torch(A, b, c).cpu().numpy()          # the autograd path computes the same thing
    assert _close(lp, lpt)
    P = {k: v.detach().cpu() for k, v in model.named_parameters()}
    An, bn, cn = A.cpu().numpy(), b.cpu().numpy(), c.cpu().numpy()
    for k in (0, 7, B - 1):
        ref, _ = oc.forward(graph, P, An[k], bn[k], cn[k], T)
        assert _close(lp[k], ref.numpy()), k
    assert np.isfinite(lp).all() and np.allclose(np.exp(lp).sum(2), 1.0, atol=1e-5)


def test_sparse_instances_use_the_adjacency(cuda_device):
    from deep_dantzig_b200.ml.models.s2v import Model
    from deep_dantzig_b200 import solver
    torch.manual_seed(4)
    model = Model('bipartite', 10, 3, on_cuda=True, verbose_init=False)
    A, b, c = solver.generate(6, 0, 16, 60, 30, density=0.3)
    with torch.no_grad():
        lp = model.forward_batch(A, b, c).cpu().numpy()
    P = {k: v.detach().cpu() for k, v in model.named_parameters()}
    for k in (0, 5, 15):
        ref, _ = oc.forward('bipartite', P, A[k].cpu().numpy(), b[k].cpu().numpy(), c[k].cpu().numpy(), 3)
        assert _close(lp[k], ref.numpy())


@pytest.mark.parametrize('graph', ['bipartite', 'complete'])
def test_forward_item_dropin(cuda_device, graph):
    """Model.forward(item) with the reference's item dictionaries and an in_loss subset."""
    from deep_dantzig_b200.ml.models.s2v import Model
    torch.manual_seed(5)
    model = Model(graph, 12, 3, on_cuda=True, verbose_init=False)
    A, b, c = orl.generate_instance(50, 20, 11)
    item = oc.item_complete(A, b, c) if graph == 'complete' else oc.item_bipartite(A, b, c)
    item['in_loss'] = [0, 3, 7, 49]
    with torch.no_grad():
        lp = model(item)
    P = {k: v.detach().cpu() for k, v in model.named_parameters()}
    ref, refp = oc.forward(graph, P, A, b, c, 3, in_loss=[0, 3, 7, 49])
    assert lp.shape == (4, 2) and model.probs.shape == (4, 2)
    assert _close(lp.cpu().numpy(), ref.numpy())
    assert _close(model.probs.cpu().numpy(), refp.numpy())


def test_drivers_end_to_end(cuda_device, tmp_path):
    """benchmark.run_benchmark / phase_transitions.compute_phaseTransitions / sweep_ratio_density on tiny settings:
    dataset generated+solved+labelled on the GPU, classifier trained with the batched forward, records saved with the
    reference's file naming and schema."""
    import json
    from deep_dantzig_b200.benchmark import run_benchmark
    from deep_dantzig_b200.phase_transitions import compute_phaseTransitions, sweep_ratio_density
    grid = {'dataset': ['randomlp'], 'graph': ['bipartite'], 'elem_type': ['lp'], 'num_elems': [48], 'p': [8],
            'rounds_s2v': [2], 'epochs': [3], 'batch_size': [16], 'learning_rate': [0.01], 'momentum': [0.9],
            'weight_decay': [0], 'seed': [0], 'm': [20], 'n': [10]}
    stamps = run_benchmark(grid, str(tmp_path), cuda=True, tag='t')
    res = json.load(open(tmp_path / ('benchmark_randomlp_res_%s.json' % stamps[0])))
    assert set(res) == {'params', 'out', 'dataset', 'seed', 'cuda', 'tag'} and len(res['out']['results']['train']) == 3
    assert (tmp_path / ('benchmark_randomlp_model_%s.json' % stamps[0])).exists()
    assert res['out']['results']['train'][-1]['total_loss'] < res['out']['results']['train'][0]['total_loss'] * 1.5
    bp = {'epochs': [2], 'seeds': [0], 'batch_sizes': [16], 'rounds_s2v': [1], 'learning_rates': [0.01],
          'momentums': [0.9], 'weight_decays': [0], 'ps': [3]}
    recs = compute_phaseTransitions(str(tmp_path), 'randomlp', bp, cuda=True, tag='t', num_elems=32, m=20, n=10)
    assert len(recs) == 1 and set(recs[0]['out']) == {'accs', 'losses'} and 3 in recs[0]['out']['accs']
    sw = sweep_ratio_density(n=20, ratios=(1.5, 3.0), densities=(1.0, 0.3), per_cell=300, chunk=128)
    assert all(v['instances'] == 300 and v['optimal'] + v['unbounded'] + v['other'] == 300 for v in sw.values())
    assert sw[(3.0, 1.0)]['optimal'] > sw[(1.5, 1.0)]['optimal']        # more rows -> fewer unbounded instances (Wendel)


def _grad_cases(golden_dir, graph):
    g = np.load(os.path.join(golden_dir, 's2v_grad_%s.npz' % graph))
    for ci in range(6):
        pre = 'case%d_' % ci
        dims = [int(v) for v in g[pre + 'dims']]
        P = {k[len(pre + 'param_'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre + 'param_')}
        G = {k[len(pre + 'grad_'):]: g[k] for k in g.files if k.startswith(pre + 'grad_')}
        yield dims, P, G, g[pre + 'A'], g[pre + 'b'], g[pre + 'c'], g[pre + 'y'], float(g[pre + 'loss']), g['weight']


@pytest.mark.parametrize('graph', ['bipartite', 'complete'])
def test_backward_kernel_matches_reference_gradients(cuda_device, golden_dir, graph):
    """ddb_s2v_loss_grad_dev (both graph variants) against loss + gradients of the UNMODIFIED reference model accumulated
    the reference's way (tests/golden/make_s2v_grad_golden.py).  fp32: 2e-4 of the largest gradient entry."""
    from deep_dantzig_b200.ml.models.s2v import Model
    for dims, P, G, A, b, c, y, loss, w in _grad_cases(golden_dir, graph):
        model = Model(graph, dims[2], dims[3], on_cuda=True, verbose_init=False)
        model.load_state_dict(P)
        model.zero_grad()
        l = model.loss_and_grad_batch(torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(c).cuda(),
                                      torch.from_numpy(y).cuda(), [float(w[0]), float(w[1])])
        assert model.last_batch_was_dense()
        assert abs(float(l) - loss) <= 1e-4 * abs(loss), dims
        scale = max(np.abs(v).max() for v in G.values())
        for k, q in model.named_parameters():
            assert np.abs(q.grad.cpu().numpy() - G[k]).max() <= 2e-4 * scale + 1e-6, (dims, k)


@pytest.mark.parametrize('graph', ['bipartite', 'complete'])
@pytest.mark.parametrize('m,n,p,T,B', [(200, 100, 40, 3, 600), (50, 20, 12, 4, 1000), (37, 19, 13, 1, 77), (120, 60, 48, 2, 300),
                                       (30, 12, 64, 3, 200), (23, 9, 5, 0, 50)])
def test_backward_kernel_vs_autograd_large_batch(cuda_device, graph, m, n, p, T, B):
    """Same loss and gradient as autograd through the batched torch restatement, on solver-produced labels."""
    from deep_dantzig_b200.ml.models.s2v import Model
    from deep_dantzig_b200 import solver
    torch.manual_seed(11)
    model = Model(graph, p, T, on_cuda=True, verbose_init=False)
    A, b, c = solver.generate(9, 0, B, m, n)
    y = solver.solve_label(A, b, c)['labels']
    w = [0.3, 0.7]
    model.zero_grad()
    l_dev = model.loss_and_grad_batch(A, b, c, y, w)
    assert model.last_batch_was_dense()
    g_dev = torch.cat([q.grad.reshape(-1) for q in model.parameters()]).clone()
    model.zero_grad()
    crit = torch.nn.NLLLoss(weight=torch.tensor(w, device='cuda'), reduction='sum')
    l_ref = crit(model.forward_batch_torch(A, b, c).reshape(-1, 2), y.long().reshape(-1))
    l_ref.backward()
    # a parameter the forward never uses (t3rc of the complete graph, quirk B10) has no autograd gradient: zeros
    g_ref = torch.cat([(q.grad if q.grad is not None else torch.zeros_like(q)).reshape(-1) for q in model.parameters()])
    assert abs(float(l_dev) - float(l_ref)) <= 2e-4 * abs(float(l_ref))
    assert float((g_dev - g_ref).abs().max()) <= 5e-4 * float(g_ref.abs().max()) + 1e-5


@pytest.mark.parametrize('m,n,p,T,B,density', [(40, 20, 8, 2, 64, 0.5), (200, 100, 40, 3, 160, 0.1), (60, 30, 12, 3, 300, 0.3),
                                               (37, 19, 13, 1, 50, 0.5), (50, 20, 64, 4, 40, 0.5), (23, 9, 5, 0, 20, 0.5)])
def test_backward_general_adjacency_vs_autograd(cuda_device, m, n, p, T, B, density):
    """Instances with zero coefficients (general adjacency; at density 0.1 also empty rows and columns) mixed with dense
    ones: the streaming kernel flags them, the general-adjacency kernel adds them; loss and gradient equal autograd through
    the batched torch restatement."""
    from deep_dantzig_b200.ml.models.s2v import Model
    from deep_dantzig_b200 import solver
    torch.manual_seed(5)
    model = Model('bipartite', p, T, on_cuda=True, verbose_init=False)
    A, b, c = solver.generate(6, 0, B, m, n, density=density)
    Ad, bd, cd = solver.generate(7, 0, B // 2, m, n)                       # dense instances in the same batch
    A, b, c = torch.cat((A, Ad)), torch.cat((b, bd)), torch.cat((c, cd))
    y = solver.solve_label(A, b, c)['labels']
    y[::3] = (torch.rand_like(y[::3].float()) < 0.3).to(y.dtype)           # labels on sparse / unbounded instances too
    w = [0.3, 0.7]
    model.zero_grad()
    l_dev = model.loss_and_grad_batch(A, b, c, y, w)
    assert not model.last_batch_was_dense()
    g_dev = torch.cat([q.grad.reshape(-1) for q in model.parameters()]).clone()
    model.zero_grad()
    crit = torch.nn.NLLLoss(weight=torch.tensor(w, device='cuda'), reduction='sum')
    l_ref = crit(model.forward_batch_torch(A, b, c).reshape(-1, 2), y.long().reshape(-1))
    l_ref.backward()
    g_ref = torch.cat([(q.grad if q.grad is not None else torch.zeros_like(q)).reshape(-1) for q in model.parameters()])
    assert abs(float(l_dev) - float(l_ref.detach())) <= 2e-4 * abs(float(l_ref.detach()))
    assert float((g_dev - g_ref).abs().max()) <= 5e-4 * float(g_ref.abs().max()) + 1e-5


def test_device_metrics_match_sklearn_and_torch(cuda_device):
    """ddb_s2v_metrics_dev: recall-1 threshold == sklearn's ROC rule (train.py:138-140), confusion counts and weighted
    NLL == the torch formulas of performance() (train.py:174-246)."""
    from sklearn.metrics import roc_curve
    from torch.utils.data import DataLoader, TensorDataset
    from deep_dantzig_b200.ml.models.s2v import Model
    from deep_dantzig_b200.ml import train as tr
    from deep_dantzig_b200 import solver
    torch.manual_seed(2)
    model = Model('bipartite', 12, 3, on_cuda=True, verbose_init=False)
    A, b, c = solver.generate(3, 0, 300, 50, 20)
    y = solver.solve_label(A, b, c)['labels'].long()
    loader = [{'A': A[k:k + 64], 'b': b[k:k + 64], 'c': c[k:k + 64], 'y': y[k:k + 64]} for k in range(0, 300, 64)]
    crit = torch.nn.NLLLoss(weight=torch.tensor([0.6, 0.4], device='cuda'), reduction='sum')
    thr = tr.recall_one_threshold(loader, model)
    with torch.no_grad():
        logp = model.forward_batch(A, b, c)
    probs = model.probs[..., 1].reshape(-1).cpu().numpy()
    yt = y.reshape(-1).cpu().numpy()
    fpr, tpr, ths = roc_curve(yt, probs, pos_label=1)
    assert thr == float(ths[np.where(tpr == 1.0)[0][0]]) == float(probs[yt == 1].min())
    got = tr.performance(loader, model, crit, thr)
    pred = probs >= np.float32(thr)
    tp, fp = int((pred & (yt == 1)).sum()), int((pred & (yt == 0)).sum())
    tn, fn = int((~pred & (yt == 0)).sum()), int((~pred & (yt == 1)).sum())
    assert got['recall'] == 1.0 and fn == 0
    assert abs(got['accuracy'] - (tp + tn) / yt.size) < 1e-12 and abs(got['precision'] - tp / (tp + fp)) < 1e-12
    assert abs(got['pred_pos'] - (tp + fp) / yt.size) < 1e-12 and abs(got['y_pos'] - (tp + fn) / yt.size) < 1e-12
    want_loss = float(crit(logp.reshape(-1, 2), y.reshape(-1)))
    assert abs(got['total_loss'] - want_loss) <= 1e-5 * abs(want_loss)


def test_streaming_training_and_reduced_solve(cuda_device):
    """Config 5 on one rank (on-GPU generation feeds the loss+gradient kernel, loss goes down) followed by config 4
    (classifier prunes rows at the recall-1 threshold, reduced LP solved and certified; results equal the full solve)."""
    from deep_dantzig_b200.ml.models.s2v import Model
    from deep_dantzig_b200.ml import train as tr
    from deep_dantzig_b200 import reduced, solver
    torch.manual_seed(0)
    model = Model('bipartite', 12, 2, on_cuda=True, verbose_init=False)
    opt = torch.optim.SGD(model.parameters(), lr=2e-5, momentum=0.9)
    m, n = 60, 20
    first = None
    for overlap in (True, False):
        hist = tr.train_on_device_stream(model, opt, m, n, steps=60, batch_per_rank=256, key=77, weight=(0.35, 0.65), overlap=overlap)
        assert hist['lps'] == 60 * 256 and np.isfinite(hist['loss']).all()
        first = hist['loss'][:5].mean() if first is None else first
    print('streaming training: loss per node %.4f -> %.4f' % (first, hist['loss'][-10:].mean()))
    assert hist['loss'][-10:].mean() < first
    A, b, c = solver.generate(78, 0, 2000, m, n)
    y = solver.solve_label(A, b, c)['labels'].long()
    loader = [{'A': A, 'b': b, 'c': c, 'y': y}]
    thr = tr.recall_one_threshold(loader, model)
    t = reduced.timing_forward_pass(model, A, b, c, thr)
    assert t['status_match'] == 2000 and t['label_match'] == 2000 and t['max_rel_x_diff'] <= 1e-9
    assert t['certified_frac_of_optimal'] == 1.0                  # recall-1 threshold on these very instances
    assert t['rows_kept_frac'] < 1.0
    # a threshold that drops active rows is caught by the certificate and repaired by the full re-solve
    t2 = reduced.timing_forward_pass(model, A, b, c, min(0.999, thr + 0.25))
    assert t2['status_match'] == 2000 and t2['label_match'] == 2000 and t2['certified_frac_of_optimal'] < 1.0


@pytest.mark.parametrize('m,n,p,T', [(200, 100, 40, 3), (50, 20, 12, 4), (64, 32, 13, 1), (120, 128, 64, 2), (40, 20, 7, 0)])
def test_dense_streaming_kernel_and_flagged_instances(cuda_device, m, n, p, T):
    """The HBM-streaming dense kernel (csrc/s2v_bipartite_dense.cu) against the oracle; instances with a zero coefficient
    inside an otherwise dense batch are flagged on the device and take the general-adjacency kernel."""
    from deep_dantzig_b200.ml.models.s2v import Model
    from deep_dantzig_b200 import solver
    torch.manual_seed(21)
    model = Model('bipartite', p, T, on_cuda=True, verbose_init=False)
    B = 700
    A, b, c = solver.generate(13, 0, B, m, n)
    sparse_ids = [3, 150, 151, B - 1]
    for k in sparse_ids:
        A[k, (k * 7) % m, (k * 3) % n] = 0.0
        A[k, 0, :n // 2] = 0.0
    with torch.no_grad():
        lp = model.forward_batch(A, b, c).cpu().numpy()
        pr = model.probs.cpu().numpy()
    P = {k: v.detach().cpu() for k, v in model.named_parameters()}
    An, bn, cn = A.cpu().numpy(), b.cpu().numpy(), c.cpu().numpy()
    for k in sparse_ids + [0, 1, 2, 149, 152, 295, 296, 400, B - 2]:
        ref, refp = oc.forward('bipartite', P, An[k], bn[k], cn[k], T)
        assert _close(lp[k], ref.numpy()), k
        assert _close(pr[k], refp.numpy()), k
    assert np.isfinite(lp).all() and np.allclose(np.exp(lp).sum(2), 1.0, atol=1e-5)
    # the general kernel alone computes the same thing on the whole batch
    os.environ['DDB_S2V_NO_DENSE_CHECK'] = '1'
    with torch.no_grad():
        lpt = model.forward_batch_torch(A, b, c).cpu().numpy()
    assert _close(lp, lpt)


@pytest.mark.parametrize('graph', ['bipartite', 'complete'])
def test_training_trajectory_matches_the_reference_loop(cuda_device, golden_dir, graph):
    """North-star bar "unchanged classifier accuracy": the mirror, trained through ``train_net`` on the GPU from the same
    initial parameters on the same BASELINE.json configs[0] data in the same order with the reference's hyper-parameters
    (src/run.py:58-72), lands on the trajectory of the UNMODIFIED reference model under the reference's own loop
    (src/ml/train.py:49-71, criterion src/benchmark.py:70-77) -- tests/golden/train_trajectory_*.npz, made by
    tests/golden/make_train_golden.py from /root/reference/src.  Per epoch: losses within 1e-3 relative, recall-1 threshold
    within 1e-4, confusion counts identical up to the sample that sits exactly on the threshold."""
    from deep_dantzig_b200.ml.models.s2v import Model
    from deep_dantzig_b200.ml import train as tr
    from oracle import randomlp as orl
    g = np.load(os.path.join(golden_dir, 'train_trajectory_%s.npz' % graph))
    m, n, p, T, epochs, batch, ntrain, ntest = [int(v) for v in g['dims']]
    lr, momentum, wd = [float(v) for v in g['hyper']]
    cfg1 = np.load(os.path.join(golden_dir, 'randomlp_config1.npz'))
    assert list(cfg1['seeds'][:ntrain + ntest]) == list(g['seeds'])
    labels = np.unpackbits(cfg1['labels_packed'], axis=1)[:ntrain + ntest, :m].astype(np.int64)
    inst = [orl.generate_instance(m, n, int(s)) for s in g['seeds']]

    def loader(lo, hi):
        out = []
        for q in range(lo, hi, batch):
            e = min(q + batch, hi)
            out.append({'A': torch.from_numpy(np.stack([inst[i][0] for i in range(q, e)])),
                        'b': torch.from_numpy(np.stack([inst[i][1] for i in range(q, e)])),
                        'c': torch.from_numpy(np.stack([inst[i][2] for i in range(q, e)])),
                        'y': torch.from_numpy(labels[q:e])})
        return out
    model = Model(graph, p, T, on_cuda=True, verbose_init=False)
    with torch.no_grad():
        for k, v in model.named_parameters():
            v.copy_(torch.from_numpy(g['init_' + k]).cuda())
    crit = torch.nn.NLLLoss(weight=torch.from_numpy(g['weight']).cuda(), reduction='sum')
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=momentum, weight_decay=wd)
    hist = tr.train_net(model, crit, opt, loader(0, ntrain), loader(ntrain, ntrain + ntest), epochs, batch, cuda=True, verbose=False)
    ref = g['history']            # [epoch][running, thr, train(loss,tp,fp,tn,fn), test(loss,tp,fp,tn,fn)]
    assert set(hist) == {'train', 'test'}
    thr = tr.train_net.last_thresholds
    for ep in range(epochs):
        for split, off in (('train', 2), ('test', 7)):
            mt = hist[split][ep]
            want_loss, tp, fp, tn, fn = ref[ep][off:off + 5]
            tot = tp + fp + tn + fn
            assert abs(mt['total_loss'] - want_loss) <= 1e-3 * abs(want_loss), (graph, ep, split, mt['total_loss'], want_loss)
            # accuracy / recall / pred_pos are ratios of the confusion counts: one sample may sit exactly on the threshold
            assert abs(mt['accuracy'] * tot - (tp + tn)) <= 1.5, (graph, ep, split)
            assert abs(mt['recall'] * (tp + fn) - tp) <= 1.5, (graph, ep, split)
            assert abs(mt['pred_pos'] * tot - (tp + fp)) <= 1.5, (graph, ep, split)
        if thr is not None:
            assert abs(thr[ep] - ref[ep][1]) <= 1e-4, (graph, ep, thr[ep], ref[ep][1])
    # and the parameters end where the reference's ended
    for k, v in model.named_parameters():
        want = g['final_' + k]
        assert np.abs(v.detach().cpu().numpy() - want).max() <= 2e-3 * max(1.0, np.abs(want).max()), k


def test_flagged_items_run_on_the_cuda_kernels(cuda_device, golden_dir):
    """MPS / PLNN items -- equality rows, bound rows, 0 / 1 node features -- through Model.forward on the GPU: the CUDA kernels
    with the item's row flags (ddb_s2v_forward_flags_dev) against the outputs of the UNMODIFIED reference model
    (tests/golden/s2v_plnn_items.npz), and against the torch restatement on a larger random sparse batch."""
    from deep_dantzig_b200.ml.models.s2v import Model
    from deep_dantzig_b200 import _lib
    g = np.load(os.path.join(golden_dir, 's2v_plnn_items.npz'))
    ctx = _lib.context(0)
    for ci in range(int(g['ncases'])):
        pre = 'bip%d_' % ci
        m, n, p, T = [int(v) for v in g[pre + 'dims']]
        A = g[pre + 'A']
        idx = [[i, j] for i in range(m) for j in range(n) if A[i, j] != 0]
        item = {'c_feats': torch.from_numpy(g[pre + 'c_feats'].copy()), 'v_feats': torch.from_numpy(g[pre + 'v_feats']),
                'e_feats': {'i': idx, 'coeffs': [float(A[i, j]) for i, j in idx]}, 'in_loss': [int(q) for q in g[pre + 'in_loss']],
                'dims': {'m': m, 'n': n}}
        model = Model('bipartite', p, T, on_cuda=True, verbose_init=False)
        model.load_state_dict({k[len(pre) + 6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre + 'param_')})
        model.cuda()
        n0 = ctx.launch_count()
        with torch.no_grad():
            got = model.forward(item).cpu().numpy()
        assert ctx.launch_count() > n0, 'the item did not go through the C ABI'
        assert got.shape == g[pre + 'logp'].shape and np.abs(got - g[pre + 'logp']).max() <= 5e-5
        pre = 'cmp%d_' % ci
        item = {'A': torch.from_numpy(g[pre + 'A']).unsqueeze(0), 'b': torch.from_numpy(g[pre + 'b']).unsqueeze(0),
                'c': torch.from_numpy(g[pre + 'c']).unsqueeze(0), 'node_features': torch.from_numpy(g[pre + 'node_features']).unsqueeze(0),
                'in_loss': [int(q) for q in g[pre + 'in_loss']]}
        model = Model('complete', p, T, on_cuda=True, verbose_init=False)
        model.load_state_dict({k[len(pre) + 6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre + 'param_')})
        model.cuda()
        n0 = ctx.launch_count()
        with torch.no_grad():
            got = model.forward(item).cpu().numpy()
        assert ctx.launch_count() > n0
        assert got.shape == g[pre + 'logp'].shape and np.abs(got - g[pre + 'logp']).max() <= 5e-5
    # batched, larger shapes: random sparse instances with random row flags, both graphs, kernels against the restatement
    from deep_dantzig_b200 import solver
    gen = torch.Generator(device='cuda').manual_seed(5)
    for graph, (B, m, n, p, T) in (('bipartite', (48, 60, 24, 16, 3)), ('complete', (48, 60, 24, 16, 2)), ('bipartite', (6, 200, 100, 40, 3)),
                                   ('complete', (6, 200, 100, 40, 3))):
        A, b, c = solver.generate(31, 0, B, m, n, density=0.3)
        ineq = (torch.rand(B, m, device='cuda', generator=gen) < 0.7).float()
        bound = (torch.rand(B, m, device='cuda', generator=gen) < 0.2).float()
        feats = (ineq, bound) if graph == 'bipartite' else ineq
        torch.manual_seed(11)
        model = Model(graph, p, T, on_cuda=True, verbose_init=False)
        with torch.no_grad():
            want = model.forward_batch_torch(A, b, c, feats)
            wprobs = model.probs
            got = model.forward_batch_cuda(A, b, c, feats)
        assert (got - want).abs().max().item() <= 5e-5 + 1e-5 * want.abs().max().item(), (graph, m, n)
        assert (model.probs - wprobs).abs().max().item() <= 5e-5


@pytest.mark.parametrize('graph,B,m,n,p,T', [('bipartite', 5, 30, 12, 8, 2), ('bipartite', 3, 90, 40, 24, 3), ('complete', 5, 30, 12, 8, 2),
                                             ('complete', 3, 90, 40, 24, 3), ('bipartite', 2, 200, 100, 40, 3), ('complete', 2, 200, 100, 40, 3)])
def test_loss_grad_with_row_flags_matches_autograd(cuda_device, graph, B, m, n, p, T):
    """ddb_s2v_loss_grad_flags_dev (MPS / PLNN items: equality / bound rows, rows outside in_loss marked by label 2) against
    autograd through the batched restatement with the same flags: loss and every parameter gradient."""
    from deep_dantzig_b200 import solver
    from deep_dantzig_b200.ml.models.s2v import Model
    gen = torch.Generator(device='cuda').manual_seed(3)
    A, b, c = solver.generate(77, 0, B, m, n, density=0.35)
    ineq = (torch.rand(B, m, device='cuda', generator=gen) < 0.7).float()
    bound = (torch.rand(B, m, device='cuda', generator=gen) < 0.2).float()
    feats = (ineq, bound) if graph == 'bipartite' else ineq
    y = (torch.rand(B, m, device='cuda', generator=gen) < 0.4).to(torch.uint8)
    out = torch.rand(B, m, device='cuda', generator=gen) < 0.25          # rows outside in_loss
    y2 = torch.where(out, torch.full_like(y, 2), y)
    w = [0.3, 0.7]
    torch.manual_seed(4)
    model = Model(graph, p, T, on_cuda=True, verbose_init=False)
    model.zero_grad()
    l_dev = model.loss_and_grad_batch(A, b, c, y2, w, feats)
    g_dev = {k: q.grad.clone() for k, q in model.named_parameters()}
    model.zero_grad()
    logp = model.forward_batch_torch(A, b, c, feats)
    crit = torch.nn.NLLLoss(weight=torch.tensor(w, device='cuda'), reduction='sum')
    keep = ~out
    l_ref = crit(logp[keep], y[keep].long())
    l_ref.backward()
    assert abs(float(l_dev) - float(l_ref.detach())) <= 2e-4 * abs(float(l_ref.detach())) + 1e-5
    ref = {k: (q.grad if q.grad is not None else torch.zeros_like(q)) for k, q in model.named_parameters()}   # unused parameters: no grad
    gmax = max(float(v.abs().max()) for v in ref.values())
    for k in ref:
        assert float((g_dev[k] - ref[k]).abs().max()) <= 2e-4 * gmax + 1e-6, k
