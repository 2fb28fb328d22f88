"""Drop-in for the reference's ``ml.models.s2v.Model`` (src/ml/models/s2v.py:9-323) with a batched B200 forward.

Same constructor ``Model(graph, p, rounds_s2v, on_cuda=False)``, same parameter names/shapes/initial scales
(s2v.py:60-89, 189-216 -- so reference ``state_dict``s load unchanged), same ``forward(item)`` contract
(log-probabilities ``(len(in_loss), 2)``, side effect ``self.probs``), same ``require_grads``.

Two execution paths:
  * inference (no autograd): ``ddb_s2v_forward_dev`` -- the hand-written CUDA kernels of csrc/s2v_forward.cu, one call
    for a whole batch ``forward_batch(A, b, c)``; this is the hot path;
  * training (autograd needed): ``forward_batch_torch`` -- a batched torch restatement of the same arithmetic (library
    kernels) so that ``loss.backward()`` works until the hand-written backward kernel lands.
Quirks B9/B10 of the reference are reproduced in both; B11 (``np.concatenate`` on a CUDA tensor) is simply fixed.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _lib

GRAPH_CODE = {'complete': 0, 'bipartite': 1}
COMPLETE_PARAMS = ['t0', 't1', 't2rr', 't2rc', 't2cr', 't3rr', 't3rc', 't3cr', 't4rr', 't4rc', 't4cr', 't6r', 't6c', 't7', 't8']
BIPARTITE_PARAMS = ['t0', 't1c', 't1v', 't2c', 't2v', 't3c', 't3v', 't4c', 't4v', 't6c', 't6v', 't7', 't8']


class Model(nn.Module):

    def __init__(self, graph, p, rounds_s2v, on_cuda=False, verbose_init=True):
        super(Model, self).__init__()
        self.verbose = False
        self.on_cuda = on_cuda
        self.graph, self.p, self.T = graph, p, rounds_s2v
        if verbose_init:
            print('rounds_s2v: %d ' % (self.T))                       # s2v.py:38
        dev = torch.device('cuda') if on_cuda else torch.device('cpu')
        C_, K = math.sqrt(1 / p), (lambda x: math.sqrt(1 / x))

        def P(scale, *shape):
            return nn.Parameter((scale * torch.randn(*shape)).float().to(dev), requires_grad=True)

        if graph == 'complete':                                        # s2v.py:60-89
            self.t0, self.t1 = P(1.0, p, 1), P(1.0, p, 1)
            self.t2rr, self.t2rc, self.t2cr = P(C_, p, p), P(C_, p, p), P(C_, p, p)
            self.t3rr, self.t3rc, self.t3cr = P(C_, p, p), P(C_, p, p), P(C_, p, p)
            self.t4rr, self.t4rc, self.t4cr = P(C_, p, 1), P(C_, p), P(C_, p)
            self.t6r, self.t6c, self.t7 = P(C_, p, p), P(C_, p, p), P(C_, p, p)
            self.t8 = P(C_, 2, 2 * p)
            self._names = COMPLETE_PARAMS
        elif graph == 'bipartite':                                     # s2v.py:189-216
            self.t0, self.t1c, self.t1v = P(1.0, p, 1), P(K(4), p, 4), P(1.0, p, 1)
            self.t2c, self.t2v = P(C_, p, p), P(C_, p, p)
            self.t3c, self.t3v = P(C_, 1, p, p), P(C_, 1, p, p)
            self.t4c, self.t4v = P(C_, 1, p, 1), P(C_, 1, p, 1)
            self.t6c, self.t6v, self.t7 = P(C_, p, p), P(C_, p, p), P(C_, p, p)
            self.t8 = P(K(2 * p + 4), 2, 2 * p + 4)
            self._names = BIPARTITE_PARAMS
        else:
            raise ValueError('Graph not recognised')
        self.probs = None
        # explicit opt-in (tests of the host-side training logic on CPU/gloo); never set implicitly: inference
        # without it goes to the CUDA kernel or raises
        self.force_torch = False

    def require_grads(self, req=True):
        for param in self.parameters():
            param.requires_grad = req

    def flat_params(self):
        """fp32 parameter block in the order/shape the C ABI documents (include/ddb200.h (4))."""
        return torch.cat([getattr(self, k).detach().reshape(-1) for k in self._names]).contiguous()

    # ------------------------------------------------------------------------------------------------------------
    # batched entry points
    # ------------------------------------------------------------------------------------------------------------
    def forward_batch(self, A, b, c):
        """A[B,m,n], b[B,m], c[B,n] float64 -> log-probs [B,m,2]; sets self.probs [B,m,2].  CUDA kernel unless
        autograd is recording for a parameter, in which case the differentiable torch path is used."""
        if self.force_torch or (torch.is_grad_enabled() and any(q.requires_grad for q in self.parameters())):
            return self.forward_batch_torch(A, b, c)
        return self.forward_batch_cuda(A, b, c)

    def forward_batch_cuda(self, A, b, c, feats=None):
        """CUDA forward.  ``feats`` = node flags of MPS / PLNN items (see ``forward_batch_torch``): they go to
        ``ddb_s2v_forward_flags_dev`` as uint8 [B,m] arrays."""
        if not (A.is_cuda and b.is_cuda and c.is_cuda):
            raise _lib.DdbError('the classifier forward kernel needs CUDA tensors; there is no CPU fallback')
        B, m, n = A.shape
        A, b, c = A.double().contiguous(), b.double().contiguous(), c.double().contiguous()
        dev = A.device
        params = self.flat_params().to(dev)
        ctx = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
        assert params.numel() == ctx.lib.ddb_s2v_param_count(GRAPH_CODE[self.graph], self.p)
        logp = torch.empty(B, m, 2, dtype=torch.float32, device=dev)
        probs = torch.empty(B, m, 2, dtype=torch.float32, device=dev)
        vp = lambda t: C.c_void_p(t.data_ptr())
        if feats is None:
            rc = ctx.lib.ddb_s2v_forward_dev(ctx.handle, GRAPH_CODE[self.graph], B, m, n, self.p, self.T, vp(A), vp(b), vp(c),
                                             vp(params), vp(logp), vp(probs),
                                             C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        else:
            fl = (feats, None) if torch.is_tensor(feats) else feats
            ineq = fl[0].to(device=dev).reshape(B, m).ne(0).to(torch.uint8).contiguous()
            bound = None if fl[1] is None else fl[1].to(device=dev).reshape(B, m).ne(0).to(torch.uint8).contiguous()
            rc = ctx.lib.ddb_s2v_forward_flags_dev(ctx.handle, GRAPH_CODE[self.graph], B, m, n, self.p, self.T, vp(A), vp(b), vp(c),
                                                   vp(params), vp(ineq), vp(bound) if bound is not None else None,
                                                   vp(logp), vp(probs), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, 'ddb_s2v_forward_dev')
        self.probs = probs
        return logp

    def device_backward_supported(self, A):
        """The hand-written backward covers the bipartite graph -- dense instances (csrc/s2v_backward.cu, the reference's
        default model) and instances with zero coefficients (csrc/s2v_bipartite_general_backward.cu) -- and the complete
        graph (csrc/s2v_complete_backward.cu, m + 1 <= 256), all with p <= 64; only what the kernels do not fit
        (``DdbError`` "do not fit") trains through ``forward_batch_torch`` + autograd."""
        return self.graph in GRAPH_CODE and self.p <= 64 and A.is_cuda and not self.force_torch

    def loss_and_grad_batch(self, A, b, c, labels, weight, feats=None):
        """One training step's loss and gradient on the device: ``ddb_s2v_loss_grad_dev``.  Replaces the reference's
        per-instance ``loss = criterion(model(x), y); loss.backward()`` accumulation (train.py:59-65) with
        ``criterion = NLLLoss(weight, size_average=False)`` (benchmark.py:70-75).  labels [B,m] (0/1), weight = [w0, w1].
        ACCUMULATES into ``param.grad`` like ``backward()`` does and returns the summed loss (0-dim fp64 tensor on the device).
        ``feats`` = node flags of MPS / PLNN items (see ``forward_batch_torch``; ``ddb_s2v_loss_grad_flags_dev``); a label of 2
        marks a row outside the item's ``in_loss`` set."""
        if not (A.is_cuda and b.is_cuda and c.is_cuda):
            raise _lib.DdbError('the classifier backward kernel needs CUDA tensors; there is no CPU fallback')
        B, m, n = A.shape
        dev = A.device
        A, b, c = A.double().contiguous(), b.double().contiguous(), c.double().contiguous()
        y = labels.to(device=dev, dtype=torch.uint8).contiguous()
        params = self.flat_params().to(dev)
        ctx = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
        grad = torch.empty(params.numel(), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float64, device=dev)
        flag = torch.empty(1, dtype=torch.int32, device=dev)
        vp = lambda t: C.c_void_p(t.data_ptr())
        if feats is None:
            rc = ctx.lib.ddb_s2v_loss_grad_dev(ctx.handle, GRAPH_CODE[self.graph], B, m, n, self.p, self.T, vp(A), vp(b), vp(c),
                                               vp(params), vp(y), float(weight[0]), float(weight[1]), vp(grad), vp(loss), vp(flag),
                                               C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        else:
            fl = (feats, None) if torch.is_tensor(feats) else feats
            ineq = fl[0].to(device=dev).reshape(B, m).ne(0).to(torch.uint8).contiguous()
            bound = None if fl[1] is None else fl[1].to(device=dev).reshape(B, m).ne(0).to(torch.uint8).contiguous()
            rc = ctx.lib.ddb_s2v_loss_grad_flags_dev(ctx.handle, GRAPH_CODE[self.graph], B, m, n, self.p, self.T, vp(A), vp(b), vp(c),
                                                     vp(params), vp(y), vp(ineq), vp(bound) if bound is not None else None,
                                                     float(weight[0]), float(weight[1]), vp(grad), vp(loss), vp(flag),
                                                     C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, 'ddb_s2v_loss_grad_dev')
        self._last_grad_flag = flag
        off = 0
        for k in self._names:
            q = getattr(self, k)
            g = grad[off:off + q.numel()].view_as(q)
            off += q.numel()
            if q.requires_grad:
                q.grad = g.clone() if q.grad is None else q.grad.add_(g)
        return loss

    def last_batch_was_dense(self):
        """False if the last ``loss_and_grad_batch`` met an instance with a zero coefficient -- informational: such
        instances went through the general-adjacency kernel, the result is complete either way; reading it synchronises
        with the device."""
        return int(self._last_grad_flag.item()) == 0

    def forward_batch_torch(self, A, b, c, feats=None):
        """Differentiable batched restatement (same arithmetic, relu-sum identity, quirks B9/B10).  ``feats`` carries the
        node flags of items that are not plain random LPs (MPS / PLNN items): bipartite ->
        (is_inequality [B,m], is_bound [B,m]); complete -> node_features [B,m] (1 for an inequality row, 0 for an equality).
        Without it every row is a non-bound inequality, which is what the CUDA kernels assume."""
        if self.graph == 'complete':
            scores = self._scores_complete(A, b, c, feats)
        else:
            scores = self._scores_bipartite(A, b, c, feats)
        self.probs = F.softmax(scores, dim=2)
        return F.log_softmax(scores, dim=2)

    def _scores_bipartite(self, A, b, c, feats=None):
        B, m, n = A.shape
        p = self.p
        A32, rhs, cv = A.float(), b.float(), c.float()
        Ab = F.normalize(torch.cat((A32, -rhs.unsqueeze(2)), 2), p=2, dim=2)
        An = Ab[:, :, :n]
        adj = (A32 != 0).float()
        is_ineq = torch.ones_like(rhs) if feats is None else feats[0].to(rhs)
        is_bound = torch.zeros_like(rhs) if feats is None else feats[1].to(rhs)
        cfe = torch.stack((is_ineq, -Ab[:, :, n], is_bound, torch.bmm(An, cv.unsqueeze(2)).squeeze(2)), 2)  # [B,m,4]
        Sp, Sn = F.relu(An).sum(2), F.relu(-An).sum(2)                     # [B,m]
        Cp, Cn = F.relu(An).sum(1), F.relu(-An).sum(1)                     # [B,n]
        t4c, t4v = self.t4c[0, :, 0], self.t4v[0, :, 0]
        w3cp, w3cn = self.t3c[0] @ F.relu(t4c), self.t3c[0] @ F.relu(-t4c)
        w3vp, w3vn = self.t3v[0] @ F.relu(t4v), self.t3v[0] @ F.relu(-t4v)
        base_c = self.t0.unsqueeze(0) + torch.einsum('kf,bif->bki', self.t1c, cfe) \
            + w3cp.view(1, p, 1) * Sp.unsqueeze(1) + w3cn.view(1, p, 1) * Sn.unsqueeze(1)           # [B,p,m]
        base_v = self.t0.unsqueeze(0) + self.t1v.unsqueeze(0) * cv.unsqueeze(1) \
            + w3vp.view(1, p, 1) * Cp.unsqueeze(1) + w3vn.view(1, p, 1) * Cn.unsqueeze(1)           # [B,p,n]
        base = torch.cat((base_c, base_v), 2)
        cadj = adj / adj.sum(1, keepdim=True).clamp_min(1e-12)             # columns sum to one
        radj = adj.transpose(1, 2) / adj.sum(2).unsqueeze(1).clamp_min(1e-12)
        mu = torch.zeros(B, p, m + n, device=A.device)
        for _ in range(self.T):
            agg_v = torch.bmm(mu[:, :, :m], cadj)                           # [B,p,n]
            agg_c = torch.bmm(mu[:, :, m:], radj)                           # [B,p,m]
            term2 = torch.cat((self.t2c @ agg_v, self.t2v @ agg_c), 2)      # variables first (B9)
            mu = F.relu(base + term2)
        u6 = self.t6c @ mu[:, :, :m].mean(2, keepdim=True) + self.t6v @ mu[:, :, m:].mean(2, keepdim=True)   # [B,p,1]
        emb = F.relu(torch.cat((u6.expand(-1, -1, m), self.t7 @ mu[:, :, :m]), 1))                             # [B,2p,m]
        emb = torch.cat((emb, cfe.transpose(1, 2)), 1)
        return (self.t8 @ emb).transpose(1, 2)                                                                  # [B,m,2]

    def _scores_complete(self, A, b, c, feats=None):
        B, m, n = A.shape
        p = self.p
        nfeat = None if feats is None else feats.to(A.device).float()         # [B,m] node features (s2v.py:100: u1 = t0 + t1 feat)
        Ab = F.normalize(torch.cat((A.double(), b.double().unsqueeze(2)), 2), p=2, dim=2).float()
        c0 = torch.cat((c.float(), torch.zeros(B, 1, device=A.device)), 1).unsqueeze(1)
        G = torch.cat((Ab, c0), 1)                                          # [B,m+1,n+1]
        W = torch.bmm(G, G.transpose(1, 2))
        W = W * (1.0 - torch.eye(m + 1, device=A.device)).unsqueeze(0)
        Wp, Wn = F.relu(W[:, :m, :m]).sum(2), F.relu(-W[:, :m, :m]).sum(2)  # [B,m]
        wc = W[:, m, :m]
        sp, sn = F.relu(wc).sum(1), F.relu(-wc).sum(1)                      # [B]
        t4rr = self.t4rr[:, 0]
        w3p, w3n = self.t3rr @ F.relu(t4rr), self.t3rr @ F.relu(-t4rr)
        relu_rc = F.relu(self.t4rc).unsqueeze(0) * sp.unsqueeze(1) + F.relu(-self.t4rc).unsqueeze(0) * sn.unsqueeze(1)
        scal = relu_rc @ self.t4rc                                          # [B]   (B10)
        relu_cr = F.relu(self.t4cr).unsqueeze(0) * sp.unsqueeze(1) + F.relu(-self.t4cr).unsqueeze(0) * sn.unsqueeze(1)
        u3c = relu_cr @ self.t3cr.t()                                       # [B,p]
        u1 = (self.t0 + self.t1).unsqueeze(0) if nfeat is None else self.t0.unsqueeze(0) + self.t1.unsqueeze(0) * nfeat.unsqueeze(1)
        base_r = u1 + w3p.view(1, p, 1) * Wp.unsqueeze(1) + w3n.view(1, p, 1) * Wn.unsqueeze(1) \
            + scal.view(B, 1, 1)
        mu = torch.zeros(B, p, m + 1, device=A.device)
        for _ in range(self.T):
            mur, muc = mu[:, :, :m], mu[:, :, m:]
            ur = base_r + self.t2rr @ mur + self.t2rc @ muc
            ucst = self.t0.unsqueeze(0) + self.t2cr @ mur.mean(2, keepdim=True) + u3c.unsqueeze(2)
            mu = F.relu(torch.cat((ur, ucst), 2))
        u6 = self.t6r @ mu[:, :, :m].mean(2, keepdim=True) + self.t6c @ mu[:, :, m:]
        feats = F.relu(torch.cat((u6.expand(-1, -1, m), self.t7 @ mu[:, :, :m]), 1))
        return (self.t8 @ feats).transpose(1, 2)

    # ------------------------------------------------------------------------------------------------------------
    # reference-format single item
    # ------------------------------------------------------------------------------------------------------------
    def _item_feats(self, item):
        """Node flags of an item when they are not the random-LP ones (then the CUDA kernels do not apply), else None."""
        if self.graph == 'complete':
            nf = item.get('node_features')
            if nf is None:
                return None
            nf = torch.as_tensor(np.asarray(nf)).reshape(-1).float()
            m = nf.numel() - 1                                   # trailing entry = the cost node
            return None if bool((nf[:m] == 1).all()) else nf[:m].reshape(1, m)
        cf = item['c_feats']
        cf = cf.reshape(-1, cf.shape[-1])
        if bool((cf[:, 0] == 1).all()) and bool((cf[:, 2] == 0).all()):
            return None
        return (cf[:, 0].reshape(1, -1).float(), cf[:, 2].reshape(1, -1).float())

    def _item_to_abc(self, item):
        if self.graph == 'complete':
            lp = item['lp'] if 'lp' in item else item
            A, b, c = [torch.as_tensor(np.asarray(lp[k])) if not torch.is_tensor(lp[k]) else lp[k] for k in ('A', 'b', 'c')]
            A = A.reshape(1, A.shape[-2], A.shape[-1])
            return A.double(), b.reshape(1, -1).double(), c.reshape(1, -1).double()
        dims = item['dims']
        m, n = int(dims['m']), int(dims['n'])
        cf = item['c_feats']
        cf = cf.reshape(-1, cf.shape[-1])
        A = torch.zeros(m, n, dtype=torch.float64)
        idx = torch.as_tensor([[int(q) for q in pr] for pr in item['e_feats']['i']], dtype=torch.long).reshape(-1, 2)
        A[idx[:, 0], idx[:, 1]] = torch.as_tensor([float(q) for q in item['e_feats']['coeffs']], dtype=torch.float64)
        b = cf[:, 1].double().cpu()
        c = item['v_feats'].reshape(-1).double().cpu()
        return A.unsqueeze(0), b.unsqueeze(0), c.unsqueeze(0)

    def loss_and_grad_item(self, item, weight):
        """``loss_and_grad_batch`` for ONE reference-format item (MPS / PLNN LPs: any shape, equality / bound flags, an
        ``in_loss`` subset of the rows): what ``criterion(model(item), y).backward()`` does in the reference's loop
        (train.py:59-65), on the hand-written kernels."""
        A, b, c = self._item_to_abc(item)
        dev = self.t0.device
        m = A.shape[1]
        lab = item['node_labels'] if self.graph == 'complete' else item['c_labels']
        lab = torch.as_tensor(np.asarray(lab)).reshape(-1) if not torch.is_tensor(lab) else lab.reshape(-1)
        in_loss = torch.as_tensor([int(q) for q in item['in_loss']], dtype=torch.long)
        y = torch.full((m,), 2, dtype=torch.uint8)
        y[in_loss] = lab[in_loss].to(torch.uint8)
        feats = self._item_feats(item)
        if feats is not None:
            feats = feats.to(dev) if torch.is_tensor(feats) else tuple(f.to(dev) for f in feats)
        return self.loss_and_grad_batch(A.to(dev), b.to(dev), c.to(dev), y.reshape(1, m).to(dev), weight, feats)

    def forward(self, item):
        if self.graph not in GRAPH_CODE:
            raise ValueError('Graph not recognised')
        A, b, c = self._item_to_abc(item)
        dev = self.t0.device
        feats = self._item_feats(item)
        if feats is None:
            logp = self.forward_batch(A.to(dev), b.to(dev), c.to(dev))
        else:
            # MPS / PLNN items (equality rows, bound rows): the same CUDA kernels with the item's node flags
            # (ddb_s2v_forward_flags_dev); the differentiable torch restatement only while autograd is recording
            feats = feats.to(dev) if torch.is_tensor(feats) else tuple(f.to(dev) for f in feats)
            if self.force_torch or (torch.is_grad_enabled() and any(q.requires_grad for q in self.parameters())):
                logp = self.forward_batch_torch(A.to(dev), b.to(dev), c.to(dev), feats)
            else:
                logp = self.forward_batch_cuda(A.to(dev), b.to(dev), c.to(dev), feats)
        in_loss = [int(q) for q in item['in_loss']]
        self.probs = self.probs[0, in_loss]
        return logp[0, in_loss]
