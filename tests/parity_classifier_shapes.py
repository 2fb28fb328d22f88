"""Differential sweep of the classifier kernels over shapes / embedding sizes / rounds (not collected by pytest):
forward (dense streaming kernel or general kernel, whichever the library picks) vs the oracle on sampled instances, and the
loss+gradient kernel vs autograd over the batched restatement.  python tests/parity_classifier_shapes.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    from deep_dantzig_b200 import solver
    from deep_dantzig_b200.ml.models.s2v import Model
    from oracle import classifier as oc
    cases = [(200, 100, 40, 3), (200, 100, 64, 3), (200, 100, 13, 4), (50, 20, 12, 3), (50, 20, 13, 1), (8, 4, 3, 2), (16, 2, 5, 1), (12, 24, 6, 2),
             (100, 128, 16, 2), (33, 16, 7, 0), (64, 50, 33, 2), (300, 56, 20, 2), (400, 104, 24, 3), (37, 19, 13, 2), (500, 250, 40, 3), (31, 100, 9, 3)]
    bad = 0
    for (m, n, p, T) in cases:
        torch.manual_seed(m * 7 + p)
        B = 96
        A, b, c = solver.generate(31, 0, B, m, n)
        y = (torch.rand(B, m, device='cuda') < 0.4).to(torch.uint8)
        line = {'shape': [m, n], 'p': p, 'T': T}
        for graph in ('bipartite', 'complete'):
            model = Model(graph, p, T, on_cuda=True, verbose_init=False)
            try:
                with torch.no_grad():
                    lp = model.forward_batch(A, b, c).cpu().numpy()
            except Exception as exc:      # shapes a kernel does not hold must say so loudly, never compute something else
                line[graph + '_forward'] = 'unsupported: %s' % str(exc)[:90]
                continue
            P = {k: v.detach().cpu() for k, v in model.named_parameters()}
            worst = 0.0
            for k in (0, 17, B - 1):
                ref, _ = oc.forward(graph, P, A[k].cpu().numpy(), b[k].cpu().numpy(), c[k].cpu().numpy(), T)
                d = np.abs(lp[k] - ref.numpy())
                worst = max(worst, float((d / (5e-5 + 1e-5 * np.abs(ref.numpy()))).max()))
            line[graph + '_forward_err_over_tol'] = worst
            bad += worst > 1.0
        model = Model('bipartite', p, T, on_cuda=True, verbose_init=False)
        if p <= 64:
            try:
                model.zero_grad()
                l_dev = float(model.loss_and_grad_batch(A, b, c, y, [0.3, 0.7]))
                g_dev = torch.cat([q.grad.reshape(-1) for q in model.parameters()]).clone()
                model.zero_grad()
                crit = torch.nn.NLLLoss(weight=torch.tensor([0.3, 0.7], device='cuda'), reduction='sum')
                l_ref = crit(model.forward_batch_torch(A, b, c).reshape(-1, 2), y.long().reshape(-1))
                l_ref.backward()
                g_ref = torch.cat([q.grad.reshape(-1) for q in model.parameters()])
                line['grad_rel_err'] = float((g_dev - g_ref).abs().max() / g_ref.abs().max())
                line['loss_rel_err'] = abs(l_dev - float(l_ref)) / abs(float(l_ref))
                bad += line['grad_rel_err'] > 5e-4 or line['loss_rel_err'] > 2e-4
            except Exception as exc:      # shapes the backward kernel does not hold report that loudly
                line['grad'] = 'unsupported: %s' % str(exc)[:80]
        print(json.dumps(line), flush=True)
    print('SUMMARY: %d cases, %d outside tolerance' % (len(cases), bad))


if __name__ == '__main__':
    main()
