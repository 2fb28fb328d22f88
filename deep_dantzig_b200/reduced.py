"""BASELINE.json config 4: batched classifier inference + reduced-LP solve on the predicted active set, certified
against the full LP.  New capability behind the reference's timing harness: the reference only compares the forward
time with the stored solver time (src/data/plnn_stats.py:124-134) and picks the recall-1 probability threshold so that
no active constraint is dropped (src/ml/train.py:102-116, 138-140); it never solves the reduced LP.

Certificate (no trust in the classifier needed): the reduced LP is a relaxation of the full LP, so if its optimum x*
exists and satisfies every one of the m rows (``violations == 0``), x* is optimal for the full LP and the labels the
kernel evaluates on all m rows are the full LP's labels.  Instances without the certificate (unbounded or infeasible
reduced LP, violated rows) are re-solved with all rows, on the device."""
import torch

from . import solver


def predict_row_mask(model, A, b, c, prob_thresh):
    """Classifier forward (CUDA kernel) -> uint8 mask [B,m] of the rows predicted active at `prob_thresh`."""
    with torch.no_grad():
        model.forward_batch_cuda(A, b, c)
    return (model.probs[..., 1] >= prob_thresh).to(torch.uint8).contiguous()


def solve_reduced_certified(A, b, c, row_mask, threshold=solver.DEFAULT_THRESHOLD):
    """Reduced solve + certificate + full re-solve of the uncertified instances.  Returns the SolveResult of the batch
    (same fields as ``solver.solve_label``) plus ``certified`` [B] bool (result came from the reduced LP)."""
    res = solver.solve_label(A, b, c, threshold=threshold, row_mask=row_mask)
    certified = (res['status'] == solver.ST_OPTIMAL) & (res['violations'] == 0)
    redo = (~certified).nonzero().flatten()
    if redo.numel() > 0:
        full = solver.solve_label(A[redo].contiguous(), b[redo].contiguous(), c[redo].contiguous(), threshold=threshold)
        for k in ('status', 'x', 'obj', 'labels', 'n_active', 'pivots', 'ties', 'violations'):
            res[k][redo] = full[k]
    res['certified'] = certified
    return res


def timing_forward_pass(model, A, b, c, prob_thresh, threshold=solver.DEFAULT_THRESHOLD, repeats=1):
    """Device-timed comparison on one resident batch: full solve vs classifier forward + reduced solve (+ re-solves).
    Mirrors what plnn_stats.py:80-149 reports (forward time / solver time) and adds the reduced-solve leg.
    Returns a dict of milliseconds, rates and agreement counts (labels/status of the pruned path == full solve)."""
    ev = lambda: torch.cuda.Event(enable_timing=True)
    B, m, n = A.shape
    best = None
    for _ in range(max(1, repeats)):
        e = [ev() for _ in range(5)]
        e[0].record()
        full = solver.solve_label(A, b, c, threshold=threshold)
        e[1].record()
        mask = predict_row_mask(model, A, b, c, prob_thresh)
        e[2].record()
        red = solver.solve_label(A, b, c, threshold=threshold, row_mask=mask)
        e[3].record()
        certified = (red['status'] == solver.ST_OPTIMAL) & (red['violations'] == 0)
        redo = (~certified).nonzero().flatten()
        if redo.numel() > 0:
            again = solver.solve_label(A[redo].contiguous(), b[redo].contiguous(), c[redo].contiguous(), threshold=threshold)
            for k in ('status', 'x', 'obj', 'labels', 'n_active'):
                red[k][redo] = again[k]
        e[4].record()
        torch.cuda.synchronize()
        t = {'full_ms': e[0].elapsed_time(e[1]), 'forward_ms': e[1].elapsed_time(e[2]),
             'reduced_ms': e[2].elapsed_time(e[3]), 'resolve_ms': e[3].elapsed_time(e[4])}
        if best is None or t['forward_ms'] + t['reduced_ms'] + t['resolve_ms'] < best['forward_ms'] + best['reduced_ms'] + best['resolve_ms']:
            best = t
    opt_full = full['status'] == solver.ST_OPTIMAL
    same_status = ((red['status'] == solver.ST_OPTIMAL) == opt_full)
    same_labels = (red['labels'] == full['labels']).all(dim=1)
    relx = ((red['x'] - full['x']).abs().amax(dim=1) / full['x'].abs().amax(dim=1).clamp_min(1e-300))[opt_full]
    pruned_ms = best['forward_ms'] + best['reduced_ms'] + best['resolve_ms']
    best.update({
        'instances': B, 'm': m, 'n': n, 'prob_thresh': float(prob_thresh),
        'rows_kept_frac': float(mask.float().mean()),
        'certified_frac': float(certified.float().mean()),
        'certified_frac_of_optimal': float(certified[opt_full].float().mean()) if bool(opt_full.any()) else 0.0,
        'status_match': int(same_status.sum()), 'label_match': int(same_labels.sum()),
        'max_rel_x_diff': float(relx.max()) if relx.numel() else 0.0,
        'forward_over_full': best['forward_ms'] / best['full_ms'],
        'pruned_over_full': pruned_ms / best['full_ms'],
        'full_lps_per_sec': B / best['full_ms'] * 1e3, 'pruned_lps_per_sec': B / pruned_ms * 1e3})
    return best
