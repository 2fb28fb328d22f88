"""Which side is off when GPU and HiGHS disagree by more than 1e-9?  For the worst instances of a large sample, compute
the vertex of the agreed active set in extended precision (float128 iterative refinement of A_B x = b_B) and measure both
against it.  Not collected by pytest; executes oracle/ (test infrastructure).   python tests/parity_outliers.py [count]"""
import json, multiprocessing as mp, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def _worker(args):
    os.environ['OMP_NUM_THREADS'] = '1'
    from oracle import randomlp as oracle
    A, b, c = args
    r = oracle.solve_batch(A, b, c)
    return r['status'], r['labels'], r['obj'], r['x']


def exact_vertex(A, b, rows):
    AB = A[rows].astype(np.longdouble); bB = b[rows].astype(np.longdouble)
    x = np.linalg.solve(A[rows], b[rows]).astype(np.longdouble)
    for _ in range(4):
        r = bB - AB @ x
        x = x + np.linalg.solve(A[rows], r.astype(np.float64)).astype(np.longdouble)
    return x


def main():
    import torch
    from deep_dantzig_b200 import solver
    m, n, N = 200, 100, int(sys.argv[1]) if len(sys.argv) > 1 else 40960
    cores = os.cpu_count() or 1
    res = solver.generate_solve_label(777, 0, N, m, n, keep_instances=True)
    A, b, c = res['A'].cpu().numpy(), res['b'].cpu().numpy(), res['c'].cpu().numpy()
    parts = np.array_split(np.arange(N), cores * 8)
    with mp.get_context('fork').Pool(cores) as pool:
        out = pool.map(_worker, [(A[p], b[p], c[p]) for p in parts])
    cst = np.concatenate([o[0] for o in out]); clab = np.concatenate([o[1] for o in out]); cx = np.concatenate([o[3] for o in out])
    gst = res['status'].cpu().numpy(); glab = res['labels'].cpu().numpy(); gx = res['x'].cpu().numpy()
    both = np.nonzero((cst == 2) & (gst == 2))[0]
    relx = np.abs(gx[both] - cx[both]).max(axis=1) / np.abs(cx[both]).max(axis=1)
    worst = both[np.argsort(-relx)[:12]]
    mism = both[(glab[both] != clab[both]).any(axis=1)]
    rows_out = []
    for i in list(dict.fromkeys(list(mism) + list(worst))):
        rows = np.nonzero(glab[i])[0]                       # the GPU's active set (exactly n rows)
        xe = exact_vertex(A[i], b[i], rows)
        sl = (b[i].astype(np.longdouble) - A[i].astype(np.longdouble) @ xe)
        feasible = bool(sl.min() > -1e-12)                  # the GPU's vertex is primal feasible for all m rows
        scale = float(np.abs(xe).max())
        rows_out.append({'instance': int(i), 'gpu_labels': int(glab[i].sum()), 'highs_labels': int(clab[i].sum()),
                         'labels_equal': bool((glab[i] == clab[i]).all()), 'max_abs_x': scale,
                         'cond_AB': float(np.linalg.cond(A[i][rows])),
                         'gpu_rel_err_vs_exact': float(np.abs(gx[i] - xe).max() / scale),
                         'highs_rel_err_vs_exact': float(np.abs(cx[i] - xe).max() / scale),
                         'gpu_vertex_feasible_all_rows': feasible,
                         'highs_max_abs_slack_on_gpu_active_rows': float(np.abs(b[i][rows] - A[i][rows] @ cx[i]).max())})
    print(json.dumps({'shape': [m, n], 'instances': N, 'label_mismatch_instances': [int(q) for q in mism], 'outliers': rows_out}))


if __name__ == '__main__':
    main()
