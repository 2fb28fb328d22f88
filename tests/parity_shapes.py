"""Differential sweep over many shapes (not collected by pytest): GPU solve vs the oracle on 150 instances per shape,
including the awkward ones (n = 1, m = n, m < n, odd sizes, m >> n).  python tests/parity_shapes.py"""
import json, multiprocessing as mp, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def _worker(args):
    os.environ['OMP_NUM_THREADS'] = '1'
    from oracle import randomlp as oracle
    A, b, c = args
    r = oracle.solve_batch(A, b, c)
    return r['status'], r['labels'], r['obj'], r['x']


def main():
    import torch
    from deep_dantzig_b200 import solver, _lib
    shapes = [(1, 1), (2, 1), (3, 2), (5, 5), (4, 7), (7, 3), (9, 8), (16, 8), (17, 9), (33, 7), (40, 39), (41, 40), (64, 32), (65, 33), (100, 10),
              (128, 64), (129, 63), (150, 100), (199, 99), (201, 101), (220, 110), (222, 111), (230, 112), (256, 100), (300, 100), (356, 100),
              (357, 100), (400, 60), (90, 70), (260, 130)]
    N = 150
    cores = os.cpu_count() or 1
    bad = []
    with mp.get_context('fork').Pool(cores) as pool:
        for (m, n) in shapes:
            res = solver.generate_solve_label(4242, 0, N, m, n, keep_instances=True)
            A, b, c = res['A'].cpu().numpy(), res['b'].cpu().numpy(), res['c'].cpu().numpy()
            parts = np.array_split(np.arange(N), cores)
            out = pool.map(_worker, [(A[p], b[p], c[p]) for p in parts if len(p)])
            cst = np.concatenate([o[0] for o in out]); clab = np.concatenate([o[1] for o in out]); cx = np.concatenate([o[3] for o in out])
            gst = res['status'].cpu().numpy(); glab = res['labels'].cpu().numpy(); gx = res['x'].cpu().numpy()
            opt = cst == 2
            st_bad = int((((gst == 2) != opt)).sum())
            both = opt & (gst == 2)
            lab_bad = int((glab[both] != clab[both]).any(axis=1).sum())
            relx = float((np.abs(gx[both] - cx[both]).max(axis=1) / np.maximum(np.abs(cx[both]).max(axis=1), 1e-300)).max()) if both.any() else 0.0
            plan = _lib.context(0).solve_plan(m, n)
            line = {'shape': [m, n], 'plan': plan, 'optimal': int(opt.sum()), 'status_mismatch': st_bad, 'label_mismatch': lab_bad, 'max_rel_x': relx,
                    'gpu_status_hist': {int(k): int(v) for k, v in zip(*np.unique(gst, return_counts=True))}}
            print(json.dumps(line), flush=True)
            if st_bad or lab_bad or relx > 1e-7:
                bad.append(line)
    print('SUMMARY: %d shapes, %d with a mismatch' % (len(shapes), len(bad)))


if __name__ == '__main__':
    main()
