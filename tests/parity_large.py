"""Large-sample parity run (not collected by pytest: no test_ prefix): GPU generate -> solve -> label against the CPU
oracle (HiGHS dual simplex, polished to the certified vertex, + the reference's labelling) on tens of thousands of instances; one JSON line per shape.
    python tests/parity_large.py [count_200x100] [count_50x20]
Lives under tests/ because it executes oracle/ (test infrastructure)."""
import json, multiprocessing as mp, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def _worker(args):
    os.environ['OMP_NUM_THREADS'] = '1'
    from oracle import randomlp as oracle
    A, b, c = args
    r = oracle.solve_batch(A, b, c)
    return r['status'], r['labels'], r['obj'], r['x'], r['oracle_tie'], r['certified']


def main():
    import torch
    from deep_dantzig_b200 import solver
    counts = {(200, 100): int(sys.argv[1]) if len(sys.argv) > 1 else 40960, (50, 20): int(sys.argv[2]) if len(sys.argv) > 2 else 131072}
    cores = os.cpu_count() or 1
    pool = mp.get_context('fork').Pool(cores)
    for (m, n), N in counts.items():
        res = solver.generate_solve_label(777, 0, N, m, n, keep_instances=True)
        torch.cuda.synchronize()
        A, b, c = res['A'].cpu().numpy(), res['b'].cpu().numpy(), res['c'].cpu().numpy()
        parts = np.array_split(np.arange(N), cores * 8)
        t0 = time.perf_counter()
        out = pool.map(_worker, [(A[p], b[p], c[p]) for p in parts])
        dt = time.perf_counter() - t0
        cst = np.concatenate([o[0] for o in out]); clab = np.concatenate([o[1] for o in out])
        cobj = np.concatenate([o[2] for o in out]); cx = np.concatenate([o[3] for o in out])
        otie = np.concatenate([o[4] for o in out]); cert = np.concatenate([o[5] for o in out])
        gst = res['status'].cpu().numpy(); glab = res['labels'].cpu().numpy()
        gobj = res['obj'].cpu().numpy(); gx = res['x'].cpu().numpy(); ties = res['ties'].cpu().numpy()
        opt = cst == 2
        same_status = (gst == 2) == opt
        same_labels = (glab == clab).all(axis=1)
        both = opt & (gst == 2)
        relobj = np.abs(gobj[both] - cobj[both]) / np.maximum(np.abs(cobj[both]), 1e-300)
        relx = np.abs(gx[both] - cx[both]).max(axis=1) / np.maximum(np.abs(cx[both]).max(axis=1), 1e-300)
        nact = glab[both].sum(axis=1)
        print(json.dumps({'shape': [m, n], 'instances': N, 'optimal_cpu': int(opt.sum()), 'status_mismatch': int((~same_status).sum()),
                          'label_mismatch_instances': int((same_status & ~same_labels).sum()),
                          'label_match_pct': 100.0 * float((same_status & same_labels).mean()),
                          'max_rel_obj_diff': float(relobj.max()), 'max_rel_x_diff': float(relx.max()),
                          'optimal_with_exactly_n_labels': int((nact == n).sum()), 'ties_reported': int(ties[both].sum()),
                          'oracle_ties': int(otie.sum()), 'oracle_uncertified': int((opt & ~cert.astype(bool)).sum()),
                          'cpu_seconds': dt, 'cpu_cores': cores,
                          'oracle': 'scipy HiGHS dual simplex, then the certified extended-precision vertex of its active set (oracle/randomlp.py: polish_vertex) + reference labelling (1e-7 threshold)'}), flush=True)
    pool.close(); pool.join()


if __name__ == '__main__':
    main()
