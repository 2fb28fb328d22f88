"""BASELINE.json configs[1] at its stated size (not collected by pytest: no test_ prefix): 1 000 000 random LPs,
m = 200, n = 100, fp64, generated (Philox, counter = global instance index), solved and labelled on the device; one JSON
line.  The whole job's labels / status / objective / x stay resident (the job's product: 200 MB + 4 MB + 8 MB + 800 MB).
Parity subset as SURVEY.md section 8(d) defines it: the first 10 000 instances downloaded and re-solved by the CPU oracle
(HiGHS dual simplex + the reference's labelling, gurobi_lp.py:435-443) and 1 000 numpy-seeded instances
(seeds = 0 + 685 i, randomlp_dataset.py:37-42 with seed 0) through the host-buffer entry point.  Full-size properties:
every optimal instance carries exactly n labels, n_active equals the label row sum, status in {optimal, unbounded}.
    python tests/run_config2_full.py [total] [parity_count]          (or under torchrun: the index range is sharded)
Lives under tests/ because it executes oracle/ (test infrastructure)."""
import json, multiprocessing as mp, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

M, N, KEY, CHUNK = 200, 100, 2026, 32768


def _worker(args):
    os.environ['OMP_NUM_THREADS'] = '1'
    from oracle import randomlp as oracle
    A, b, c = args
    r = oracle.solve_batch(A, b, c)
    return r['status'], r['labels'], r['obj'], r['x']


def _exact_vertex(A, b, rows):
    """Vertex of the given active set in extended precision (float128 iterative refinement of A_B x = b_B)."""
    AB = A[rows].astype(np.longdouble); bB = b[rows].astype(np.longdouble)
    x = np.linalg.solve(A[rows], b[rows]).astype(np.longdouble)
    for _ in range(4):
        x = x + np.linalg.solve(A[rows], (bB - AB @ x).astype(np.float64)).astype(np.longdouble)
    return x


def _explain(A, b, c, gx, glab, cx, clab, i):
    """Which side is off on a label mismatch: both x against the exact vertex of the GPU's active set, primal feasibility of
    that vertex over all m rows and its dual feasibility (multipliers of the active rows <= 0 for min c'x, Ax <= b)."""
    rows = np.nonzero(glab[i])[0]
    xe = _exact_vertex(A[i], b[i], rows)
    sl = b[i].astype(np.longdouble) - A[i].astype(np.longdouble) @ xe
    y = np.linalg.solve(A[i][rows].T, c[i])
    scale = float(np.abs(xe).max())
    return {'instance': int(i), 'gpu_labels': int(glab[i].sum()), 'highs_labels': int(clab[i].sum()),
            'gpu_rel_err_vs_exact': float(np.abs(gx[i] - xe).max() / scale),
            'highs_rel_err_vs_exact': float(np.abs(cx[i] - xe).max() / scale),
            'gpu_vertex_feasible_all_rows': bool(sl.min() > -1e-12), 'gpu_vertex_dual_feasible': bool(y.max() <= 1e-12),
            'highs_max_abs_slack_on_gpu_active_rows': float(np.abs(b[i][rows] - A[i][rows] @ cx[i]).max())}


def _compare(g, cpu_parts, inst=None):
    cst = np.concatenate([o[0] for o in cpu_parts]); clab = np.concatenate([o[1] for o in cpu_parts])
    cobj = np.concatenate([o[2] for o in cpu_parts]); cx = np.concatenate([o[3] for o in cpu_parts])
    gst, glab, gobj, gx, ties = g
    opt = cst == 2
    same_status = (gst == 2) == opt
    same_labels = (glab == clab).all(axis=1)
    both = opt & (gst == 2)
    relobj = np.abs(gobj[both] - cobj[both]) / np.maximum(np.abs(cobj[both]), 1e-300)
    relx = np.abs(gx[both] - cx[both]).max(axis=1) / np.maximum(np.abs(cx[both]).max(axis=1), 1e-300)
    mism = np.nonzero(same_status & ~same_labels)[0]
    expl = [_explain(inst[0], inst[1], inst[2], gx, glab, cx, clab, i) for i in mism[:8]] if inst is not None else []
    return {'instances': int(len(cst)), 'optimal_cpu': int(opt.sum()), 'status_mismatch': int((~same_status).sum()),
            'label_mismatch_instances': int((same_status & ~same_labels).sum()),
            'label_match_pct': 100.0 * float((same_status & same_labels).mean()),
            'max_rel_obj_diff': float(relobj.max()), 'max_rel_x_diff': float(relx.max()), 'ties_reported': int(ties[both].sum()),
            'label_mismatches_explained': expl}


def main():
    import torch
    import torch.distributed as dist
    from deep_dantzig_b200 import solver
    from oracle import randomlp as oracle
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    npar = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
    cores = os.cpu_count() or 1
    pool = mp.get_context('fork').Pool(cores) if rank == 0 else None      # forked before the CUDA context exists
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    lo, hi = rank * total // world, (rank + 1) * total // world          # contiguous block of the global index range
    cnt = hi - lo
    full = solver._alloc_outputs(cnt, M, N, dev)                           # the job's product, resident
    solver.generate_solve_label(KEY, 0, 4096, M, N, device=dev)           # warm-up (context, scratch, clocks)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launches = 0
    for off in range(0, cnt, CHUNK):
        nb = min(CHUNK, cnt - off)
        view = solver.SolveResult({k: v[off:off + nb] for k, v in full.items()})
        solver.generate_solve_label(KEY, lo + off, nb, M, N, device=dev, out=view)
        launches += 1
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    st = full['status']; nact = full['n_active']; piv = full['pivots']
    optm = st == 2
    stats = torch.stack([optm.sum(), (st == 5).sum(), ((st != 2) & (st != 5)).sum(), (optm & (nact == N)).sum(),
                         full['ties'][optm].sum(), (optm & (full['labels'].sum(dim=1, dtype=torch.int64) != nact)).sum(), full['labels'].sum(dtype=torch.int64),
                         piv[:, 3].sum(dtype=torch.int64), piv[optm][:, 3].sum(dtype=torch.int64)]).to(torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats)
    if rank != 0:
        dist.destroy_process_group()
        return
    s = [int(v) for v in stats.tolist()]
    line = {'config': 'BASELINE.json configs[1]: %d random LPs m=200 n=100 fp64 generated + solved + labelled' % total,
            'n_gpus': world, 'instances': total, 'chunk': CHUNK, 'calls_per_rank': launches, 'philox_key': KEY,
            'device_seconds': float(ms.item()) / 1e3, 'lps_per_sec': total / (float(ms.item()) / 1e3),
            'optimal': s[0], 'unbounded': s[1], 'other_status': s[2], 'optimal_with_exactly_n_labels': s[3],
            'ties_on_optimal': s[4], 'label_count_inconsistent': s[5], 'labels_set': s[6],
            'mean_pivots_all': s[7] / total, 'mean_pivots_optimal': s[8] / max(s[0], 1)}
    # ---- parity subset 1: the first npar instances of the job, downloaded and re-solved by the oracle ----
    A, b, c = solver.generate(KEY, 0, npar, M, N, device=dev)
    A, b, c = A.cpu().numpy(), b.cpu().numpy(), c.cpu().numpy()
    g = tuple(full[k][:npar].cpu().numpy() for k in ('status', 'labels', 'obj', 'x', 'ties'))
    parts = np.array_split(np.arange(npar), cores * 8)
    t0 = time.perf_counter()
    out = pool.map(_worker, [(A[p], b[p], c[p]) for p in parts])
    dt = time.perf_counter() - t0
    line['parity_first_instances'] = dict(_compare(g, out, (A, b, c)), cpu_seconds=dt, cpu_cores=cores, cpu_lps_per_sec=npar / dt,
                                          oracle='scipy HiGHS dual simplex + reference labelling (1e-7 threshold)')
    # ---- parity subset 2: 1 000 numpy-seeded instances (the reference's own generator bits), host-buffer entry point ----
    seeds = oracle.seed_schedule(0, 1000)
    inst = [oracle.generate_instance(M, N, int(sd)) for sd in seeds]
    A = np.ascontiguousarray(np.stack([i[0] for i in inst])); b = np.ascontiguousarray(np.stack([i[1] for i in inst]))
    c = np.ascontiguousarray(np.stack([i[2] for i in inst]))
    r = solver.solve_label_host(A, b, c, device=local)
    g = (np.asarray(r['status']), np.asarray(r['labels']), np.asarray(r['obj']), np.asarray(r['x']), np.asarray(r['ties']))
    parts = np.array_split(np.arange(1000), cores * 4)
    out = pool.map(_worker, [(A[p], b[p], c[p]) for p in parts])
    line['parity_numpy_seeded'] = dict(_compare(g, out, (A, b, c)), seed_step=int(seeds[1] - seeds[0]))
    pool.close(); pool.join()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
