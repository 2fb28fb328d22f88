"""GPU parity tests of the generate -> solve -> label path: CUDA (through the C ABI) vs the CPU oracle.

Bars (BASELINE.json north_star): status predicate equal on every instance; active-set labels bit-exact on
non-degenerate instances; objective and x within 1e-9 relative; ties counted and reported, not hidden."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import randomlp as oracle
from oracle import philox

pytestmark = pytest.mark.gpu

REL_TOL = 1e-9          # north_star: objectives and primal values within 1e-9 relative in fp64


def _numpy_batch(m, n, seeds):
    A = np.empty((len(seeds), m, n)); b = np.empty((len(seeds), m)); c = np.empty((len(seeds), n))
    for i, s in enumerate(seeds):
        A[i], b[i], c[i] = oracle.generate_instance(m, n, s)
    return A, b, c


def _dev(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]


def _check_against_oracle(res, ref, A, b, c, rel_tol=REL_TOL, what=''):
    """The bar of BASELINE.json north_star against the POLISHED oracle (HiGHS dual simplex, then the certified
    extended-precision vertex of its active set -- oracle.randomlp.polish_vertex): status predicate equal everywhere,
    labels bit-exact, x and objective within 1e-9 relative.  Prints what it tolerates nowhere but counts: ties the kernel
    reported, and instances where HiGHS' raw x would have been labelled differently from its own polished vertex."""
    st = np.asarray(res['status']); lab = np.asarray(res['labels'])
    ok_ref = ref['status'] == 2
    assert ((st == 2) == ok_ref).all(), 'status predicate differs on %s' % np.flatnonzero((st == 2) != ok_ref)[:10]
    assert set(np.unique(st[~ok_ref])) <= {5}, 'non-optimal instances of this generator are unbounded'
    assert (lab[~ok_ref] == 0).all()
    mism = np.flatnonzero((lab[ok_ref] != ref['labels'][ok_ref]).any(axis=1))
    ties = np.asarray(res['ties'])[ok_ref]
    n_oracle_ties = int(ref['oracle_tie'].sum()) if 'oracle_tie' in ref else 0
    print('%s%d instances, %d optimal, label mismatches %d, ties reported by the kernel %d, oracle ties (raw HiGHS x labelled '
          'differently from its polished vertex) %d, uncertified %d'
          % (what + ': ' if what else '', len(st), ok_ref.sum(), len(mism), int((ties > 0).sum()), n_oracle_ties,
             int((~ref['certified'][ok_ref]).sum()) if 'certified' in ref else -1))
    assert len(mism) == 0, 'active-set mismatch on %s (kernel tie flags there: %s)' % (mism[:10], ties[mism[:10]])
    if ok_ref.any():
        obj = np.asarray(res['obj'])[ok_ref]; x = np.asarray(res['x'])[ok_ref]
        rel_o = np.abs(obj - ref['obj'][ok_ref]) / np.abs(ref['obj'][ok_ref])
        rel_x = np.abs(x - ref['x'][ok_ref]).max(axis=1) / np.abs(ref['x'][ok_ref]).max(axis=1)
        assert rel_o.max() <= rel_tol, rel_o.max()
        assert rel_x.max() <= rel_tol, rel_x.max()
        assert (np.asarray(res['n_active'])[ok_ref] == ref['n_active'][ok_ref]).all()
    return int(ok_ref.sum())


def _check_against_fixture(r, g, rel_tol=REL_TOL):
    m, n = int(g['m']), int(g['n'])
    want = np.unpackbits(g['labels_packed'], axis=1)[:, :m]
    ok = g['status'] == 2
    assert ((r['status'] == 2) == ok).all(), np.flatnonzero((r['status'] == 2) != ok)[:10]
    assert set(np.unique(r['status'][~ok])) <= {5}
    mism = np.flatnonzero((r['labels'] != want).any(axis=1))
    assert len(mism) == 0, 'label mismatch on instances %s' % mism[:10]
    assert (r['n_active'][ok] == n).all()
    rel_o = np.abs(r['obj'][ok] - g['obj'][ok]) / np.abs(g['obj'][ok])
    assert rel_o.max() <= rel_tol, rel_o.max()
    xi = g['x_index']
    rel_x = np.abs(r['x'][xi] - g['x']).max(axis=1) / np.abs(g['x']).max(axis=1)
    assert rel_x.max() <= rel_tol, rel_x.max()
    print('%d instances (%d optimal): labels equal, max rel obj %.2e, max rel x (%d rows) %.2e, kernel ties %d, oracle ties %d'
          % (len(ok), ok.sum(), rel_o.max(), len(xi), rel_x.max(), int((r['ties'] > 0).sum()), int(g['oracle_tie'].sum())))


def _to_np(res):
    return {k: v.cpu().numpy() for k, v in res.items() if torch.is_tensor(v)}


def test_known_answer_vectors(cuda_device, golden_dir):
    from deep_dantzig_b200 import solver
    with open(os.path.join(golden_dir, 'randomlp_kat.json')) as f:
        kat = json.load(f)
    for it in kat['instances']:
        A, b, c = oracle.generate_instance(it['m'], it['n'], it['seed'])
        r = _to_np(solver.solve_label(*_dev(A[None], b[None], c[None])))
        assert (r['status'][0] == 2) == (it['status'] == 2), it
        assert list(np.flatnonzero(r['labels'][0])) == it['active'], it
        if it['status'] == 2:
            assert abs(r['obj'][0] - it['objval']) <= REL_TOL * abs(it['objval'])
        else:
            assert r['status'][0] == 5


def test_config1_golden_fixture(cuda_device, golden_dir):
    """BASELINE.json configs[0] at its stated size: (50,20), seed 3231, all 1 000 instances (seeds 3231 + 578 i) --
    against the committed fixture of the polished oracle, through the HOST-buffer ABI."""
    from deep_dantzig_b200 import solver
    g = np.load(os.path.join(golden_dir, 'randomlp_config1.npz'))
    assert len(g['seeds']) == 1000
    A, b, c = _numpy_batch(50, 20, list(g['seeds']))
    r = solver.solve_label_host(A, b, c)
    _check_against_fixture(r, g)
    assert (r['ties'] == 0).all()


def test_config2_golden_fixture(cuda_device, golden_dir):
    """BASELINE.json configs[1] parity subset: (200,100), 2 000 instances from the reference's own
    generator bits (numpy legacy stream, seeds 0 + 685 i): status, labels, objective of all, x of the first 256 optimal."""
    from deep_dantzig_b200 import solver
    g = np.load(os.path.join(golden_dir, 'randomlp_config2.npz'))
    assert len(g['seeds']) == 2000
    A, b, c = _numpy_batch(200, 100, list(g['seeds']))
    r = solver.solve_label_host(A, b, c)
    _check_against_fixture(r, g)


@pytest.mark.parametrize('name', ['randomlp_300x150.npz', 'randomlp_400x100.npz', 'randomlp_500x250.npz'])
def test_beyond_one_sm_fixtures(cuda_device, golden_dir, name):
    """Shapes whose tableau does not fit the register file / shared memory of one SM -- 64 instances each, BASELINE.json
    configs[3] shape (500,250) among them."""
    from deep_dantzig_b200 import solver
    g = np.load(os.path.join(golden_dir, name))
    assert len(g['seeds']) >= 64
    A, b, c = _numpy_batch(int(g['m']), int(g['n']), list(g['seeds']))
    r = _to_np(solver.solve_label(*_dev(A, b, c)))
    _check_against_fixture(r, g)


def test_ten_thousand_instances_hold_the_1e9_bar(cuda_device):
    """North-star bar at scale: 10 240 Philox instances of (200,100), downloaded and re-solved by the live polished oracle
    on every host core: statuses and labels equal on all of them, x and objective within 1e-9 relative."""
    from deep_dantzig_b200 import solver
    N = 10240
    r = solver.generate_solve_label(777, 0, N, 200, 100, keep_instances=True)
    A, b, c = r['A'].cpu().numpy(), r['b'].cpu().numpy(), r['c'].cpu().numpy()
    ref = oracle.solve_batch_parallel(A, b, c)
    ok = ref['status'] == 2
    assert ref['certified'][ok].all(), 'polished oracle could not certify %d instances' % (~ref['certified'][ok]).sum()
    _check_against_oracle(_to_np(r), ref, A, b, c, what='(200,100) x %d' % N)


SWEEP_CELLS = [(ratio, dens) for ratio in (1.25, 1.5, 2.0, 3.0, 4.0) for dens in (0.5, 0.1)]


@pytest.mark.parametrize('ratio,density', SWEEP_CELLS)
def test_sparse_sweep_cells_vs_oracle(cuda_device, ratio, density):
    """BASELINE.json configs[2]: every (m/n, density < 1) cell of the sweep at n = 100, 200 downloaded Philox instances per
    cell against the polished oracle.  The Bernoulli mask creates exact zeros in entering columns and near-parallel rows."""
    from deep_dantzig_b200 import solver
    n = 100; m = int(round(ratio * n)); N = 200
    r = solver.generate_solve_label(31337, 0, N, m, n, density=density, keep_instances=True)
    A, b, c = r['A'].cpu().numpy(), r['b'].cpu().numpy(), r['c'].cpu().numpy()
    assert abs((A != 0).mean() - density) < 0.02
    ref = oracle.solve_batch_parallel(A, b, c)
    _check_against_oracle(_to_np(r), ref, A, b, c, what='m/n=%.2f density=%.1f' % (ratio, density))


def test_handmade_sparse_degeneracies(cuda_device):
    """What a Bernoulli mask does at density 0.1, made by hand on every instance: an all-zero row, a duplicated row (same
    right-hand side), a zero in the first crash pivot position, an all-zero column pair avoided (keeps the LP bounded)."""
    from deep_dantzig_b200 import solver
    m, n, N = 60, 20, 96
    A, b, c = _numpy_batch(m, n, [911 * i + 5 for i in range(N)])
    rng = np.random.RandomState(3)
    for i in range(N):
        kind = i % 4
        score = (A[i] @ c[i]) / np.linalg.norm(A[i], axis=1)
        first = int(np.argsort(score)[0])                       # the row the crash takes first
        if kind == 0:                                           # all-zero row: 0 <= b_i, never active
            z = int(rng.randint(m)); A[i, z] = 0.0; b[i, z] = abs(b[i, z]) + 0.1
        elif kind == 1:                                         # exact duplicate of a row (same rhs): a degenerate vertex if active
            src, dst = rng.choice(m, 2, replace=False); A[i, dst] = A[i, src]; b[i, dst] = b[i, src]
        elif kind == 2:                                         # zero where the first crash pivot would be (largest |entry| of the first row)
            A[i, first, int(np.abs(A[i, first]).argmax())] = 0.0
        else:                                                   # a very sparse first crash row
            keep = rng.choice(n, 2, replace=False); row = np.zeros(n); row[keep] = A[i, first, keep]; A[i, first] = row
    r = solver.solve_label_host(A, b, c)
    ref = oracle.solve_batch(A, b, c)
    ok = ref['status'] == 2
    assert ((r['status'] == 2) == ok).all()
    dup = (np.arange(N) % 4 == 1)
    # duplicated active rows are genuine ties (n + 1 rows at the vertex): labels must still be what thresholding gives
    assert (r['labels'][ok] == ref['labels'][ok]).all()
    assert (r['n_active'][ok] == ref['n_active'][ok]).all()
    assert (r['n_active'][ok & ~dup] == n).all()
    deg = ok & (ref['n_active'] > n)
    assert deg.sum() >= 3 and (r['ties'][deg] > 0).all()        # reported, not hidden
    assert np.abs(r['obj'][ok] - ref['obj'][ok]).max() <= REL_TOL * np.abs(ref['obj'][ok]).max()


@pytest.mark.parametrize('m,n,N', [(10, 5, 400), (50, 20, 400), (200, 100, 160), (30, 20, 100), (120, 100, 40),
                                   (33, 17, 100), (64, 32, 100), (300, 100, 24)])
def test_parity_vs_oracle(cuda_device, m, n, N):
    from deep_dantzig_b200 import solver
    seeds = [685 * i for i in range(N)]
    A, b, c = _numpy_batch(m, n, seeds)
    r = _to_np(solver.solve_label(*_dev(A, b, c)))
    ref = oracle.solve_batch(A, b, c)
    nopt = _check_against_oracle(r, ref, A, b, c)
    assert (r['pivots'][:, 3] == r['pivots'][:, :3].sum(axis=1)).all()
    assert (r['pivots'][:, 0] == n).all()
    assert (r['violations'] == 0).all()
    assert 0 <= nopt <= N
    if m >= 2 * n:
        assert nopt > 0          # at m = 1.2 n essentially every instance is unbounded (Wendel)


def test_parity_large_global_tableau(cuda_device):
    """BASELINE.json configs[3] shape (500,250): the tableau does not fit one SM -> thread-block-cluster kernel (live tableau
    in the distributed shared memory of four SMs), and the L2/HBM-streamed plan it replaces gives the same answers."""
    from deep_dantzig_b200 import solver, _lib
    ctx = _lib.context(0)
    assert ctx.solve_plan(500, 250) == 6
    A, b, c = _numpy_batch(500, 250, [0, 1, 2, 3, 4, 5])
    r = _to_np(solver.solve_label(*_dev(A, b, c)))
    ref = oracle.solve_batch(A, b, c)
    _check_against_oracle(r, ref, A, b, c)
    try:
        ctx.set_solve_plan(2)
        r2 = _to_np(solver.solve_label(*_dev(A, b, c)))
    finally:
        ctx.set_solve_plan(-1)
    _check_against_oracle(r2, ref, A, b, c)
    assert (r['status'] == r2['status']).all() and (r['labels'] == r2['labels']).all()


@pytest.mark.parametrize('m,n,N', [(400, 100, 600), (357, 100, 300), (484, 100, 300), (470, 86, 300), (385, 73, 300),
                                   (250, 100, 400), (292, 100, 300), (229, 100, 300), (265, 80, 300),
                                   (300, 150, 300), (260, 130, 300), (280, 120, 300), (400, 150, 200), (380, 125, 200)])
def test_wide_row_variants_agree_with_the_cluster_kernel(cuda_device, m, n, N):
    """Shapes with 256 < m - n <= 384 live rows and n <= 100 (the m/n = 4 cells of the configs[2] sweep) run on the hybrid
    row-per-thread kernel with twelve warps, one LP per SM (plan 0), instead of the thread-block-cluster kernel; shapes
    with 128 < m - n <= 192 on its six-warp variant, two LPs per SM; 100 < n <= 150 on 151-column variants: same statuses, labels and pivot counts as plan 6 on
    the same Philox batch, x and objective within 1e-9."""
    from deep_dantzig_b200 import solver, _lib
    ctx = _lib.context(0)
    assert ctx.solve_plan(m, n) == 0
    dA, db, dc = solver.generate(99, 0, N, m, n)
    try:
        ctx.set_solve_plan(0)
        r0 = _to_np(solver.solve_label(dA, db, dc))
        ctx.set_solve_plan(6)
        r6 = _to_np(solver.solve_label(dA, db, dc))
    finally:
        ctx.set_solve_plan(-1)
    for k in ('status', 'labels', 'n_active'):
        assert (r0[k] == r6[k]).all(), k
    assert (r0['pivots'] == r6['pivots']).all(axis=1).mean() >= 0.99
    ok = r0['status'] == 2
    assert ok.sum() > 0
    assert np.abs(r0['x'][ok] - r6['x'][ok]).max() <= 1e-9 * np.abs(r6['x'][ok]).max()
    assert np.abs(r0['obj'][ok] - r6['obj'][ok]).max() <= 1e-9 * np.abs(r6['obj'][ok]).max()
    # the same on a sparse cell of the sweep (density 0.1) and on reduced LPs (labels + 30 % random rows kept)
    import torch
    sA, sb, sc = solver.generate(100, 0, N // 2, m, n, density=0.1)
    keep = torch.maximum(torch.from_numpy(r0['labels']).to(dA.device),
                         (torch.rand(N, m, device=dA.device) < 0.3).to(torch.uint8)).contiguous()
    try:
        ctx.set_solve_plan(0)
        s0 = _to_np(solver.solve_label(sA, sb, sc))
        m0 = _to_np(solver.solve_label(dA, db, dc, row_mask=keep))
        ctx.set_solve_plan(6)
        s6 = _to_np(solver.solve_label(sA, sb, sc))
        m6 = _to_np(solver.solve_label(dA, db, dc, row_mask=keep))
    finally:
        ctx.set_solve_plan(-1)
    for k in ('status', 'labels', 'n_active', 'violations'):
        assert (s0[k] == s6[k]).all(), ('sparse', k)
        assert (m0[k] == m6[k]).all(), ('masked', k)
    assert (m0['status'][ok] == 2).all() and (m0['violations'][ok] == 0).all() and (m0['labels'][ok] == r0['labels'][ok]).all()
    sok = s0['status'] == 2
    assert np.abs(s0['x'][sok] - s6['x'][sok]).max() <= 1e-9 * np.abs(s6['x'][sok]).max()


@pytest.mark.parametrize('m,n,N', [(300, 150, 200), (400, 100, 150), (500, 250, 80), (260, 130, 200), (450, 200, 60), (600, 300, 24),
                                   (301, 151, 100)])
def test_cluster_kernel_agrees_with_the_streamed_plan(cuda_device, m, n, N):
    """The cluster kernel (plan 6) and the global-memory plan (plan 2) implement the same algorithm on different memory:
    same statuses, labels and pivot counts on Philox batches large enough that every cluster solves several instances;
    x and objective within 1e-9."""
    from deep_dantzig_b200 import solver, _lib
    ctx = _lib.context(0)
    wide = (n <= 100 and m - n <= 384) or (n <= 150 and max(n, m - n) <= 256)      # wide row-per-thread variants of plan 0
    assert ctx.solve_plan(m, n) == (0 if wide else 6)
    dA, db, dc = solver.generate(4711, 0, N, m, n)
    try:
        ctx.set_solve_plan(6)
        r6 = _to_np(solver.solve_label(dA, db, dc))
        ctx.set_solve_plan(2)
        r2 = _to_np(solver.solve_label(dA, db, dc))
    finally:
        ctx.set_solve_plan(-1)
    for k in ('status', 'labels', 'n_active'):
        assert (r6[k] == r2[k]).all(), k
    ok = r6['status'] == 2
    assert ok.sum() > 0 or m < 2 * n
    assert (r6['pivots'][:, 0] == r2['pivots'][:, 0]).all()
    same_path = (r6['pivots'] == r2['pivots']).all(axis=1).mean()
    print('(%d,%d): identical pivot counts on %.1f %% of %d instances' % (m, n, 100 * same_path, N))
    if ok.any():
        assert np.abs(r6['x'][ok] - r2['x'][ok]).max() <= 1e-9 * np.abs(r2['x'][ok]).max()
        assert np.abs(r6['obj'][ok] - r2['obj'][ok]).max() <= 1e-9 * np.abs(r2['obj'][ok]).max()


@pytest.mark.parametrize('m,n,N,density', [(200, 100, 3000, 1.0), (228, 100, 400, 1.0), (100, 100, 300, 1.0), (101, 100, 300, 1.0),
                                           (173, 87, 400, 1.0), (144, 72, 400, 1.0), (151, 75, 300, 1.0), (90, 30, 500, 1.0),
                                           (12, 5, 300, 1.0), (200, 100, 600, 0.5), (150, 100, 400, 0.1), (228, 100, 300, 0.1)])
def test_column_block_kernel_agrees_with_row_per_thread(cuda_device, m, n, N, density):
    """Plan 7 (rows over lanes, columns over warps; the automatic choice for 72 <= n <= 100) runs the algorithm of plan 0 with the
    same arithmetic per tableau entry: same statuses, labels, ties and PIVOT COUNTS on every instance, x / objective to 1e-9
    (in practice to the last bits) -- at its largest tile (m - n = 128), with no live rows (m = n), at odd and small n, and on
    sparse instances, where singular static crash bases are flagged by both kernels and re-solved by the generic one."""
    from deep_dantzig_b200 import solver, _lib
    ctx = _lib.context(0)
    assert ctx.solve_plan(m, n) == (7 if n >= 72 else 0)
    dA, db, dc = solver.generate(9001, 0, N, m, n, density=density)
    try:
        ctx.set_solve_plan(7)
        r7 = _to_np(solver.solve_label(dA, db, dc))
        ctx.set_solve_plan(0)
        r0 = _to_np(solver.solve_label(dA, db, dc))
    finally:
        ctx.set_solve_plan(-1)
    for k in ('status', 'labels', 'n_active', 'ties', 'violations', 'pivots'):
        assert (r7[k] == r0[k]).all(), k
    assert set(np.unique(r7['status'])) <= {2, 3, 5}
    ok = r7['status'] == 2
    if ok.any():
        assert np.abs(r7['x'][ok] - r0['x'][ok]).max() <= 1e-9 * np.abs(r0['x'][ok]).max()
        assert np.abs(r7['obj'][ok] - r0['obj'][ok]).max() <= 1e-9 * np.abs(r0['obj'][ok]).max()
    # reduced LPs (row masks): random masks that keep ~85 % of the rows, one instance with fewer kept rows than columns
    # (flagged for the generic kernel by both)
    if m > n + 8:
        gen = torch.Generator(device=dA.device).manual_seed(m * 1000 + n)
        mask = (torch.rand(N, m, device=dA.device, generator=gen) < 0.85).to(torch.uint8)
        mask[0, n - 3:] = 0
        try:
            ctx.set_solve_plan(7)
            q7 = _to_np(solver.solve_label(dA, db, dc, row_mask=mask))
            ctx.set_solve_plan(0)
            q0 = _to_np(solver.solve_label(dA, db, dc, row_mask=mask))
        finally:
            ctx.set_solve_plan(-1)
        for k in ('status', 'labels', 'n_active', 'ties', 'violations', 'pivots'):
            assert (q7[k] == q0[k]).all(), ('mask', k)
        okm = q7['status'] == 2
        if okm.any():
            assert np.abs(q7['x'][okm] - q0['x'][okm]).max() <= 1e-9 * np.abs(q0['x'][okm]).max()


def test_plans_agree_bit_for_bit(cuda_device):
    """shared-memory and global-memory tableau kernels run the same arithmetic."""
    from deep_dantzig_b200 import solver, _lib
    ctx = _lib.context(0)
    A, b, c = _numpy_batch(60, 24, list(range(200)))
    dA, db, dc = _dev(A, b, c)
    try:
        ctx.set_solve_plan(1)
        r1 = _to_np(solver.solve_label(dA, db, dc))
        ctx.set_solve_plan(2)
        r2 = _to_np(solver.solve_label(dA, db, dc))
    finally:
        ctx.set_solve_plan(-1)
    for k in ('status', 'labels', 'pivots', 'n_active'):
        assert (r1[k] == r2[k]).all(), k
    ok = r1['status'] == 2
    assert (r1['x'][ok] == r2['x'][ok]).all() and (r1['obj'][ok] == r2['obj'][ok]).all()


@pytest.mark.parametrize('density', [1.0, 0.1])
def test_streamed_global_tableau_agrees_bit_for_bit(cuda_device, density):
    """(200,100) is a shape where the global-memory plan runs its 512-thread streaming instantiation (m n > 16384)
    while the shared-memory plan still fits: same arithmetic, bit for bit -- dense instances and sparse ones (entries
    of the entering column that are exactly zero: rows the update skips), Philox batches of 600 so that persistent
    CTAs run several instances each.  With DDB_PLAN2_RING=1|2|3 in the environment the same test covers the bulk-TMA
    row-ring variant of that instantiation (even n), including the requested-ahead rows it discards (recorded green
    in gpurun_out/pytest_solve_ring.log / profiles/plan2_tma_ring_r01.txt)."""
    from deep_dantzig_b200 import solver, _lib
    ctx = _lib.context(0)
    dA, db, dc = solver.generate(91, 0, 600, 200, 100, density=density)
    try:
        ctx.set_solve_plan(1)
        r1 = _to_np(solver.solve_label(dA, db, dc))
        ctx.set_solve_plan(2)
        r2 = _to_np(solver.solve_label(dA, db, dc))
    finally:
        ctx.set_solve_plan(-1)
    assert (r1['status'] == 2).sum() > 100
    for k in ('status', 'labels', 'pivots', 'n_active', 'ties'):
        assert (r1[k] == r2[k]).all(), k
    ok = r1['status'] == 2
    assert (r1['x'][ok] == r2['x'][ok]).all() and (r1['obj'][ok] == r2['obj'][ok]).all()


def test_host_and_device_flavours_agree(cuda_device):
    from deep_dantzig_b200 import solver
    A, b, c = _numpy_batch(50, 20, list(range(3000)))            # > one host chunk boundary is exercised below
    rh = solver.solve_label_host(A, b, c)
    rd = _to_np(solver.solve_label(*_dev(A, b, c)))
    for k in ('status', 'labels', 'pivots', 'n_active', 'ties'):
        assert (rh[k] == rd[k]).all(), k
    ok = rh['status'] == 2
    assert (rh['x'][ok] == rd['x'][ok]).all()


def test_degeneracies_and_infeasibility_at_the_column_block_kernel_shape(cuda_device):
    """The hand-made cases of the test above at a shape the column-block-per-warp kernel takes automatically (170 x 80), plus
    infeasible instances (a row and its negation with contradictory right-hand sides: phase 1 must prove status 3) -- against
    the polished oracle."""
    from deep_dantzig_b200 import solver, _lib
    m, n, N = 170, 80, 120
    assert _lib.context(0).solve_plan(m, n) == 7
    A, b, c = _numpy_batch(m, n, [577 * i + 11 for i in range(N)])
    rng = np.random.RandomState(8)
    infeasible = np.zeros(N, bool)
    for i in range(N):
        kind = i % 5
        score = (A[i] @ c[i]) / np.linalg.norm(A[i], axis=1)
        first = int(np.argsort(score)[0])
        if kind == 0:
            z = int(rng.randint(m)); A[i, z] = 0.0; b[i, z] = abs(b[i, z]) + 0.1
        elif kind == 1:
            src, dst = rng.choice(m, 2, replace=False); A[i, dst] = A[i, src]; b[i, dst] = b[i, src]
        elif kind == 2:
            A[i, first, int(np.abs(A[i, first]).argmax())] = 0.0
        elif kind == 3:
            keep = rng.choice(n, 2, replace=False); row = np.zeros(n); row[keep] = A[i, first, keep]; A[i, first] = row
        else:                                                   # a_s x <= b_s and -a_s x <= -b_s - 1: empty feasible set
            src, dst = rng.choice(m, 2, replace=False); A[i, dst] = -A[i, src]; b[i, dst] = -b[i, src] - 1.0
            infeasible[i] = True
    r = solver.solve_label_host(A, b, c)
    ref = oracle.solve_batch(A, b, c)
    ok = ref['status'] == 2
    assert ((r['status'] == 2) == ok).all()
    assert (r['status'][infeasible] == 3).all() and (ref['status'][infeasible] != 2).all()
    assert (r['labels'][~ok] == 0).all()
    assert (r['labels'][ok] == ref['labels'][ok]).all()
    assert (r['n_active'][ok] == ref['n_active'][ok]).all()
    if ok.any():
        assert np.abs(r['obj'][ok] - ref['obj'][ok]).max() <= REL_TOL * np.abs(ref['obj'][ok]).max()
        assert np.abs(r['x'][ok] - ref['x'][ok]).max() <= REL_TOL * np.abs(ref['x'][ok]).max()


def test_edge_cases(cuda_device):
    from deep_dantzig_b200 import solver, _lib
    # empty batch
    r = solver.solve_label_host(np.zeros((0, 5, 3)), np.zeros((0, 5)), np.zeros((0, 3)))
    assert r['status'].shape == (0,)
    # fewer rows than columns: no vertex, generic instance is unbounded
    A, b, c = _numpy_batch(3, 5, [0, 1])
    r = solver.solve_label_host(A, b, c)
    ref = oracle.solve_batch(A, b, c)
    assert ((r['status'] == 2) == (ref['status'] == 2)).all() and (r['labels'] == 0).all()
    # a bounded box: min -x-y s.t. x<=1, y<=1, -x-y<=1  -> x=(1,1), rows 0,1 active
    A = np.array([[[1.0, 0], [0, 1.0], [-1, -1]]]); b = np.array([[1.0, 1, 1]]); c = np.array([[-1.0, -1]])
    r = solver.solve_label_host(A, b, c)
    assert r['status'][0] == 2 and list(r['labels'][0]) == [1, 1, 0]
    assert r['x'][0] == pytest.approx([1.0, 1.0]) and r['obj'][0] == pytest.approx(-2.0)
    # infeasible: x <= -1 and -x <= -1 (x >= 1)
    A = np.array([[[1.0], [-1.0]]]); b = np.array([[-1.0, -1.0]]); c = np.array([[1.0]])
    r = solver.solve_label_host(A, b, c)
    assert r['status'][0] == 3 and (r['labels'] == 0).all()
    # degenerate vertex (three lines through one point in 2-D): still optimal, tie reported through n_active > n
    A = np.array([[[1.0, 0], [0, 1.0], [1.0, 1.0], [-1.0, 0], [0, -1.0]]]); b = np.array([[1.0, 1, 2, 5, 5]]); c = np.array([[-1.0, -1]])
    r = solver.solve_label_host(A, b, c)
    assert r['status'][0] == 2 and list(r['labels'][0]) == [1, 1, 1, 0, 0] and r['n_active'][0] == 3 and r['ties'][0] >= 1
    # bad arguments fail loudly
    with pytest.raises(ValueError):
        solver.solve_label_host(np.zeros((2, 4, 2)), np.zeros((2, 3)), np.zeros((2, 2)))
    with pytest.raises(_lib.DdbError):
        _lib.context(0).solve_plan(10, 600)


def test_reduced_lp_row_mask(cuda_device):
    """Rows left out by a mask that keeps every active row do not change the optimum (config-4 reduced solve)."""
    from deep_dantzig_b200 import solver
    A, b, c = _numpy_batch(50, 20, list(range(64)))
    full = solver.solve_label_host(A, b, c)
    rng = np.random.RandomState(0)
    mask = np.maximum(full['labels'], (rng.rand(64, 50) < 0.3).astype(np.uint8))
    ok = full['status'] == 2
    red = solver.solve_label_host(A[ok], b[ok], c[ok], row_mask=mask[ok])
    assert (red['status'] == 2).all() and (red['labels'] == full['labels'][ok]).all()
    assert (red['violations'] == 0).all()
    # oracle leg: HiGHS (polished) on the kept rows A[mask] only, labels / violations over all rows at that optimum
    ored = oracle.solve_batch(A[ok], b[ok], c[ok], row_mask=mask[ok])
    assert (ored['status'] == 2).all() and (red['labels'] == ored['labels']).all() and (ored['violations'] == 0).all()
    assert np.abs(red['x'] - ored['x']).max() <= REL_TOL * np.abs(ored['x']).max()
    assert np.abs(red['obj'] - ored['obj']).max() <= REL_TOL * np.abs(ored['obj']).max()
    assert np.abs(red['obj'] - full['obj'][ok]).max() <= 1e-9 * np.abs(full['obj'][ok]).max()
    assert red['pivots'][:, 3].mean() < full['pivots'][ok, 3].mean()
    # dropping an active row is detected: the reduced optimum violates it (or the reduced LP is unbounded)
    bad = mask[ok].copy()
    first_active = full['labels'][ok].argmax(axis=1)
    bad[np.arange(bad.shape[0]), first_active] = 0
    r2 = solver.solve_label_host(A[ok], b[ok], c[ok], row_mask=bad)
    assert ((r2['status'] != 2) | (r2['violations'] > 0)).all()
    # ... and the kernel's reduced solve still equals the oracle's solve of the same (wrong) reduced LP, instance by instance
    o2 = oracle.solve_batch(A[ok], b[ok], c[ok], row_mask=bad)
    assert ((r2['status'] == 2) == (o2['status'] == 2)).all()
    both = (r2['status'] == 2)
    assert (r2['labels'][both] == o2['labels'][both]).all()
    assert (r2['violations'][both] == o2['violations'][both]).all()
    assert np.abs(r2['x'][both] - o2['x'][both]).max() <= REL_TOL * np.abs(o2['x'][both]).max()


def test_reduced_lp_row_mask_at_config4_shape(cuda_device):
    """The reduced solve at BASELINE.json configs[3] shape (500,250) against the oracle on A[mask]: labels of the full LP
    plus 20 % random rows kept."""
    from deep_dantzig_b200 import solver
    g_seeds = [685 * i for i in range(24)]
    A, b, c = _numpy_batch(500, 250, g_seeds)
    full = oracle.solve_batch_parallel(A, b, c)
    ok = full['status'] == 2
    rng = np.random.RandomState(2)
    mask = np.maximum(full['labels'], (rng.rand(len(g_seeds), 500) < 0.2).astype(np.uint8))
    red = solver.solve_label_host(A[ok], b[ok], c[ok], row_mask=mask[ok])
    ored = oracle.solve_batch_parallel(A[ok], b[ok], c[ok], row_mask=mask[ok])
    assert (red['status'] == 2).all() and (ored['status'] == 2).all()
    assert (red['labels'] == ored['labels']).all() and (red['labels'] == full['labels'][ok]).all()
    assert (red['violations'] == 0).all()
    assert np.abs(red['x'] - ored['x']).max() <= REL_TOL * np.abs(ored['x']).max()


def test_philox_generator_matches_cpu_restatement(cuda_device):
    from deep_dantzig_b200 import solver
    for (m, n, dens) in [(50, 20, 1.0), (33, 17, 1.0), (64, 32, 0.3)]:
        A, b, c, x0 = solver.generate(99, 5, 4, m, n, density=dens, want_x0=True)
        A, b, c, x0 = [t.cpu().numpy() for t in (A, b, c, x0)]
        for k in range(4):
            Ar, br, cr, x0r = philox.generate_instance(99, 5 + k, m, n, dens)
            assert ((A[k] == 0) == (Ar == 0)).all()
            np.testing.assert_allclose(A[k], Ar, rtol=1e-12, atol=1e-14)   # libm log/sincos differ by a few ulp
            np.testing.assert_allclose(x0[k], x0r, rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(c[k], cr, rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(b[k], br, rtol=1e-11, atol=1e-12)
    # chunk-independence: instance i is a function of (key, i) only
    A1, _, _ = solver.generate(7, 0, 8, 20, 10)
    A2, _, _ = solver.generate(7, 4, 4, 20, 10)
    assert (A1[4:] == A2).all()


def test_fused_generate_solve_label_vs_oracle_on_downloaded_instances(cuda_device):
    """Throughput mode: the oracle consumes the instances the device generated."""
    from deep_dantzig_b200 import solver
    r = solver.generate_solve_label(2024, 0, 300, 200, 100, keep_instances=True)
    A, b, c = r['A'].cpu().numpy(), r['b'].cpu().numpy(), r['c'].cpu().numpy()
    ref = oracle.solve_batch(A[:120], b[:120], c[:120])
    rn = {k: v[:120] for k, v in _to_np(r).items()}
    _check_against_oracle(rn, ref, A[:120], b[:120], c[:120])
    # and the non-materialising call returns the same answers
    r2 = _to_np(solver.generate_solve_label(2024, 0, 300, 200, 100))
    assert (r2['status'] == r['status'].cpu().numpy()).all() and (r2['labels'] == r['labels'].cpu().numpy()).all()


def test_full_size_properties(cuda_device):
    """BASELINE.json configs[1] shape at scale, through size-independent properties (no oracle at this size)."""
    from deep_dantzig_b200 import solver
    B, m, n = 20000, 200, 100
    r = solver.generate_solve_label(11, 0, B, m, n, keep_instances=True)
    st = r['status'].cpu().numpy(); lab = r['labels']; na = r['n_active'].cpu().numpy()
    assert set(np.unique(st)) <= {2, 5}
    frac_unb = (st == 5).mean()
    assert abs(frac_unb - 0.472) < 0.02, frac_unb              # Wendel: 2^-m sum_{k<n} C(m,k) at m = 2n -> ~0.472
    ok = st == 2
    # ties (a slack inside [1e-8, 1e-6], or a thresholded label that disagrees with the final basis -- an
    # ill-conditioned vertex) are rare but legitimate at this scale: counted, reported, bounded, never hidden
    tie = r['ties'].cpu().numpy() > 0
    n_tie = int(tie.sum())
    print('instances with a reported tie: %d of %d' % (n_tie, B))
    assert n_tie <= B // 1000
    assert (na[ok & ~tie] == n).all() and (na[~ok] == 0).all()   # non-degenerate: exactly n active rows
    # ill-conditioned vertices (|x| >> 1) get a step of iterative refinement on the device: at this scale every optimal
    # instance ends with exactly n thresholded labels (instance 16738 of this stream had 98 without it)
    assert (na[ok] == n).all()
    # KKT certificate on the device data, independent of the solver: primal feasibility and objective consistency
    okt = torch.from_numpy(ok & ~tie).cuda()
    A, b, c, x = r['A'][okt], r['b'][okt], r['c'][okt], r['x'][okt]
    slack = b - torch.bmm(A, x.unsqueeze(2)).squeeze(2)
    assert slack.min().item() >= -1e-7
    assert ((slack.abs() <= 1e-7) == lab[okt].bool()).all()
    assert torch.allclose((c * x).sum(1), r['obj'][okt], rtol=1e-12, atol=0)
    # full optimality certificate on EVERY optimal instance, independent of any solver: with B = the labelled rows,
    #   dual:   c = -A_B' y with y >= 0            primal: x is the vertex A_B x = b_B (and feasible, above)
    K = int(okt.sum().item())
    rows = lab[okt].bool().nonzero()[:, 1].reshape(K, n)                       # exactly n labelled rows per instance
    AB = torch.gather(A, 1, rows.unsqueeze(2).expand(K, n, n))
    bB = torch.gather(b, 1, rows)
    y = torch.linalg.solve(AB.transpose(1, 2), -c.unsqueeze(2)).squeeze(2)
    assert y.min().item() >= -1e-7
    xv = torch.linalg.solve(AB, bB.unsqueeze(2)).squeeze(2)
    xv = xv + torch.linalg.solve(AB, (bB - torch.bmm(AB, xv.unsqueeze(2)).squeeze(2)).unsqueeze(2)).squeeze(2)   # one refinement step
    relv = ((x - xv).abs().amax(1) / xv.abs().amax(1))
    print('optimal instances %d: max relative distance of x from the vertex of its labelled rows %.2e' % (K, relv.max().item()))
    assert relv.max().item() <= 1e-9
    # linearity: scaling the objective scales the optimum, labels unchanged; row permutation permutes labels
    r2 = solver.solve_label(r['A'][:512], r['b'][:512], (r['c'][:512] * 3.0).contiguous())
    assert (r2['labels'] == lab[:512]).all()
    okh = torch.from_numpy(ok[:512]).cuda()
    assert torch.allclose(r2['obj'][okh], 3.0 * r['obj'][:512][okh], rtol=1e-9, atol=0)
    perm = torch.randperm(m, device='cuda')
    r3 = solver.solve_label(r['A'][:512][:, perm].contiguous(), r['b'][:512][:, perm].contiguous(), r['c'][:512])
    assert (r3['labels'] == lab[:512][:, perm]).all()


def test_dropin_dataset_and_linprog(cuda_device):
    """The reference-facing classes: same names, arguments, item layout and error behaviour."""
    from deep_dantzig_b200.data.randomlp_dataset import RandomLPDataset
    from deep_dantzig_b200.data.gurobi_lp import LinProg
    ds = RandomLPDataset(10, 5, num_lps=6, seed=0)
    ods = oracle.RandomLPDataset(10, 5, num_lps=6, seed=0)
    assert len(ds) == 6
    for i in range(7):
        a, o_ = ds[i], ods[i]
        assert (a['lp']['A'] == o_['lp']['A']).all() and (a['lp']['b'] == o_['lp']['b']).all()
        assert a['labels'] == o_['labels']
    for s, so in zip(ds.get_lp_params(), ods.get_lp_params()):
        assert s['id'] == so['id'] and s['success'] == so['success'] and s['active'] == so['active']
        assert (s['sc'] in (1, 2)) == (so['sc'] in (1, 2))
        if s['success']:
            assert s['objval'] == pytest.approx(so['objval'], rel=1e-9)
    p = RandomLPDataset.create_lp_problem(10, 5, 0, with_stats=True)     # the reference's own main() case
    assert list(p['active']) == [0, 5, 6, 8, 9] and p['stats']['objval'] == pytest.approx(-2.56554105335413, rel=1e-9)
    A, b, c = oracle.generate_instance(10, 5, 3)
    lp = LinProg(A, b, c, 'min', ['<'] * 10)
    assert lp.get_statuscode() == 1
    lp.optimize()
    assert lp.get_statuscode() == 2 and list(lp.get_active_constraints()) == [0, 3, 4, 6, 9]
    assert lp.model.objVal == pytest.approx(2.368364743974648, rel=1e-9) and lp.model.status == 2
    lp = LinProg(*oracle.generate_instance(10, 5, 1))
    lp.optimize()
    assert lp.get_statuscode() not in (1, 2)
    with pytest.raises(AttributeError):
        lp.model.objVal
    with pytest.raises(ValueError):
        LinProg(A, b, c, 'sideways')
    # 'max' and '>' are sign flips of the canonical form
    lp = LinProg(-A, -b, -c, 'max', ['>'] * 10)
    lp.optimize()
    assert list(lp.get_active_constraints()) == [0, 3, 4, 6, 9]
    ph = RandomLPDataset(50, 20, num_lps=32, seed=5, generator='philox')
    assert len(ph) == 32 and ph[0]['lp']['A'].shape == (50, 20)


_EXPERIMENTS = os.environ.get('DDB_EXPERIMENTS') == '1'      # `make -C deep_dantzig_b200/csrc experiments` build loaded


@pytest.mark.parametrize('plan0', [0, 4, 7] + ([3, 5] if _EXPERIMENTS else []))
@pytest.mark.parametrize('m,n,N', [(10, 5, 300), (50, 20, 300), (100, 50, 100), (200, 100, 200)])
def test_register_resident_and_generic_kernels_agree(cuda_device, m, n, N, plan0):
    """The register-resident kernels (0: row per thread -- hybrid register + shared-memory rows at (200,100) --, 4: warp-tiled,
    7: rows over lanes and columns over warps, the default from n = 72;
    with DDB_EXPERIMENTS=1 also the measured negative results 3: 2-D register tile, 5: software-pipelined rows) and plan 1
    (tableau in shared memory) implement the same algorithm: same statuses, labels and pivot path."""
    from deep_dantzig_b200 import solver, _lib
    ctx = _lib.context(0)
    assert ctx.solve_plan(m, n) == (7 if n >= 72 else 0)
    A, b, c = _numpy_batch(m, n, [17 + 3 * i for i in range(N)])
    dA, db, dc = _dev(A, b, c)
    try:
        ctx.set_solve_plan(plan0)
        r0 = _to_np(solver.solve_label(dA, db, dc))
        ctx.set_solve_plan(1)
        r1 = _to_np(solver.solve_label(dA, db, dc))
    finally:
        ctx.set_solve_plan(-1)
    assert (r0['status'] == r1['status']).all()
    assert (r0['labels'] == r1['labels']).all()
    assert (r0['n_active'] == r1['n_active']).all()
    assert (r0['pivots'] == r1['pivots']).all()
    ok = r0['status'] == 2
    assert np.abs(r0['x'][ok] - r1['x'][ok]).max() <= 1e-9 * np.abs(r1['x'][ok]).max()
    assert np.abs(r0['obj'][ok] - r1['obj'][ok]).max() <= 1e-9 * np.abs(r1['obj'][ok]).max()


@pytest.mark.parametrize('m,n,B,dens', [(400, 100, 300, 1.0), (250, 100, 400, 1.0), (300, 150, 300, 1.0), (400, 150, 200, 0.1),
                                        (484, 100, 200, 0.5)])
def test_in_kernel_generator_on_the_wide_row_variants(cuda_device, m, n, B, dens):
    """Fused mode 1 (the instance drawn inside the solver CTA) on the six- / twelve-warp and 151-column row variants: the
    same instance bits as the generator entry point and the same results as solving the materialised instances
    (tools/check_gen_variants.py is the same check as a script)."""
    from deep_dantzig_b200 import solver, _lib
    ctx = _lib.context(0)
    A, b, c = solver.generate(61, 7, B, m, n, density=dens)
    want = solver.solve_label(A, b, c)
    ctx.set_fused_mode(1)
    try:
        got = solver.generate_solve_label(61, 7, B, m, n, density=dens)
        keep = solver.generate_solve_label(61, 7, B, m, n, density=dens, keep_instances=True)
    finally:
        ctx.set_fused_mode(0)
    assert (keep['A'] == A).all() and (keep['b'] == b).all() and (keep['c'] == c).all()
    for k in ('status', 'labels', 'pivots', 'n_active', 'ties', 'x'):
        assert (got[k] == want[k]).all() and (keep[k] == want[k]).all(), k


def test_fused_in_kernel_generator_is_bit_identical(cuda_device):
    """The fused call draws each instance inside the solver CTA.  Same (key, index) -> the same instance bits as the
    generator entry point, and -- same kernel arithmetic on the same bits -- the same results bit for bit as solving the
    materialised instances; chunk- and rank-independent (first_instance offsets)."""
    from deep_dantzig_b200 import solver, _lib
    ctx = _lib.context(0)
    try:
        for mode in (1, 2):           # 1: in-kernel generation (one launch), 2: generator kernel + solver kernel (automatic choice)
            ctx.set_fused_mode(mode)
            for (m, n, B, dens) in [(200, 100, 700, 1.0), (50, 20, 3000, 1.0), (150, 100, 300, 0.5), (64, 32, 500, 1.0),
                                    (125, 100, 900, 0.1), (75, 20, 1000, 1.0)]:
                A, b, c = solver.generate(5150, 100, B, m, n, density=dens)
                want = _to_np(solver.solve_label(A, b, c))
                keep = solver.generate_solve_label(5150, 100, B, m, n, density=dens, keep_instances=True)
                assert (keep['A'] == A).all() and (keep['b'] == b).all() and (keep['c'] == c).all()
                lean = _to_np(solver.generate_solve_label(5150, 100, B, m, n, density=dens))
                host = solver.generate_solve_label_host(5150, 100, B, m, n, density=dens)
                tail = _to_np(solver.generate_solve_label(5150, 100 + B // 2, B - B // 2, m, n, density=dens))
                for k in ('status', 'labels', 'pivots', 'n_active', 'ties', 'violations', 'x', 'obj'):
                    for name, got in (('keep', _to_np(keep)), ('lean', lean), ('host', host)):
                        eq = (got[k] == want[k]) | ((got[k] != got[k]) & (want[k] != want[k])) if k == 'obj' else (got[k] == want[k])
                        assert eq.all(), (mode, m, n, k, name)
                    eq = (tail[k] == want[k][B // 2:]) | ((tail[k] != tail[k]) if k == 'obj' else False)
                    assert np.all(eq), (mode, m, n, k, 'tail')
    finally:
        ctx.set_fused_mode(0)
        ctx.set_solve_plan(-1)
    # odd n: the two-kernel fallback, same contract
    r = _to_np(solver.generate_solve_label(9, 0, 200, 33, 17))
    A, b, c = solver.generate(9, 0, 200, 33, 17)
    w = _to_np(solver.solve_label(A, b, c))
    assert (r['status'] == w['status']).all() and (r['labels'] == w['labels']).all()


def test_host_entry_with_pageable_and_pinned_buffers(cuda_device):
    """ddb_solve_label_host stages pageable caller memory through its pinned ring and DMAs pinned memory in place: same
    results either way, across several chunks (DDB_HOST_CHUNK_MB-sized) and three slots."""
    from deep_dantzig_b200 import solver
    A, b, c = _numpy_batch(50, 20, list(range(40000)))       # 327 MB of instances: two chunks
    want = _to_np(solver.solve_label(*_dev(A[:6000], b[:6000], c[:6000])))
    page = solver.solve_label_host(A, b, c)
    pA, pb, pc = [torch.from_numpy(a_).pin_memory().numpy() for a_ in (A, b, c)]
    pin = solver.solve_label_host(pA, pb, pc, out=solver._host_outputs(40000, 50, 20, True))
    for k in ('status', 'labels', 'pivots', 'n_active', 'ties', 'violations'):
        assert (page[k] == pin[k]).all(), k
        assert (page[k][:6000] == want[k]).all(), k
    ok = page['status'] == 2
    assert (page['x'][ok] == pin['x'][ok]).all() and (page['obj'][ok] == pin['obj'][ok]).all()


def test_singular_crash_basis_is_handed_to_the_generic_kernel(cuda_device):
    """A duplicated top-ranked row makes the static crash basis singular: plan 0 flags the instance and the
    generic kernel re-solves it on the device; results still match the oracle."""
    from deep_dantzig_b200 import solver
    A, b, c = _numpy_batch(50, 20, list(range(40)))
    for i in range(0, 40, 2):
        score = (A[i] @ c[i]) / np.linalg.norm(A[i], axis=1)
        r0, r1 = np.argsort(score)[:2]
        A[i, r1] = A[i, r0]; b[i, r1] = b[i, r0] + 0.5          # parallel, never-active copy inside the crash basis
    r = solver.solve_label_host(A, b, c)
    ref = oracle.solve_batch(A, b, c)
    assert ((r['status'] == 2) == (ref['status'] == 2)).all()
    ok = ref['status'] == 2
    assert (r['labels'][ok] == ref['labels'][ok]).all()
    assert np.abs(r['obj'][ok] - ref['obj'][ok]).max() <= 1e-9 * np.abs(ref['obj'][ok]).max()


def test_linprog_ops_and_objective_sense(cuda_device):
    """LinProg(A, b, c, obj, ops) with '<', '>', '=' rows and 'min' / 'max' (reference gurobi_lp.py:392-426) against the
    oracle's LinProg (HiGHS with A_eq): status, objective, x and active rows."""
    from deep_dantzig_b200.data.gurobi_lp import LinProg
    # min -x s.t. x + y = 1, y >= 0, x <= 0.7  ->  x = 0.7, y = 0.3; active: the equality and x <= 0.7
    A = np.array([[1.0, 1.0], [0.0, 1.0], [1.0, 0.0]]); b = np.array([1.0, 0.0, 0.7]); c = np.array([-1.0, 0.0])
    lp = LinProg(A, b, c, 'min', ['=', '>', '<']); lp.optimize()
    assert lp.get_statuscode() == 2 and lp.x == pytest.approx([0.7, 0.3]) and lp.model.objVal == pytest.approx(-0.7)
    assert list(lp.get_active_constraints()) == [0, 2]
    lp = LinProg(A, b, -c, 'max', ['=', '>', '<']); lp.optimize()
    assert lp.get_statuscode() == 2 and lp.model.objVal == pytest.approx(0.7) and list(lp.get_active_constraints()) == [0, 2]
    # random instances with mixed senses
    rs = np.random.RandomState(5)
    checked = 0
    for trial in range(40):
        m, n = 14, 5
        A = rs.randn(m, n); x0 = rs.randn(n); cc = np.abs(rs.randn(n))
        ops = ['<'] * m
        slack = np.abs(rs.randn(m))
        for i in rs.choice(m, 4, replace=False):
            ops[i] = '>'
        eqs = rs.choice([i for i in range(m) if ops[i] == '<'], 2, replace=False)
        for i in eqs:
            ops[i] = '='
        bb = A.dot(x0) + np.array([0.0 if ops[i] == '=' else (slack[i] if ops[i] == '<' else -slack[i]) for i in range(m)])
        ours = LinProg(A, bb, cc, 'min', ops); ours.optimize()
        ref = oracle.LinProg(A, bb, cc, 'min', ops); ref.optimize()
        assert (ours.model.status == 2) == (ref.model.status == 2), trial
        if ref.model.status == 2:
            assert ours.model.objVal == pytest.approx(ref.model.objVal, rel=1e-8, abs=1e-9)
            assert list(ours.get_active_constraints()) == list(ref.get_active_constraints()), trial
            checked += 1
    assert checked >= 10
