"""numpy model of the device algorithm (dev tool; NOT the oracle, NOT on the product path).

Primal-orientation condensed tableau for  min c'x, Ax<=b, x free:
    sigma_i = s_i - sum_j P[i][j] nu_j      (row i = constraint whose slack is basic)
    z       = z0  + sum_j g[j]  nu_j
nu_j is a free x_j before the crash, afterwards the slack of an active constraint.
Stages: static-order crash (n Gauss-Jordan pivots, row picked by cosine score, column = max |entry|),
phase 1 (dual-simplex type, artificial costs ghat=1) -> s>=0, phase 2 (Dantzig primal simplex) -> g>=0,
x = xv - D (sigma*_B0) from the frozen crash rows.
"""
import numpy as np

TOL_PIV = 1e-9
TOL_FEAS = 1e-9

def solve(A, b, c, stats=None):
    m, n = A.shape
    P = A.astype(np.float64).copy(); s = b.astype(np.float64).copy(); g = c.astype(np.float64).copy()
    nrm = np.sqrt((A * A).sum(1)); nrm[nrm == 0] = 1
    score = (A @ c) / nrm
    order = np.argsort(score, kind='stable')
    live = np.ones(m, bool)            # row still holds a basic slack
    colvar = -np.ones(n, int)          # constraint whose slack is nonbasic in column j (-1: free x_j)
    rowfree = -np.ones(m, int)         # for crashed rows: which x_j is basic there
    npiv = [0, 0, 0]

    def pivot(r, k, rows):
        p = P[r, k]
        P[r, :] /= p; P[r, k] = 1.0 / p; s[r] /= p
        colk = P[rows, k].copy()
        colk[rows == r] = 0
        P[rows, :] -= np.outer(colk, P[r, :])
        P[rows, k] = np.where(rows == r, P[r, k], -colk / p)
        s[rows] -= colk * s[r]
        return p

    # crash
    allrows = np.arange(m)
    t = 0; oi = 0
    while t < n and oi < m:
        r = order[oi]; oi += 1
        cand = np.where(colvar < 0, np.abs(P[r]), -1.0)
        k = int(np.argmax(cand))
        if cand[k] < 1e-7:
            continue                     # skip (model only; device v1 reports numeric failure)
        gk = g[k]
        p = pivot(r, k, allrows[live | (rowfree >= 0)])
        g -= gk * P[r]; g[k] = -gk / p
        live[r] = False; rowfree[r] = k; colvar[k] = r
        t += 1; npiv[0] += 1
    if t < n:
        return dict(status=12, pivots=npiv)
    colvar0 = colvar.copy()
    D = P[~live & (rowfree >= 0)].copy(); xv = s[~live].copy(); xrow = rowfree[~live].copy()
    rows = allrows[live]
    rowvar = allrows.copy()            # constraint whose slack is basic in row i (live rows)
    gh = np.ones(n)

    def spivot(r, k):
        gk = g[k]; ghk = gh[k]
        p = pivot(r, k, rows)
        g[:] -= gk * P[r]; g[k] = -gk / p
        gh[:] -= ghk * P[r]; gh[k] = -ghk / p
        rowvar[r], colvar[k] = colvar[k], rowvar[r]

    status = 2
    maxit = 50 * (m + n)
    # phase 1
    while True:
        if len(rows) == 0: break
        i = int(np.argmin(s[rows])); r = rows[i]
        if s[r] >= -TOL_FEAS: break
        row = P[r]
        ok = row < -TOL_PIV
        if not ok.any(): status = 3; break
        rat = np.where(ok, np.maximum(gh, 0) / np.where(ok, -row, 1), np.inf)
        k = int(np.argmin(rat))
        spivot(r, k); npiv[1] += 1
        if npiv[1] > maxit: status = 7; break
    # phase 2
    while status == 2:
        k = int(np.argmin(g))
        if g[k] >= -TOL_FEAS: break
        col = P[rows, k]
        ok = col > TOL_PIV
        if not ok.any(): status = 5; break
        rat = np.where(ok, np.maximum(s[rows], 0) / np.where(ok, col, 1), np.inf)
        i = int(np.argmin(rat)); r = rows[i]
        spivot(r, k); npiv[2] += 1
        if npiv[2] > maxit: status = 7; break
    out = dict(status=status, pivots=npiv)
    if status == 2:
        sig = np.zeros(n)
        pos = {rowvar[r]: r for r in rows}
        for j in range(n):
            q = colvar0[j]
            if q in pos: sig[j] = s[pos[q]]
        xb = xv - D @ sig
        x = np.zeros(n); x[xrow] = xb
        out.update(x=x, obj=float(c @ x), basis=np.sort(colvar))
    return out

if __name__ == '__main__':
    import sys
    from scipy.optimize import linprog
    m, n, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    bad = 0; piv = []; errx = []; erro = []; ninf = 0
    for sd in range(N):
        np.random.seed(sd)
        A = np.random.randn(m, n); b = A.dot(np.random.randn(n)) + np.absolute(np.random.randn(m)); c = np.absolute(np.random.randn(n))
        r = linprog(c, A_ub=A, b_ub=b, bounds=(None, None), method='highs-ds')
        o = solve(A, b, c)
        piv.append(o['pivots'])
        if r.status == 0:
            sl = b - A @ r.x; sl[np.abs(sl) <= 1e-7] = 0; act = (sl == 0).nonzero()[0]
            if o['status'] != 2: bad += 1; print('status mismatch', sd, o['status']); continue
            sl2 = b - A @ o['x']; sl2[np.abs(sl2) <= 1e-7] = 0; act2 = (sl2 == 0).nonzero()[0]
            if not (np.array_equal(act, act2) and np.array_equal(act2, o['basis'])): bad += 1; print('active mismatch', sd)
            xe = np.linalg.solve(A[act], b[act])
            errx.append(max(np.abs(o['x'] - r.x).max() / np.abs(r.x).max(), 0)); erro.append(abs(o['obj'] - r.fun) / abs(r.fun))
            errx[-1] = max(errx[-1], np.abs(o['x'] - xe).max() / np.abs(xe).max())
        else:
            ninf += 1
            if o['status'] != 5: bad += 1; print('status mismatch', sd, r.status, o['status'])
    piv = np.array(piv)
    print('N', N, 'bad', bad, 'unbounded', ninf, 'mean pivots crash/p1/p2', piv.mean(0), 'max', piv.max(0))
    if errx: print('max rel err x', max(errx), 'obj', max(erro))
