"""Pivot-count study for the simplex variants considered in DESIGN.md (dev tool, numpy only).

Condensed (Tucker) tableau for the dual  min b'y, (-A')y = c, y>=0  of  min c'x, Ax<=b, x free.
rows = basic y (active constraints), cols = nonbasic y. rhs = y_B, cost row = d_N = primal slacks.
"""
import sys, time
import numpy as np
from scipy.optimize import linprog

def gen(m, n, seed):
    np.random.seed(seed)
    A = np.random.randn(m, n)
    b = A.dot(np.random.randn(n)) + np.absolute(np.random.randn(m))
    c = np.absolute(np.random.randn(n))
    return A, b, c

def ref(A, b, c):
    r = linprog(c, A_ub=A, b_ub=b, bounds=(None, None), method='highs-ds')
    if r.status == 0:
        s = b - A @ r.x
        s[np.abs(s) <= 1e-7] = 0
        return 2, (s == 0).nonzero()[0], r.fun, r.nit
    return (5 if r.status == 3 else r.status + 100), None, None, r.nit

def two_phase_artificial(A, b, c, tol=1e-9):
    """A: textbook; full tableau n x m with artificial basis; returns status, basis, pivots(p1,p2), work entries"""
    m, n = A.shape
    T = np.zeros((n + 2, m + 1))
    T[:n, :m] = -A.T
    T[:n, m] = c
    T[n, :m] = b              # phase-2 cost row (reduced costs as basis of artificials has 0 cost)
    T[n + 1, :m] = -T[:n, :m].sum(0)  # phase-1 cost row
    T[n + 1, m] = -c.sum()
    basis = -np.ones(n, dtype=int)  # -1 artificial
    alive = np.ones(m, dtype=bool)
    piv = [0, 0]
    work = 0
    for phase in (1, 2):
        crow = n + 1 if phase == 1 else n
        while True:
            d = T[crow, :m].copy()
            d[~alive] = np.inf
            for bi in basis:
                if bi >= 0: d[bi] = np.inf
            j = int(np.argmin(d))
            if d[j] >= -tol: break
            col = T[:n, j]
            ratios = np.where(col > tol, T[:n, m] / np.where(col > tol, col, 1), np.inf)
            r = int(np.argmin(ratios))
            if not np.isfinite(ratios[r]):
                return 'dual_unbounded', None, piv, work
            T[r] /= T[r, j]
            for i in range(n + 2):
                if i != r: T[i] -= T[i, j] * T[r]
            basis[r] = j
            piv[phase - 1] += 1
            work += (n + 2) * (m + 1)
        if phase == 1:
            if -T[n + 1, m] > 1e-7:
                return 5, None, piv, work
            # drive out remaining artificials (degenerate) - ignore (prob 0)
            if (basis < 0).any():
                return 'art_in_basis', None, piv, work
    return 2, np.sort(basis), piv, work

def crash(A, b, c, order=None):
    """Gauss-Jordan n pivots -> condensed tableau. Returns T (n x (m-n)), rhs y_B (n), d (m-n), basis idx (n), nonbasis idx."""
    m, n = A.shape
    M = np.zeros((n + 1, m + 1))
    M[:n, :m] = -A.T
    M[:n, m] = c
    M[n, :m] = b
    if order is None: order = np.arange(m)
    rows_left = np.ones(n, dtype=bool)
    basis = -np.ones(n, dtype=int)
    k = 0
    for j in order:
        col = np.abs(M[:n, j]) * rows_left
        r = int(np.argmax(col))
        if col[r] < 1e-3: continue
        M[r] /= M[r, j]
        for i in range(n + 1):
            if i != r: M[i] -= M[i, j] * M[r]
        rows_left[r] = False
        basis[r] = j
        k += 1
        if k == n: break
    nb = np.array([j for j in range(m) if j not in set(basis)])
    return M[:n][:, nb].copy(), M[:n, m].copy(), M[n, nb].copy(), basis, nb, M[n, m]

def pivot(T, y, d, basis, nb, r, k):
    """condensed pivot: basic row r leaves, nonbasic col k enters."""
    p = T[r, k]
    rowr = T[r].copy(); colk = T[:, k].copy()
    yr = y[r]; dk = d[k]
    T -= np.outer(colk, rowr) / p
    T[r] = rowr / p
    T[:, k] = -colk / p
    T[r, k] = 1.0 / p
    y -= colk * (yr / p); y[r] = yr / p
    d -= dk * rowr / p; d[k] = -dk / p
    basis[r], nb[k] = nb[k], basis[r]

def primal_on_dual(T, y, d, basis, nb, tol=1e-9, maxit=100000):
    """phase 2 primal simplex on dual LP: needs y>=0; fix d<0."""
    it = 0
    while it < maxit:
        k = int(np.argmin(d))
        if d[k] >= -tol: return 2, it
        col = T[:, k]
        ratios = np.where(col > tol, y / np.where(col > tol, col, 1), np.inf)
        r = int(np.argmin(ratios))
        if not np.isfinite(ratios[r]): return 3, it   # dual LP unbounded -> primal infeasible
        pivot(T, y, d, basis, nb, r, k); it += 1
    return 7, it

def dual_on_dual(T, y, d, basis, nb, tol=1e-9, maxit=100000):
    """dual simplex on the dual LP (== primal simplex on original): needs d>=0; fix y<0."""
    it = 0
    while it < maxit:
        r = int(np.argmin(y))
        if y[r] >= -tol: return 2, it
        row = T[r]
        ratios = np.where(row < -tol, d / np.where(row < -tol, -row, 1), np.inf)
        k = int(np.argmin(ratios))
        if not np.isfinite(ratios[k]): return 5, it   # dual LP infeasible -> primal unbounded
        pivot(T, y, d, basis, nb, r, k); it += 1
    return 7, it

def composite_phase1_y(T, y, d, basis, nb, tol=1e-9, maxit=100000):
    """make y>=0 by minimising sum of infeasibilities (textbook composite, recompute w each it)."""
    it = 0
    while it < maxit:
        inf = y < -tol
        if not inf.any(): return 2, it
        w = T[inf].sum(0)       # increasing nonbasic k changes y_i by -T[i,k]; infeas sum decreases if sum T[inf,k] < 0
        k = int(np.argmin(w))
        if w[k] >= -tol: return 5, it   # infeasible dual LP -> primal unbounded
        col = T[:, k]
        # ratio: feasible rows with col>0 block at y/col ; infeasible rows with col<0 reach zero at y/col (positive) - let them pass (textbook: block at last)
        rat = np.full(len(y), np.inf)
        f = (~inf) & (col > tol); rat[f] = y[f] / col[f]
        r = int(np.argmin(rat))
        if not np.isfinite(rat[r]):
            g = inf & (col < -tol); rat[g] = y[g] / col[g]
            r = int(np.argmax(np.where(np.isfinite(rat), rat, -1)))
        pivot(T, y, d, basis, nb, r, k); it += 1
    return 7, it

def composite_phase1_d(T, y, d, basis, nb, tol=1e-9, maxit=100000):
    """make d>=0 (primal feasible) by the mirrored composite on columns."""
    it = 0
    while it < maxit:
        inf = d < -tol
        if not inf.any(): return 2, it
        w = T[:, inf].sum(1)   # pivot in row r: d_k' = d_k - d_kk*T[r,k]/p ...
        # mirror: treat transposed negative tableau. leaving row r chosen by most positive w
        r = int(np.argmax(w))
        if w[r] <= tol: return 3, it
        row = T[r]
        rat = np.full(len(d), np.inf)
        f = (~inf) & (row < -tol); rat[f] = d[f] / -row[f]
        k = int(np.argmin(rat))
        if not np.isfinite(rat[k]):
            g = inf & (row > tol); rat[g] = d[g] / -row[g]
            k = int(np.argmax(np.where(np.isfinite(rat), rat, -1)))
        pivot(T, y, d, basis, nb, r, k); it += 1
    return 7, it

if __name__ == '__main__':
    m, n, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    strat = sys.argv[4]
    tot = {}
    for s in range(N):
        A, b, c = gen(m, n, s)
        st, act, obj, nit = ref(A, b, c)
        if strat == 'A':
            st2, basis, piv, work = two_phase_artificial(A, b, c)
            ok = (st2 == st) and (st != 2 or np.array_equal(basis, act))
            print(s, st, st2, piv, nit, ok)
        else:
            order = None
            if 'h' in strat:
                score = (A @ c) / np.linalg.norm(A, axis=1)   # want -a_i.c > 0 -> small score first
                if 'b' in strat:
                    score = score + 0.5 * b / np.linalg.norm(A, axis=1)
                order = np.argsort(score)
            T, y, d, basis, nb, z = crash(A, b, c, order)
            ninf_y = (y < -1e-9).sum(); ninf_d = (d < -1e-9).sum()
            if 'P' in strat:   # primal feasible first, then dual simplex
                s1, it1 = composite_phase1_d(T, y, d, basis, nb)
                s2, it2 = dual_on_dual(T, y, d, basis, nb) if s1 == 2 else (s1, 0)
            else:
                s1, it1 = composite_phase1_y(T, y, d, basis, nb)
                s2, it2 = primal_on_dual(T, y, d, basis, nb) if s1 == 2 else (s1, 0)
            ok = (s2 == st) and (st != 2 or np.array_equal(np.sort(basis), act))
            print(s, st, s2, 'inf_y', ninf_y, 'inf_d', ninf_d, 'it', it1, it2, 'highs', nit, ok)

def self_dual(T, y, d, basis, nb, tol=1e-9, maxit=100000):
    """parametric self-dual simplex (Vanderbei ch.7) on the condensed tableau.
    y + mu*yb >= 0, d + mu*db >= 0 ; yb, db start at 1."""
    n, q = T.shape
    yb = np.ones(n); db = np.ones(q)
    it = 0; nrow = ncol = 0
    while it < maxit:
        # mu* = smallest mu keeping everything nonneg
        mr = np.where(yb > tol, -y / np.where(yb > tol, yb, 1), -np.inf)
        mc = np.where(db > tol, -d / np.where(db > tol, db, 1), -np.inf)
        # entries with yb<=tol and y<-tol: infeasible regardless (shouldn't happen)
        r = int(np.argmax(mr)); k = int(np.argmax(mc))
        mu = max(mr[r], mc[k])
        if mu <= tol: return 2, it, nrow, ncol
        if mr[r] >= mc[k]:
            # row r leaves (dual-type pivot): ratio over row entries T[r,k]<0 of (d+mu*db)/-T
            row = T[r]
            num = d + mu * db
            rat = np.where(row < -tol, num / np.where(row < -tol, -row, 1), np.inf)
            k = int(np.argmin(rat))
            if not np.isfinite(rat[k]): return 5, it, nrow, ncol
            nrow += 1
        else:
            col = T[:, k]
            num = y + mu * yb
            rat = np.where(col > tol, num / np.where(col > tol, col, 1), np.inf)
            r = int(np.argmin(rat))
            if not np.isfinite(rat[r]): return 3, it, nrow, ncol
            ncol += 1
        # pivot incl. perturbation vectors
        p = T[r, k]; colk = T[:, k].copy(); rowr = T[r].copy()
        ybr = yb[r]; dbk = db[k]
        yb -= colk * (ybr / p); yb[r] = ybr / p
        db -= dbk * rowr / p; db[k] = -dbk / p
        pivot(T, y, d, basis, nb, r, k); it += 1
    return 7, it, nrow, ncol

def run_sd(m, n, N, heur):
    for s in range(N):
        A, b, c = gen(m, n, s)
        st, act, obj, nit = ref(A, b, c)
        order = None
        if heur:
            order = np.argsort((A @ c) / np.linalg.norm(A, axis=1))
        T, y, d, basis, nb, z = crash(A, b, c, order)
        s2, it, nr, nc = self_dual(T, y, d, basis, nb)
        ok = (s2 == st) and (st != 2 or np.array_equal(np.sort(basis), act))
        print(s, st, s2, 'it', it, 'row', nr, 'col', nc, 'highs', nit, ok)

def hint_primal(A, b, c, x0, tol=1e-9, bland=False):
    """primal orientation. rows: basic slacks (m, shrinking to m-n live), cols: nonbasic (free x offsets -> slacks).
    returns status, active set, (crash pivots, phase2 pivots)"""
    m, n = A.shape
    P = A.copy()                 # s = sval - P @ xN
    s = b - A @ x0
    assert (s > 0).all()
    g = c.copy()                 # reduced costs of nonbasic
    row_alive = np.ones(m, bool)
    col_free = np.ones(n, bool)  # column still a free variable
    rowvar = np.arange(m)        # slack index for live rows
    colvar = -np.ones(n, int)    # slack index for nonbasic slack columns
    it1 = it2 = 0
    while True:
        if col_free.any():
            cand = np.where(col_free, np.abs(g), -1)
            j = int(np.argmax(cand))
            if cand[j] <= tol:
                # zero reduced cost free var: still needs to be made basic for a vertex; pick any direction
                sign = 1.0
            else:
                sign = -np.sign(g[j])  # move x_j in direction 'sign' decreases cost
            phase = 1
        else:
            j = int(np.argmin(g))
            if g[j] >= -tol: break
            sign = 1.0
            phase = 2
        col = P[:, j] * sign      # s decreases at rate col
        rat = np.where(row_alive & (col > tol), s / np.where(col > tol, col, 1), np.inf)
        r = int(np.argmin(rat))
        if not np.isfinite(rat[r]):
            return 5, None, (it1, it2)
        # pivot
        p = P[r, j]
        rowr = P[r].copy(); colj = P[:, j].copy(); sr = s[r]; gj = g[j]
        P -= np.outer(colj, rowr) / p
        P[r] = rowr / p; P[:, j] = -colj / p; P[r, j] = 1 / p
        s -= colj * (sr / p); s[r] = sr / p
        g -= gj * rowr / p; g[j] = -gj / p
        if col_free[j]:
            col_free[j] = False; row_alive[r] = False
            colvar[j] = rowvar[r]; rowvar[r] = -1
            it1 += 1
        else:
            colvar[j], rowvar[r] = rowvar[r], colvar[j]
            it2 += 1
    return 2, np.sort(colvar), (it1, it2)

def run_hint(m, n, N):
    for sd in range(N):
        np.random.seed(sd)
        A = np.random.randn(m, n); x0 = np.random.randn(n)
        b = A.dot(x0) + np.absolute(np.random.randn(m)); c = np.absolute(np.random.randn(n))
        st, act, obj, nit = ref(A, b, c)
        s2, a2, its = hint_primal(A, b, c, x0)
        ok = (s2 == st) and (st != 2 or np.array_equal(a2, act))
        print(sd, st, s2, its, 'highs', nit, ok)

def artcost_phase1(T, y, d, basis, nb, tol=1e-9, maxit=100000, yhat0=None):
    """Phase 1 for primal feasibility (d>=0) using an artificial positive 'rhs' yhat (all ones) and Dantzig
    pricing on d (most negative), ratio test on yhat.  (In primal orientation: dual simplex with artificial costs.)"""
    n, q = T.shape
    yh = np.ones(n) if yhat0 is None else yhat0.copy()
    it = 0
    while it < maxit:
        k = int(np.argmin(d))
        if d[k] >= -tol: return 2, it
        col = T[:, k]
        rat = np.where(col > tol, yh / np.where(col > tol, col, 1), np.inf)
        r = int(np.argmin(rat))
        if not np.isfinite(rat[r]): return 3, it     # primal infeasible
        p = T[r, k]; colk = T[:, k].copy(); yhr = yh[r]
        yh -= colk * (yhr / p); yh[r] = yhr / p
        pivot(T, y, d, basis, nb, r, k); it += 1
    return 7, it

def run_artcost(m, n, N, heur=True):
    tot = []
    for s in range(N):
        A, b, c = gen(m, n, s)
        st, act, obj, nit = ref(A, b, c)
        order = np.argsort((A @ c) / np.linalg.norm(A, axis=1)) if heur else None
        T, y, d, basis, nb, z = crash(A, b, c, order)
        s1, it1 = artcost_phase1(T, y, d, basis, nb)
        s2, it2 = dual_on_dual(T, y, d, basis, nb) if s1 == 2 else (s1, 0)
        ok = (s2 == st) and (st != 2 or np.array_equal(np.sort(basis), act))
        print(s, st, s2, 'it', it1, it2, 'highs', nit, ok)
        tot.append(it1 + it2)
    print('mean', np.mean(tot))

def crash_dynamic(A, b, c, mode='viol'):
    """crash choosing at each step the constraint (column of M) dynamically: most violated current slack first
    (d_j most negative); fallback: heuristic cosine order. Pivot row = max |entry| among remaining artificial rows."""
    m, n = A.shape
    M = np.zeros((n + 1, m + 1))
    M[:n, :m] = -A.T; M[:n, m] = c; M[n, :m] = b
    score = (A @ c) / np.linalg.norm(A, axis=1)
    rows_left = np.ones(n, dtype=bool)
    used = np.zeros(m, dtype=bool)
    basis = -np.ones(n, dtype=int)
    nrm = np.linalg.norm(A, axis=1)
    for k in range(n):
        d = M[n, :m].copy()
        if mode == 'viol':
            cand = np.where(used, np.inf, d / nrm)
            j = int(np.argmin(cand))
            if cand[j] >= -1e-9:
                cand = np.where(used, np.inf, score); j = int(np.argmin(cand))
        elif mode == 'mix':
            cand = np.where(used, np.inf, np.minimum(d / nrm, 0) * 10 + score); j = int(np.argmin(cand))
        col = np.abs(M[:n, j]) * rows_left
        r = int(np.argmax(col))
        M[r] /= M[r, j]
        for i in range(n + 1):
            if i != r: M[i] -= M[i, j] * M[r]
        rows_left[r] = False; basis[r] = j; used[j] = True
    nb = np.array([j for j in range(m) if not used[j]])
    return M[:n][:, nb].copy(), M[:n, m].copy(), M[n, nb].copy(), basis, nb, M[n, m]

def run_dyn(m, n, N, mode):
    tot = []
    for s in range(N):
        A, b, c = gen(m, n, s)
        st, act, obj, nit = ref(A, b, c)
        T, y, d, basis, nb, z = crash_dynamic(A, b, c, mode)
        ninf = (d < -1e-9).sum()
        s1, it1 = artcost_phase1(T, y, d, basis, nb)
        s2, it2 = dual_on_dual(T, y, d, basis, nb) if s1 == 2 else (s1, 0)
        ok = (s2 == st) and (st != 2 or np.array_equal(np.sort(basis), act))
        print(s, st, s2, 'ninf', ninf, 'it', it1, it2, 'highs', nit, ok)
        tot.append(it1 + it2)
    print('mean', np.mean(tot))
