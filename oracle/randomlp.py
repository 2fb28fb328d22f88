"""Oracle for G1-G3 and S1-S6 of SURVEY.md section 8(a).  Test infrastructure (see oracle/__init__.py)."""
import multiprocessing as mp
import os

import numpy as np
from scipy.linalg import lu_factor, lu_solve
from scipy.optimize import linprog

# Gurobi status codes the reference switches on (gurobi_lp.py:447-461)
LOADED, OPTIMAL, INFEASIBLE, INF_OR_UNBD, UNBOUNDED = 1, 2, 3, 4, 5
ITERATION_LIMIT, NUMERIC = 7, 12
ACTIVE_THRESHOLD = 1e-7            # gurobi_lp.py:437

# scipy/HiGHS status -> Gurobi code
_HIGHS_TO_GUROBI = {0: OPTIMAL, 1: ITERATION_LIMIT, 2: INFEASIBLE, 3: UNBOUNDED, 4: NUMERIC}


def seed_schedule(seed, num_lps):
    """randomlp_dataset.py:37-42 -- reseed the global legacy stream, draw the stride, lay out per-LP seeds."""
    np.random.seed(seed)
    step = np.random.randint(1, 1000)
    return [seed + i * step for i in range(num_lps)]


def generate_instance(m, n, seed=None):
    """randomlp_dataset.py:76-84 -- draw order is m*n, n, m, n normals from the (re)seeded global stream."""
    if seed is not None:
        np.random.seed(seed)
    A = np.random.randn(m, n)
    b = A.dot(np.random.randn(n)) + np.absolute(np.random.randn(m))
    c = np.absolute(np.random.randn(n))
    return A, b, c


def active_constraints(A, b, x, threshold=ACTIVE_THRESHOLD):
    """gurobi_lp.py:435-443 -- slack = b - A x, zero out |slack| <= 1e-7, indices where slack == 0."""
    slack = b - A.dot(x)
    slack[np.abs(slack) <= threshold] = 0
    return (slack == 0).nonzero()[0]


def polish_vertex(A, b, c, x, feas_tol=1e-9):
    """Checker-side sharpening of a solver's x (not part of the reference's algorithm; gurobi_lp.py:440-441 takes the
    solver's x as is).  HiGHS returns x to its own 1e-7 feasibility tolerance, i.e. up to ~1e-8 relative off the
    vertex and with up to 5e-7 of slack on active rows -- too coarse to check a 1e-9 bar or an absolute 1e-7 label
    threshold.  So: take the n rows of smallest |slack| at the solver's x as its active set, solve A_B x = b_B with
    iterative refinement in np.longdouble (residual in extended precision, correction through the float64 LU), and
    accept the result only with an OPTIMALITY CERTIFICATE -- primal feasible for all m rows and dual feasible
    (c = A_B' y, y <= 0) -- which makes it the optimal vertex whatever the solver's tolerances were.

    Returns (x, certified).  Uncertified instances (degenerate / ambiguous active set) keep the solver's x."""
    m, n = A.shape
    if m < n:
        return x, False
    slack = b - A.dot(x)
    rows = np.sort(np.argsort(np.abs(slack), kind='stable')[:n])
    AB = A[rows]
    try:
        lu = lu_factor(AB)
    except (ValueError, np.linalg.LinAlgError):
        return x, False
    if not np.all(np.isfinite(lu[0])) or np.abs(np.diag(lu[0])).min() < 1e-13:
        return x, False
    ABl = AB.astype(np.longdouble)
    bBl = b[rows].astype(np.longdouble)
    xl = lu_solve(lu, b[rows]).astype(np.longdouble)
    for _ in range(3):
        r = bBl - ABl.dot(xl)
        xl = xl + lu_solve(lu, r.astype(np.float64)).astype(np.longdouble)
    xp = xl.astype(np.float64)
    scale = max(1.0, float(np.abs(xp).max()))
    sl = b.astype(np.longdouble) - A.astype(np.longdouble).dot(xl)
    y = lu_solve(lu, c, trans=1)                       # A_B' y = c ; optimal iff y <= 0  (min c'x, A x <= b)
    ok = bool(sl.min() >= -feas_tol * scale) and bool(y.max() <= feas_tol * max(1.0, float(np.abs(y).max())))
    ok = ok and bool(np.abs(xp - x).max() <= 1e-5 * scale)
    return (xp, True) if ok else (x, False)


class _ModelView(object):
    """The two attributes the reference reads off the Gurobi model (randomlp_dataset.py:117, gurobi_lp.py:462)."""

    def __init__(self):
        self.status = LOADED
        self._obj = None

    @property
    def objVal(self):
        if self._obj is None:                       # Gurobi raises when no solution is available
            raise AttributeError('objVal is unavailable: model has no solution')
        return self._obj


class LinProg(object):
    """Method surface of the reference's LinProg (gurobi_lp.py:11-29, 428-465) over HiGHS dual simplex."""

    def __init__(self, A, b, c, obj='min', ops=None, method='highs-ds', polish=True):
        self.A, self.b, self.c = np.asarray(A, float), np.asarray(b, float), np.asarray(c, float)
        self.m, self.n = self.A.shape
        if obj not in ('min', 'max'):
            raise ValueError                        # gurobi_lp.py:421
        self.obj, self.ops = obj, ops
        if ops is not None and any(op not in ('<', '>', '=') for op in ops):
            raise ValueError                        # gurobi_lp.py:409
        self.model = _ModelView()
        self.x = None
        self._method = method
        self._polish = polish
        self.nit = 0
        self.x_raw = None
        self.certified = False

    def optimize(self):
        ops = self.ops if self.ops else ['<'] * self.m
        ub = [i for i, op in enumerate(ops) if op != '=']
        eq = [i for i, op in enumerate(ops) if op == '=']
        sign = np.array([-1.0 if ops[i] == '>' else 1.0 for i in ub])
        cost = self.c if self.obj == 'min' else -self.c
        res = linprog(cost,
                      A_ub=self.A[ub] * sign[:, None] if ub else None, b_ub=self.b[ub] * sign if ub else None,
                      A_eq=self.A[eq] if eq else None, b_eq=self.b[eq] if eq else None,
                      bounds=(None, None), method=self._method)
        self.model.status = _HIGHS_TO_GUROBI.get(res.status, NUMERIC)
        self.nit = int(getattr(res, 'nit', 0))
        if res.status == 0:
            self.x = self.x_raw = np.array(res.x, float)
            plain = self.obj == 'min' and not eq and all(op == '<' for op in ops)
            if self._polish and plain:
                self.x, self.certified = polish_vertex(self.A, self.b, self.c, self.x_raw)
            self.model._obj = float(self.c.dot(self.x))

    def get_statuscode(self):
        return self.model.status

    def get_active_constraints(self):
        return active_constraints(self.A, self.b, self.x)


def create_lp_problem(m, n, seed=None, with_stats=False, method='highs-ds', polish=True):
    """randomlp_dataset.py:65-128 restated; non-optimal instances get objval None instead of raising (B2)."""
    A, b, c = generate_instance(m, n, seed)
    ops = ['<'] * m
    lp = LinProg(A, b, c, 'min', ops, method=method, polish=polish)
    lp.optimize()
    sc = lp.get_statuscode()
    success = sc in (LOADED, OPTIMAL)
    active = lp.get_active_constraints() if success else []
    member = set(int(i) for i in active)
    labels = [(i, 1 if i in member else 0) for i in range(m)]
    stats = None
    if with_stats:
        stats = {'id': seed, 'm': m, 'n': n, 'eq': 0, 'ineq': m, 'active': len(active), 'sc': sc,
                 'objval': lp.model.objVal if success else None, 'success': success}
    return {'A': A, 'b': b, 'c': c, 'active': active, 'labels': labels, 'stats': stats,
            'x': lp.x, 'nit': lp.nit}


def solve_batch(A, b, c, threshold=ACTIVE_THRESHOLD, method='highs-ds', polish=True, row_mask=None):
    """Batch view used by the parity tests: arrays shaped like the C-ABI outputs (include/ddb200.h).

    x / obj / labels come from the polished vertex (polish_vertex) when it carries an optimality certificate, else
    from the solver's raw x.  raw_labels = labels of the solver's raw x; oracle_tie[i] = the two label sets differ
    (the solver's tolerance, not the instance, decided a label) ; certified[i] = polished vertex accepted.
    row_mask[B,m] (optional): solve only the kept rows of each instance (the reduced LP of SURVEY.md 8(f) rank 1);
    labels are still computed over all m rows at the reduced optimum, violations = rows with slack < -threshold."""
    B, m, n = A.shape
    status = np.zeros(B, np.int32); x = np.zeros((B, n)); obj = np.zeros(B)
    labels = np.zeros((B, m), np.uint8); n_active = np.zeros(B, np.int32); nit = np.zeros(B, np.int32)
    raw_labels = np.zeros((B, m), np.uint8); raw_x = np.zeros((B, n))
    certified = np.zeros(B, bool); oracle_tie = np.zeros(B, bool); violations = np.zeros(B, np.int32)
    min_inactive = np.full(B, np.inf); max_active = np.zeros(B)
    for i in range(B):
        keep = np.flatnonzero(row_mask[i]) if row_mask is not None else None
        Ai, bi = (A[i], b[i]) if keep is None else (A[i][keep], b[i][keep])
        if Ai.shape[0] == 0:
            status[i] = UNBOUNDED
            continue
        lp = LinProg(Ai, bi, c[i], 'min', None, method=method, polish=polish)
        lp.optimize()
        status[i] = lp.get_statuscode(); nit[i] = lp.nit
        if status[i] == OPTIMAL:
            x[i] = lp.x; raw_x[i] = lp.x_raw; obj[i] = lp.model.objVal; certified[i] = lp.certified
            slack = b[i] - A[i].dot(lp.x)
            act = np.abs(slack) <= threshold
            labels[i] = act; n_active[i] = act.sum()
            raw_labels[i] = np.abs(b[i] - A[i].dot(lp.x_raw)) <= threshold
            oracle_tie[i] = (raw_labels[i] != labels[i]).any()
            violations[i] = (slack < -threshold).sum()
            if (~act).any(): min_inactive[i] = np.abs(slack[~act]).min()
            if act.any(): max_active[i] = np.abs(slack[act]).max()
    return dict(status=status, x=x, obj=obj, labels=labels, n_active=n_active, nit=nit, raw_labels=raw_labels,
                raw_x=raw_x, certified=certified, oracle_tie=oracle_tie, violations=violations,
                min_inactive=min_inactive, max_active=max_active)


def _pool_job(args):
    os.environ['OMP_NUM_THREADS'] = '1'
    A, b, c, kw = args
    return solve_batch(A, b, c, **kw)


def solve_batch_parallel(A, b, c, procs=None, **kw):
    """solve_batch over a fork pool (one instance block per task); same return layout."""
    B = A.shape[0]
    procs = procs or min(os.cpu_count() or 1, 64)
    if procs <= 1 or B < 2 * procs:
        return solve_batch(A, b, c, **kw)
    mask = kw.pop('row_mask', None)
    parts = [p for p in np.array_split(np.arange(B), min(B, procs * 4)) if len(p)]
    jobs = [(A[p], b[p], c[p], dict(kw, row_mask=None if mask is None else mask[p])) for p in parts]
    with mp.get_context('fork').Pool(procs) as pool:
        outs = pool.map(_pool_job, jobs)
    return {k: np.concatenate([o[k] for o in outs]) for k in outs[0]}


class RandomLPDataset(object):
    """randomlp_dataset.py:12-63 restated (plain sequence; torch's Dataset base adds nothing to the semantics)."""

    def __init__(self, m, n, num_lps=1, test=False, seed=3231):
        self.m, self.n, self.seed, self.test_mode = m, n, seed, test
        self._seeds = seed_schedule(seed, num_lps)
        self._problems = [create_lp_problem(m, n, seed=s, with_stats=True) for s in self._seeds]

    def __len__(self):
        return len(self._problems)

    def __getitem__(self, idx):
        p = self._problems[idx % len(self._problems)]
        return {'lp': {'A': p['A'], 'b': p['b'], 'c': p['c']}, 'labels': p['labels']}

    def get_lp_params(self):
        return [p['stats'] for p in self._problems]
