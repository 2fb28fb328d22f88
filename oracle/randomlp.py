"""Oracle for G1-G3 and S1-S6 of SURVEY.md section 8(a).  Test infrastructure (see oracle/__init__.py)."""
import numpy as np
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

    def __init__(self, A, b, c, obj='min', ops=None, method='highs-ds'):
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
        self.nit = 0

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
            self.x = np.array(res.x, float)
            self.model._obj = float(self.c.dot(self.x))

    def get_statuscode(self):
        return self.model.status

    def get_active_constraints(self):
        return active_constraints(self.A, self.b, self.x)


def create_lp_problem(m, n, seed=None, with_stats=False, method='highs-ds'):
    """randomlp_dataset.py:65-128 restated; non-optimal instances get objval None instead of raising (B2)."""
    A, b, c = generate_instance(m, n, seed)
    ops = ['<'] * m
    lp = LinProg(A, b, c, 'min', ops, method=method)
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


def solve_batch(A, b, c, threshold=ACTIVE_THRESHOLD, method='highs-ds'):
    """Batch view used by the parity tests: arrays shaped like the C-ABI outputs (include/ddb200.h)."""
    B, m, n = A.shape
    status = np.zeros(B, np.int32); x = np.zeros((B, n)); obj = np.zeros(B)
    labels = np.zeros((B, m), np.uint8); n_active = np.zeros(B, np.int32); nit = np.zeros(B, np.int32)
    min_inactive = np.full(B, np.inf); max_active = np.zeros(B)
    for i in range(B):
        lp = LinProg(A[i], b[i], c[i], 'min', None, method=method)
        lp.optimize()
        status[i] = lp.get_statuscode(); nit[i] = lp.nit
        if status[i] == OPTIMAL:
            x[i] = lp.x; obj[i] = lp.model.objVal
            slack = b[i] - A[i].dot(lp.x)
            act = np.abs(slack) <= threshold
            labels[i] = act; n_active[i] = act.sum()
            if (~act).any(): min_inactive[i] = np.abs(slack[~act]).min()
            if act.any(): max_active[i] = np.abs(slack[act]).max()
    return dict(status=status, x=x, obj=obj, labels=labels, n_active=n_active, nit=nit,
                min_inactive=min_inactive, max_active=max_active)


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
