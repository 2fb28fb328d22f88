"""Drop-in for the reference's ``data.gurobi_lp.LinProg`` (src/data/gurobi_lp.py:9-29, 428-465) on the B200 solver.

The reference builds a Gurobi model term by term in Python (``dot``/``_add_constraints``, :370-412 -- m*n interpreter
level calls per LP) and calls ``model.optimize()``.  Here the model *is* the (A, b, c) arrays and ``optimize()`` is a
batch-of-one call into ``ddb_solve_label_host``; ``LinProg.solve_many`` is the batched form the dataset uses.
"""
import numpy as np

from .. import solver
from .._lib import DEFAULT_THRESHOLD, ST_LOADED, ST_OPTIMAL

STATUSCODES = {1: 'loaded', 2: 'optimal', 3: 'infeasible', 4: 'inf_or_unbd', 5: 'unbounded', 6: 'cutoff',
               7: 'iteration_limit', 8: 'node_limit', 9: 'time_limit', 10: 'solution_limit', 11: 'interrupted',
               12: 'numeric', 13: 'suboptimal', 14: 'inprogress', 15: 'user_obj_limit'}   # gurobi_lp.py:447-461


class _ModelView(object):
    """``lp.model.status`` / ``lp.model.objVal`` as the reference reads them (gurobi_lp.py:462, randomlp_dataset.py:117)."""

    def __init__(self):
        self.status = ST_LOADED
        self._obj = None

    @property
    def objVal(self):
        if self._obj is None:
            raise AttributeError('objVal is unavailable: the model has no optimal solution')   # as Gurobi (B2)
        return self._obj


class LinProg(object):

    def __init__(self, A, b, c, obj='min', ops=None, device=0):
        self.A = np.ascontiguousarray(A, dtype=np.float64)
        self.m, self.n = self.A.shape
        self.b = np.ascontiguousarray(b, dtype=np.float64)
        self.c = np.ascontiguousarray(c, dtype=np.float64)
        if obj not in ('min', 'max'):
            raise ValueError                                    # gurobi_lp.py:421
        if ops is not None:
            for op in ops:
                if op not in ('<', '>', '='):
                    raise ValueError                            # gurobi_lp.py:409
        self.obj, self.ops = obj, ops
        self.x = None
        self.model = _ModelView()
        self._device = device
        self._res = None

    def _canonical(self):
        """min c'x, Ax <= b: '>' rows and 'max' objectives are sign flips; an '=' row a.x = b_i becomes the pair
        a.x <= b_i, -a.x <= -b_i (the second halves are appended behind the m original rows, so the first m labels the
        device returns are the labels of the caller's rows: |b_i - a_i.x| does not depend on the sign)."""
        A, b, c = self.A, self.b, self.c
        if self.ops is not None and any(op == '>' for op in self.ops):
            sign = np.array([-1.0 if op == '>' else 1.0 for op in self.ops])
            A, b = A * sign[:, None], b * sign
        if self.ops is not None and any(op == '=' for op in self.ops):
            eq = [i for i, op in enumerate(self.ops) if op == '=']
            A, b = np.vstack((A, -A[eq])), np.concatenate((b, -b[eq]))
        if self.obj == 'max':
            c = -c
        return np.ascontiguousarray(A), np.ascontiguousarray(b), c

    def optimize(self):
        A, b, c = self._canonical()
        res = solver.solve_label_host(A[None], b[None], c[None], DEFAULT_THRESHOLD, device=self._device)
        self._res = res
        self.model.status = int(res['status'][0])
        if self.model.status == ST_OPTIMAL:
            self.x = res['x'][0].copy()
            self.model._obj = float(self.c.dot(self.x))
        return

    def get_active_constraints(self):
        """gurobi_lp.py:435-443: indices with |b - A x| <= 1e-7 (computed on the device from the same A, b, x)."""
        return np.flatnonzero(self._res['labels'][0][:self.m]).astype(np.int64)

    def get_statuscode(self):
        s = self.model.status
        if s not in (1, 2):
            print(STATUSCODES.get(s, str(s)))                   # gurobi_lp.py:463-464
        return s

    @staticmethod
    def solve_many(A, b, c, threshold=DEFAULT_THRESHOLD, device=0):
        """Batched form: A[B,m,n], b[B,m], c[B,n] numpy -> SolveResult of numpy arrays."""
        return solver.solve_label_host(A, b, c, threshold, device=device)
