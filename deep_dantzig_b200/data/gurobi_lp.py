"""Drop-in for the reference's ``data.gurobi_lp.LinProg`` (src/data/gurobi_lp.py:9-29, 428-465) on the B200 solver.

The reference builds a Gurobi model term by term in Python (``dot``/``_add_constraints``, :370-412 -- m*n interpreter
level calls per LP) and calls ``model.optimize()``.  Here the model *is* the (A, b, c) arrays and ``optimize()`` is a
batch-of-one call into ``ddb_solve_label_host``; ``LinProg.solve_many`` is the batched form the dataset uses.
"""
import json
import os
import time

import numpy as np
import torch

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


    # ------------------------------------------------------------------------------------------------------------
    # MPS / PLNN ingestion (reference gurobi_lp.py:64-368).  Gurobi's ``read`` is replaced by data.mps.read_mps; the
    # ``<name>.info`` side file the reference expects next to every ``.mps`` (optimal x and the names of the active
    # constraints, written by a Gurobi run that is not part of the reference tree) is produced by ``solve_mps`` below
    # with the B200 solver.
    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _info(mps_path):
        with open(os.path.splitext(mps_path)[0] + '.info', 'r') as f:
            return json.load(f)

    @staticmethod
    def solve_mps(mps_path, device=0, write_info=True, threshold=DEFAULT_THRESHOLD):
        """Solve an MPS model on the GPU and (optionally) write ``<name>.info`` = {'active': names of the matrix
        constraints with |b - a.x| <= threshold, 'x_opt': {variable: value}, 'objval', 'sc', 'time'}."""
        from .mps import read_mps
        from .mps2numpy import model2numpy
        model = read_mps(mps_path)
        item = model2numpy(model, standardize=True)
        by_index = {i: name for name, i in item['cnames'].items()}
        ops = [item['csenses'][by_index[i]] for i in range(item['A'].shape[0])]
        lp = LinProg(item['A'], item['b'], item['c'], item['obj'], ops, device=device)
        t0 = time.perf_counter()
        lp.optimize()
        dt = time.perf_counter() - t0
        info = {'sc': int(lp.model.status), 'time': dt, 'active': [], 'x_opt': {}, 'objval': None}
        if lp.model.status == ST_OPTIMAL:
            vs = model.getVars()
            info['x_opt'] = {v.VarName: float(lp.x[j]) for j, v in enumerate(vs)}
            matrix = set(c.ConstrName for c in model.getConstrs())
            act = lp.get_active_constraints()
            info['active'] = [by_index[int(i)] for i in act if by_index[int(i)] in matrix]
            sign = -1.0 if model.ModelSense == -1 else 1.0           # model2numpy minimises -c for a 'max' model
            info['objval'] = sign * float(lp.model.objVal) + model.ObjCon
        if write_info:
            with open(os.path.splitext(mps_path)[0] + '.info', 'w') as f:
                json.dump(info, f)
        return info

    @staticmethod
    def ineq_num(mps_path):
        """gurobi_lp.py:64-92: counts of (in)active matrix inequalities and equalities of one LP."""
        from .mps import read_mps
        model = read_mps(mps_path)
        mineq = set(c.ConstrName for c in model.getConstrs() if c.Sense != '=')
        meq = set(c.ConstrName for c in model.getConstrs() if c.Sense == '=')
        active = set(LinProg._info(mps_path)['active'])
        num_active = len(active & mineq)
        num_inactive = len(mineq - active)
        return {'path': mps_path, 'num_active_ineq': num_active, 'num_inactive_ineq': num_inactive, 'num_ineq': len(mineq),
                'num_eq': len(meq), 'num_constrs': len(meq) + len(mineq), 'num_pos': num_active,
                'num_neg': num_inactive + len(meq)}

    @staticmethod
    def has_matrix_inequalities(mps_path):
        from .mps import read_mps
        return any(c.Sense != '=' for c in read_mps(mps_path).getConstrs())

    @staticmethod
    def mps_to_bipartite_graph(mps_path):
        """gurobi_lp.py:189-262: constraint / variable / edge features by name, the active set and, per variable, whether
        the optimum sits on its lower or upper bound."""
        from .mps import INF, read_mps
        model = read_mps(mps_path)
        c_feats = {c.ConstrName: {'sense': c.Sense, 'rhs': c.RHS} for c in model.getConstrs()}
        v_feats = {v.VarName: {'lb': v.LB if v.LB > -INF else None, 'ub': v.UB if v.UB < INF else None, 'obj': v.Obj}
                   for v in model.getVars()}
        e_feats = [{'vname': v.VarName, 'cname': c.ConstrName, 'coeff': coeff}
                   for c in model.getConstrs() for v, coeff in c.terms if coeff != 0.0]
        info = LinProg._info(mps_path)
        x_opt = info['x_opt']
        v_bounds = {vname: None for vname in v_feats}
        for vname, vf in v_feats.items():
            xj = x_opt[vname]
            if vf['lb'] and xj == vf['lb']:
                v_bounds[vname] = 'lb'
            if vf['ub'] and xj == vf['ub']:
                v_bounds[vname] = 'ub'
        return {'c_feats': c_feats, 'v_feats': v_feats, 'e_feats': e_feats, 'active': list(info['active']),
                'v_bounds': v_bounds, 'mps_path': mps_path}

    @staticmethod
    def getitem_bipartite(mps_path, reference_exact=False):
        """gurobi_lp.py:94-187: the bipartite-graph item of one LP -- c_feats [m,3] = (is_inequality, rhs, is_bound),
        v_feats [n,1] = (objective coefficient), e_feats = {'i': [[row, col]...], 'coeffs': [...]}, c_labels, in_loss
        (the matrix inequalities), dims.  Variable bounds become extra constraint nodes flagged is_bound (a bound of exactly
        0.0 is skipped, as the reference's truthiness test does).  '>' rows are flipped to '<': right-hand side AND
        coefficients (the reference flips only the right-hand side -- SURVEY.md B13, a chained comparison that is never true;
        ``reference_exact=True`` reproduces that)."""
        graph = LinProg.mps_to_bipartite_graph(mps_path)
        c_feats, v_feats, e_feats = graph['c_feats'], graph['v_feats'], list(graph['e_feats'])
        v_bounds, active = graph['v_bounds'], list(graph['active'])
        for cname in c_feats:
            c_feats[cname]['is_bound'] = 0
        b_feats, b_edges, b_active = {}, [], []
        for vname, vf in v_feats.items():
            cname = None
            if vf['lb']:
                cname = '%s_lb' % vname
                b_feats[cname] = {'sense': '>', 'rhs': vf['lb'], 'is_bound': 1}
                b_edges.append({'vname': vname, 'cname': cname, 'coeff': 1.0})
            if vf['ub']:
                cname = '%s_ub' % vname
                b_feats[cname] = {'sense': '<', 'rhs': vf['ub'], 'is_bound': 1}
                b_edges.append({'vname': vname, 'cname': cname, 'coeff': 1.0})
            if v_bounds[vname] and cname is not None:
                b_active.append(cname)
        c_feats.update(b_feats)
        e_feats.extend(b_edges)
        active.extend(b_active)
        cnames, vnames = list(c_feats.keys()), list(v_feats.keys())
        m, n = len(cnames), len(vnames)
        for cname in cnames:
            c_feats[cname]['is_inequality'] = 1 if c_feats[cname]['sense'] != '=' else 0
        if not reference_exact:
            for e in e_feats:
                if c_feats[e['cname']]['sense'] == '>':
                    e['coeff'] = -1.0 * e['coeff']
        for cname in cnames:
            if c_feats[cname]['sense'] == '>':
                c_feats[cname]['rhs'] = -1.0 * c_feats[cname]['rhs']
                c_feats[cname]['sense'] = '<'
        act = set(active)
        c_labels = [bool(c_feats[c]['sense'] != '=' and c in act and c_feats[c]['is_bound'] == 0) for c in cnames]
        cf = [[c_feats[c][k] for k in ('is_inequality', 'rhs', 'is_bound')] for c in cnames]
        vf = [[v_feats[v]['obj']] for v in vnames]
        cx = {c: j for j, c in enumerate(cnames)}
        vx = {v: j for j, v in enumerate(vnames)}
        edges = {'i': [[cx[e['cname']], vx[e['vname']]] for e in e_feats], 'coeffs': [e['coeff'] for e in e_feats]}
        cf_t = torch.FloatTensor(cf)
        in_loss = [int(i) for i in torch.nonzero((cf_t[:, 0] == 1) & (cf_t[:, 2] == 0)).reshape(-1)]
        return {'c_feats': cf_t, 'v_feats': torch.FloatTensor(vf), 'e_feats': edges, 'c_labels': torch.FloatTensor(c_labels),
                'in_loss': in_loss, 'dims': {'m': m, 'n': n}, 'mps_path': mps_path}

    @staticmethod
    def getitem_complete(mps_path):
        """gurobi_lp.py:295-368 (the working body, ``getitem_complete_copy``): {'lp': {'A','b','c'} in the
        standardised form of mps2numpy.model2numpy, node_features (1 for a '<' row, 0 for '='; a trailing 0 for the cost
        node), node_labels (active matrix rows; a bound row is positive when x sits on it), in_loss, mps_path}."""
        from .mps import read_mps
        from .mps2numpy import model2numpy
        model = read_mps(mps_path)
        item = model2numpy(model, standardize=True)
        name2index, name2sense = item['cnames'], item['csenses']
        rows = item['A'].shape[0]
        node_features = [None] * rows
        for k, sense in name2sense.items():
            node_features[name2index[k]] = 1 if sense == '<' else 0
        node_labels = [None] * rows
        info = LinProg._info(mps_path)
        for c in model.getConstrs():
            node_labels[name2index[c.ConstrName]] = 0
        for c in info['active']:
            node_labels[name2index[c]] = 1
        for k, v in info['x_opt'].items():
            lb, ub = item['bounds'][k]['lb'], item['bounds'][k]['ub']
            if lb:
                node_labels[name2index[lb['name']]] = 1 if v == -lb['val'] else 0    # lower bounds are stored flipped
            if ub:
                node_labels[name2index[ub['name']]] = 1 if v == ub['val'] else 0
        node_features.append(0)
        return {'lp': {'A': item['A'], 'b': item['b'], 'c': item['c']}, 'node_features': np.asarray(node_features),
                'node_labels': np.asarray(node_labels), 'in_loss': item['in_loss'], 'mps_path': mps_path}

    getitem_complete_copy = getitem_complete
