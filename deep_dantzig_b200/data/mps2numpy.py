"""Drop-in for the reference's ``data.mps2numpy`` (src/data/mps2numpy.py:6-128): an MPS model as dense numpy arrays

    min c'x  s.t.  A x (<, =) b      with every variable bound turned into an extra '<' row,

in the reference's row order (matrix constraints first, then per variable its lower-bound row ``-x_j <= -lb`` and its
upper-bound row ``x_j <= ub``), with the same name / sense / bound bookkeeping.  The model comes from
:func:`deep_dantzig_b200.data.mps.read_mps` instead of Gurobi's ``read``."""
import numpy as np

from .mps import INF, read_mps


def _has_lb(v):
    return v.LB > -INF          # mps2numpy.py:29-35


def _has_ub(v):
    return v.UB < INF           # mps2numpy.py:36-42


def constr2numpy(model, standardize=True):
    """Yield (i, a_i, b_i, sense, name) per matrix constraint; '>' rows are flipped to '<' when standardising
    (mps2numpy.py:11-26)."""
    n = len(model.getVars())
    for i, con in enumerate(model.getConstrs()):
        ai = np.zeros(n)
        for v, coeff in con.terms:
            ai[v.index] = coeff
        bi, sense = con.RHS, con.Sense
        if standardize and sense == '>':
            ai, bi, sense = -ai, -bi, '<'
        yield i, ai, bi, sense, con.ConstrName


def bounds2numpy(model, m):
    """Variable bounds as '<' rows appended behind the m matrix rows (mps2numpy.py:28-70): a lower bound becomes
    ``-x_j <= -lb`` named ``<var>_lb``, an upper bound ``x_j <= ub`` named ``<var>_ub``."""
    vs = model.getVars()
    rows, bs, cnames, ops, bounds = [], [], {}, {}, {}
    for j, v in enumerate(vs):
        bounds[v.VarName] = {'lb': None, 'ub': None}
        if _has_lb(v):
            r = np.zeros(len(vs)); r[j] = -1.0
            rows.append(r); bs.append(-v.LB)
            name = '%s_lb' % v.VarName
            cnames[name] = m + len(rows) - 1
            ops[name] = '<'
            bounds[v.VarName]['lb'] = {'val': -v.LB, 'sense': '<', 'name': name}
        if _has_ub(v):
            r = np.zeros(len(vs)); r[j] = 1.0
            rows.append(r); bs.append(v.UB)
            name = '%s_ub' % v.VarName
            cnames[name] = m + len(rows) - 1
            ops[name] = '<'
            bounds[v.VarName]['ub'] = {'val': v.UB, 'sense': '<', 'name': name}
    return rows, bs, ops, cnames, bounds


def model2numpy(model, standardize=True):
    """mps2numpy.py:72-126 -> {'A','b','c','obj','csenses','cnames','bounds','in_loss'}; ``in_loss`` = the matrix
    inequality rows (the only constraints the classifier is asked about)."""
    n, m = len(model.getVars()), len(model.getConstrs())
    A, b = np.zeros((m, n)), np.zeros(m)
    csenses, cnames = {}, {}
    for i, ai, bi, sense, name in constr2numpy(model, standardize):
        A[i], b[i] = ai, bi
        csenses[name], cnames[name] = sense, i
    rows, bs, bsenses, bname2index, bounds = bounds2numpy(model, m)
    if rows:
        A = np.vstack((A, np.array(rows)))
        b = np.concatenate((b, np.array(bs)))
    csenses.update(bsenses)
    cnames.update(bname2index)
    in_loss = [cnames[k] for k in cnames if csenses[k] == '<' and k not in bname2index]
    c = np.array(model.Obj, dtype=np.float64)
    sense = model.ModelSense
    if standardize and sense == -1:
        c, sense = -c, 1
    return {'A': A, 'b': b, 'c': c, 'obj': 'max' if sense == -1 else 'min', 'csenses': csenses, 'cnames': cnames,
            'bounds': bounds, 'in_loss': in_loss}


def mps2numpy(fpath, standardize=True):
    """(A, b, c, ops, obj) of an MPS file -- the tuple the reference's callers unpack (plnn_dataset.py:219); ``ops`` is the
    per-row sense list in row order.  (The reference returns model2numpy's dict here, mps2numpy.py:127-130, which its own
    caller cannot unpack; the tuple is what that caller needs.)"""
    item = model2numpy(read_mps(fpath), standardize)
    by_index = {i: name for name, i in item['cnames'].items()}
    ops = [item['csenses'][by_index[i]] for i in range(item['A'].shape[0])]
    return item['A'], item['b'], item['c'], ops, item['obj']
