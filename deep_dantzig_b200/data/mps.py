"""MPS reader / writer (fixed and free format) and the thin model view the ingestion code works on.

The reference lets Gurobi parse MPS files (``read(fpath)``, src/data/mps2numpy.py:128, src/data/gurobi_lp.py:69,90,207) and
then walks the Gurobi model object (``getVars``, ``getConstrs``, ``getRow``, ``LB``/``UB``/``Obj``, ``Sense``/``RHS``,
``ModelSense``).  Gurobi is proprietary and absent here, so this module restates the published MPS format (IBM MPSX /
lp_solve / Gurobi documentation: sections NAME, OBJSENSE, ROWS, COLUMNS, RHS, RANGES, BOUNDS, ENDATA) and exposes the same
attribute names on plain Python objects, so that the code above it reads like the reference's."""
import math

INF = 1e100          # the reference's notion of "no bound" (mps2numpy.py:33-42)


class Var(object):
    def __init__(self, name, index):
        self.VarName = self.varName = name
        self.index = index
        self.LB, self.UB, self.Obj = 0.0, math.inf, 0.0      # MPS defaults: 0 <= x < +inf


class Row(object):
    """``model.getRow(constr)``: ``size()``, ``getVar(i)``, ``getCoeff(i)``."""

    def __init__(self, terms):
        self._terms = terms

    def size(self):
        return len(self._terms)

    def getVar(self, i):
        return self._terms[i][0]

    def getCoeff(self, i):
        return self._terms[i][1]


class Constr(object):
    def __init__(self, name, sense, index):
        self.ConstrName, self.Sense, self.RHS, self.index = name, sense, 0.0, index
        self.terms = []                                         # [(Var, coeff)]


class MpsModel(object):
    def __init__(self, name=''):
        self.ModelName = name
        self.ModelSense = 1                                     # 1 minimise, -1 maximise (Gurobi convention)
        self.ObjCon = 0.0
        self._vars, self._constrs = [], []
        self._vindex, self._cindex = {}, {}

    def getVars(self):
        return list(self._vars)

    def getConstrs(self):
        return list(self._constrs)

    def getRow(self, constr):
        return Row(constr.terms)

    @property
    def Obj(self):
        return [v.Obj for v in self._vars]

    def _var(self, name):
        v = self._vindex.get(name)
        if v is None:
            v = self._vindex[name] = Var(name, len(self._vars))
            self._vars.append(v)
        return v


_SECTIONS = ('NAME', 'OBJSENSE', 'OBJSENSE_MAX', 'ROWS', 'COLUMNS', 'RHS', 'RANGES', 'BOUNDS', 'ENDATA', 'OBJSENSE_MIN')


def read_mps(path):
    """Parse an MPS file (tokens are split on white space, which covers the free format and every fixed-format file whose
    names contain no blanks).  RANGES sections are not supported (none of the reference's LP families uses them)."""
    model = MpsModel()
    section, objname = None, None
    with open(path) as f:
        for raw in f:
            line = raw.rstrip('\n')
            if not line.strip() or line.lstrip().startswith('*'):
                continue
            tok = line.split()
            if not line[0].isspace():                            # section header
                head = tok[0].upper()
                if head == 'NAME':
                    model.ModelName = tok[1] if len(tok) > 1 else ''
                    section = 'NAME'
                elif head == 'OBJSENSE':
                    section = 'OBJSENSE'
                    if len(tok) > 1:
                        model.ModelSense = -1 if tok[1].upper().startswith('MAX') else 1
                elif head in ('ROWS', 'COLUMNS', 'RHS', 'RANGES', 'BOUNDS'):
                    section = head
                elif head == 'ENDATA':
                    break
                else:
                    raise ValueError('%s: unknown MPS section %r' % (path, tok[0]))
                continue
            if section == 'OBJSENSE':
                model.ModelSense = -1 if tok[0].upper().startswith('MAX') else 1
            elif section == 'ROWS':
                kind, name = tok[0].upper(), tok[1]
                if kind == 'N':
                    if objname is None:
                        objname = name                           # the first free row is the objective, later ones are dropped
                elif kind in ('L', 'G', 'E'):
                    cobj = Constr(name, {'L': '<', 'G': '>', 'E': '='}[kind], len(model._constrs))
                    model._constrs.append(cobj)
                    model._cindex[name] = cobj
                else:
                    raise ValueError('%s: unknown row type %r' % (path, kind))
            elif section == 'COLUMNS':
                if len(tok) >= 3 and tok[1].upper() == "'MARKER'":
                    continue                                     # integrality markers: the LP relaxation is what is read
                v = model._var(tok[0])
                for q in range(1, len(tok) - 1, 2):
                    rname, val = tok[q], float(tok[q + 1])
                    if rname == objname:
                        v.Obj = val
                    elif rname in model._cindex:
                        if val != 0.0:
                            model._cindex[rname].terms.append((v, val))
                    # entries of dropped free rows are ignored
            elif section == 'RHS':
                start = 1 if len(tok) % 2 == 1 else 0            # the RHS-set name is optional
                for q in range(start, len(tok) - 1, 2):
                    rname, val = tok[q], float(tok[q + 1])
                    if rname == objname:
                        model.ObjCon = -val                      # the objective's RHS is minus the constant
                    elif rname in model._cindex:
                        model._cindex[rname].RHS = val
            elif section == 'RANGES':
                raise NotImplementedError('%s: RANGES sections are not supported' % path)
            elif section == 'BOUNDS':
                kind = tok[0].upper()
                if kind in ('FR', 'MI', 'PL', 'BV'):
                    name = tok[2] if len(tok) >= 3 else tok[1]
                    val = None
                else:
                    if len(tok) >= 4:
                        name, val = tok[2], float(tok[3])
                    else:
                        name, val = tok[1], float(tok[2])        # bound-set name omitted
                v = model._var(name)
                if kind == 'UP' or kind == 'UI':
                    v.UB = val
                    if val < 0.0 and v.LB == 0.0:
                        v.LB = -math.inf                         # the customary reading of a negative upper bound
                elif kind == 'LO' or kind == 'LI':
                    v.LB = val
                elif kind == 'FX':
                    v.LB = v.UB = val
                elif kind == 'FR':
                    v.LB, v.UB = -math.inf, math.inf
                elif kind == 'MI':
                    v.LB = -math.inf
                elif kind == 'PL':
                    v.UB = math.inf
                elif kind == 'BV':
                    v.LB, v.UB = 0.0, 1.0
                else:
                    raise ValueError('%s: unknown bound type %r' % (path, kind))
    return model


def write_mps(path, A, b, c, ops, lb=None, ub=None, obj='min', name='LP', cnames=None, vnames=None):
    """Write min/max c'x s.t. A_i x (ops_i) b_i, lb <= x <= ub as a free-format MPS file (test and tooling helper: the PLNN
    problem families the reference trains on are not shipped with it)."""
    m, n = len(b), len(c)
    cnames = cnames or ['c%d' % i for i in range(m)]
    vnames = vnames or ['x%d' % j for j in range(n)]
    with open(path, 'w') as f:
        f.write('NAME %s\n' % name)
        if obj == 'max':
            f.write('OBJSENSE\n    MAX\n')
        f.write('ROWS\n N  OBJ\n')
        for i in range(m):
            f.write(' %s  %s\n' % ({'<': 'L', '>': 'G', '=': 'E'}[ops[i]], cnames[i]))
        f.write('COLUMNS\n')
        for j in range(n):
            if c[j] != 0.0:
                f.write('    %s  OBJ  %r\n' % (vnames[j], float(c[j])))
            for i in range(m):
                if A[i][j] != 0.0:
                    f.write('    %s  %s  %r\n' % (vnames[j], cnames[i], float(A[i][j])))
            if c[j] == 0.0 and all(A[i][j] == 0.0 for i in range(m)):
                f.write('    %s  OBJ  0.0\n' % vnames[j])
        f.write('RHS\n')
        for i in range(m):
            if b[i] != 0.0:
                f.write('    RHS  %s  %r\n' % (cnames[i], float(b[i])))
        f.write('BOUNDS\n')
        for j in range(n):
            lo = 0.0 if lb is None else lb[j]
            hi = math.inf if ub is None else ub[j]
            if lo == -math.inf and hi == math.inf:
                f.write(' FR BND  %s\n' % vnames[j])
                continue
            if lo == hi:
                f.write(' FX BND  %s  %r\n' % (vnames[j], float(lo)))
                continue
            if lo == -math.inf:
                f.write(' MI BND  %s\n' % vnames[j])
            elif lo != 0.0:
                f.write(' LO BND  %s  %r\n' % (vnames[j], float(lo)))
            if hi != math.inf:
                f.write(' UP BND  %s  %r\n' % (vnames[j], float(hi)))
        f.write('ENDATA\n')
