"""ctypes binding of include/ddb200.h.  Loading is lazy for the symbol table but strict: a missing library or a
missing symbol raises -- nothing falls back to Python/numpy."""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = 'libddb200.so'

# status codes == Gurobi's (reference src/data/gurobi_lp.py:447-461)
ST_LOADED, ST_OPTIMAL, ST_INFEASIBLE, ST_INF_OR_UNBD, ST_UNBOUNDED = 1, 2, 3, 4, 5
ST_ITERATION_LIMIT, ST_NUMERIC = 7, 12
DEFAULT_THRESHOLD = 1e-7          # reference src/data/gurobi_lp.py:437


class DdbError(RuntimeError):
    pass


def library_path():
    # DDB200_LIBRARY: development switch (A/B builds of the library); the shipped path is the in-tree libddb200.so
    return os.environ.get('DDB200_LIBRARY') or os.path.join(_HERE, _LIB_NAME)


_lock = threading.Lock()
_lib = None

_vp, _i32, _i64, _u64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double

# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header is exported
SIGNATURES = {
    'ddb_abi_version': (C.c_int, []),
    'ddb_last_error': (C.c_char_p, []),
    'ddb_create': (C.c_int, [C.c_int, C.POINTER(_vp)]),
    'ddb_destroy': (C.c_int, [_vp]),
    'ddb_device_info': (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(_i64)]),
    'ddb_solve_plan': (C.c_int, [_vp, C.c_int, C.c_int]),
    'ddb_set_solve_plan': (C.c_int, [_vp, C.c_int]),
    'ddb_set_fused_mode': (C.c_int, [_vp, C.c_int]),
    'ddb_generate_dev': (C.c_int, [_vp, _u64, _i64, _i64, C.c_int, C.c_int, _f64, _vp, _vp, _vp, _vp, _vp]),
    'ddb_solve_label_dev': (C.c_int, [_vp, _i64, C.c_int, C.c_int, _vp, _vp, _vp, _f64, _vp,
                                      _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'ddb_solve_label_host': (C.c_int, [_vp, _i64, C.c_int, C.c_int, _vp, _vp, _vp, _f64, _vp,
                                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'ddb_generate_solve_label_dev': (C.c_int, [_vp, _u64, _i64, _i64, C.c_int, C.c_int, _f64, _f64,
                                               _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'ddb_generate_solve_label_host': (C.c_int, [_vp, _u64, _i64, _i64, C.c_int, C.c_int, _f64, _f64,
                                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'ddb_s2v_param_count': (C.c_int, [C.c_int, C.c_int]),
    'ddb_s2v_forward_dev': (C.c_int, [_vp, C.c_int, _i64, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'ddb_s2v_forward_flags_dev': (C.c_int, [_vp, C.c_int, _i64, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'ddb_s2v_loss_grad_dev': (C.c_int, [_vp, C.c_int, _i64, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp,
                                        C.c_float, C.c_float, _vp, _vp, _vp, _vp]),
    'ddb_s2v_loss_grad_flags_dev': (C.c_int, [_vp, C.c_int, _i64, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                              C.c_float, C.c_float, _vp, _vp, _vp, _vp]),
    'ddb_s2v_metrics_dev': (C.c_int, [_vp, _i64, _vp, _vp, _vp, C.c_float, C.c_float, C.c_float, _vp, _vp]),
    'ddb_launch_count': (_i64, [_vp]),
}


def load():
    """Return the ctypes handle of libddb200.so with all prototypes set; raise DdbError if it cannot be loaded."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            raise DdbError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` or '
                           '`make -C deep_dantzig_b200/csrc`. There is no CPU fallback.' % path)
        try:
            lib = C.CDLL(path)
        except OSError as exc:
            raise DdbError('cannot load %s: %s' % (path, exc))
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                raise DdbError('%s does not export %s (stale build?)' % (path, name))
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return _lib


def abi_version():
    return int(load().ddb_abi_version())


def check(rc, what):
    if rc != 0:
        msg = load().ddb_last_error()
        raise DdbError('%s failed (%d): %s' % (what, rc, msg.decode('utf-8', 'replace') if msg else ''))


_ctx_lock = threading.Lock()
_contexts = {}


class Context(object):
    """One ddb_ctx per (process, device)."""

    def __init__(self, device):
        lib = load()
        handle = _vp()
        check(lib.ddb_create(int(device), C.byref(handle)), 'ddb_create(device=%d)' % device)
        self.lib, self.handle, self.device = lib, handle, int(device)
        sm, maj, mnr, smem = C.c_int(), C.c_int(), C.c_int(), _i64()
        check(lib.ddb_device_info(handle, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(smem)), 'ddb_device_info')
        self.sm_count, self.cc, self.smem_optin = sm.value, (maj.value, mnr.value), smem.value

    def launch_count(self):
        return int(self.lib.ddb_launch_count(self.handle))

    def solve_plan(self, m, n):
        rc = self.lib.ddb_solve_plan(self.handle, int(m), int(n))
        if rc < 0:
            check(rc, 'ddb_solve_plan')
        return rc

    def set_solve_plan(self, plan):
        check(self.lib.ddb_set_solve_plan(self.handle, int(plan)), 'ddb_set_solve_plan')

    def set_fused_mode(self, mode):
        """0 automatic, 1 in-kernel instance generation, 2 generator kernel + solver kernel."""
        check(self.lib.ddb_set_fused_mode(self.handle, int(mode)), 'ddb_set_fused_mode')

    def close(self):
        if self.handle:
            self.lib.ddb_destroy(self.handle)
            self.handle = None


def context(device=0):
    with _ctx_lock:
        ctx = _contexts.get(int(device))
        if ctx is None:
            ctx = _contexts[int(device)] = Context(int(device))
        return ctx
