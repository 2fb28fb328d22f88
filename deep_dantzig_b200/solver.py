"""Batched entry points over the C ABI (include/ddb200.h).  torch is used for device memory and streams only."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DEFAULT_THRESHOLD, ST_OPTIMAL  # noqa: F401


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.DdbError('deep_dantzig_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class SolveResult(dict):
    """status[B] i32, x[B,n] f64, obj[B] f64, labels[B,m] u8, n_active[B] i32, pivots[B,4] i32, ties[B] i32,
    violations[B] i32 -- device tensors (or numpy arrays from the host flavour)."""
    __getattr__ = dict.__getitem__


def _alloc_outputs(B, m, n, device):
    kw = dict(device=device)
    return SolveResult(
        status=torch.empty(B, dtype=torch.int32, **kw), x=torch.empty(B, n, dtype=torch.float64, **kw),
        obj=torch.empty(B, dtype=torch.float64, **kw), labels=torch.empty(B, m, dtype=torch.uint8, **kw),
        n_active=torch.empty(B, dtype=torch.int32, **kw), pivots=torch.empty(B, 4, dtype=torch.int32, **kw),
        ties=torch.empty(B, dtype=torch.int32, **kw), violations=torch.empty(B, dtype=torch.int32, **kw))


def solve_label(A, b, c, threshold=DEFAULT_THRESHOLD, row_mask=None, out=None):
    """Solve min c'x s.t. Ax<=b (x free) for a device-resident batch and label active constraints.

    A[B,m,n], b[B,m], c[B,n]: contiguous float64 CUDA tensors.  Asynchronous on the current stream.
    Batched replacement of LinProg(...).optimize()/get_statuscode()/get_active_constraints()
    (reference src/data/gurobi_lp.py:11-29, 428-465)."""
    _require_cuda()
    if A.dim() != 3 or b.dim() != 2 or c.dim() != 2:
        raise ValueError('expected A[B,m,n], b[B,m], c[B,n]')
    B, m, n = A.shape
    if tuple(b.shape) != (B, m) or tuple(c.shape) != (B, n):
        raise ValueError('shape mismatch: A %s b %s c %s' % (tuple(A.shape), tuple(b.shape), tuple(c.shape)))
    for t in (A, b, c):
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise ValueError('inputs must be contiguous float64 CUDA tensors')
    dev = A.device
    if row_mask is not None:
        if not (row_mask.is_cuda and row_mask.dtype == torch.uint8 and row_mask.is_contiguous()
                and tuple(row_mask.shape) == (B, m)):
            raise ValueError('row_mask must be a contiguous uint8 CUDA tensor [B,m]')
    ctx = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
    res = out if out is not None else _alloc_outputs(B, m, n, dev)
    rc = ctx.lib.ddb_solve_label_dev(ctx.handle, B, m, n, _ptr(A), _ptr(b), _ptr(c), float(threshold), _ptr(row_mask),
                                     _ptr(res['status']), _ptr(res['x']), _ptr(res['obj']), _ptr(res['labels']),
                                     _ptr(res['n_active']), _ptr(res['pivots']), _ptr(res['ties']),
                                     _ptr(res['violations']), _stream_ptr(dev))
    _lib.check(rc, 'ddb_solve_label_dev')
    return res


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def solve_label_host(A, b, c, threshold=DEFAULT_THRESHOLD, row_mask=None, device=0, out=None):
    """Same contract with HOST (numpy) buffers; the library does the chunked H2D/D2H copies itself and returns
    when the results are in the output arrays."""
    _require_cuda()
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    if A.ndim != 3:
        raise ValueError('expected A[B,m,n]')
    B, m, n = A.shape
    if b.shape != (B, m) or c.shape != (B, n):
        raise ValueError('shape mismatch')
    if row_mask is not None:
        row_mask = np.ascontiguousarray(row_mask, dtype=np.uint8)
        if row_mask.shape != (B, m):
            raise ValueError('row_mask must be [B,m]')
    ctx = _lib.context(device)
    res = out if out is not None else SolveResult(
        status=np.empty(B, np.int32), x=np.empty((B, n), np.float64), obj=np.empty(B, np.float64),
        labels=np.empty((B, m), np.uint8), n_active=np.empty(B, np.int32), pivots=np.empty((B, 4), np.int32),
        ties=np.empty(B, np.int32), violations=np.empty(B, np.int32))
    rc = ctx.lib.ddb_solve_label_host(ctx.handle, B, m, n, _np_ptr(A), _np_ptr(b), _np_ptr(c), float(threshold),
                                      _np_ptr(row_mask), _np_ptr(res['status']), _np_ptr(res['x']), _np_ptr(res['obj']),
                                      _np_ptr(res['labels']), _np_ptr(res['n_active']), _np_ptr(res['pivots']),
                                      _np_ptr(res['ties']), _np_ptr(res['violations']))
    _lib.check(rc, 'ddb_solve_label_host')
    return res


def generate(key, first_instance, B, m, n, density=1.0, device=0, want_x0=False):
    """Philox instance generator (throughput mode) -> A[B,m,n], b[B,m], c[B,n] (and x0[B,n]) on `device`.
    Batched replacement of RandomLPDataset._generate_problems (reference src/data/randomlp_dataset.py:58-63, 76-86)."""
    _require_cuda()
    dev = torch.device('cuda', device) if not isinstance(device, torch.device) else device
    A = torch.empty(B, m, n, dtype=torch.float64, device=dev)
    b = torch.empty(B, m, dtype=torch.float64, device=dev)
    c = torch.empty(B, n, dtype=torch.float64, device=dev)
    x0 = torch.empty(B, n, dtype=torch.float64, device=dev) if want_x0 else None
    ctx = _lib.context(dev.index)
    rc = ctx.lib.ddb_generate_dev(ctx.handle, int(key), int(first_instance), B, m, n, float(density),
                                  _ptr(A), _ptr(b), _ptr(c), _ptr(x0), _stream_ptr(dev))
    _lib.check(rc, 'ddb_generate_dev')
    return (A, b, c, x0) if want_x0 else (A, b, c)


def generate_solve_label(key, first_instance, B, m, n, density=1.0, threshold=DEFAULT_THRESHOLD, device=0,
                         keep_instances=False, out=None, instances=None):
    """Fused generate -> solve -> label in one kernel launch (the solver CTA draws its instance itself); instances are
    only materialised for the caller when keep_instances (fresh tensors) or when `instances` = (A, b, c) preallocated
    CUDA tensors are given.  Batched replacement of the loop RandomLPDataset._generate_problems -> create_lp_problem
    (reference src/data/randomlp_dataset.py:58-63, 65-128)."""
    _require_cuda()
    dev = torch.device('cuda', device) if not isinstance(device, torch.device) else device
    res = out if out is not None else _alloc_outputs(B, m, n, dev)
    A = b = c = None
    if instances is not None:
        A, b, c = instances
        for t, shp in ((A, (B, m, n)), (b, (B, m)), (c, (B, n))):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == shp):
                raise ValueError('instances must be contiguous float64 CUDA tensors A[B,m,n], b[B,m], c[B,n]')
        keep_instances = True
    elif keep_instances:
        A = torch.empty(B, m, n, dtype=torch.float64, device=dev)
        b = torch.empty(B, m, dtype=torch.float64, device=dev)
        c = torch.empty(B, n, dtype=torch.float64, device=dev)
    ctx = _lib.context(dev.index)
    rc = ctx.lib.ddb_generate_solve_label_dev(ctx.handle, int(key), int(first_instance), B, m, n, float(density),
                                              float(threshold), _ptr(res['status']), _ptr(res['x']), _ptr(res['obj']),
                                              _ptr(res['labels']), _ptr(res['n_active']), _ptr(res['pivots']),
                                              _ptr(res['ties']), _ptr(res['violations']), _ptr(A), _ptr(b), _ptr(c),
                                              _stream_ptr(dev))
    _lib.check(rc, 'ddb_generate_solve_label_dev')
    if keep_instances:
        res['A'], res['b'], res['c'] = A, b, c
    return res


def _host_outputs(B, m, n, pinned):
    if pinned:
        mk = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
        return SolveResult(
            status=mk((B,), torch.int32), x=mk((B, n), torch.float64), obj=mk((B,), torch.float64),
            labels=mk((B, m), torch.uint8), n_active=mk((B,), torch.int32), pivots=mk((B, 4), torch.int32),
            ties=mk((B,), torch.int32), violations=mk((B,), torch.int32))
    return SolveResult(
        status=np.empty(B, np.int32), x=np.empty((B, n), np.float64), obj=np.empty(B, np.float64),
        labels=np.empty((B, m), np.uint8), n_active=np.empty(B, np.int32), pivots=np.empty((B, 4), np.int32),
        ties=np.empty(B, np.int32), violations=np.empty(B, np.int32))


def generate_solve_label_host(key, first_instance, B, m, n, density=1.0, threshold=DEFAULT_THRESHOLD, device=0, out=None,
                              pinned=False):
    """Fused generate -> solve -> label with HOST (numpy) outputs: what the reference's dataset loop hands its caller.
    Nothing travels host -> device; results come back chunk by chunk while the next chunk is being solved."""
    _require_cuda()
    ctx = _lib.context(device)
    res = out if out is not None else _host_outputs(B, m, n, pinned)
    rc = ctx.lib.ddb_generate_solve_label_host(ctx.handle, int(key), int(first_instance), B, m, n, float(density),
                                               float(threshold), _np_ptr(res['status']), _np_ptr(res['x']),
                                               _np_ptr(res['obj']), _np_ptr(res['labels']), _np_ptr(res['n_active']),
                                               _np_ptr(res['pivots']), _np_ptr(res['ties']), _np_ptr(res['violations']))
    _lib.check(rc, 'ddb_generate_solve_label_host')
    return res
