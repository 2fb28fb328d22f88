"""CPU: the C-ABI library builds/loads and exports exactly the symbols include/ddb200.h declares (no compute)."""
import os
import re
import subprocess

import pytest

from deep_dantzig_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, 'include', 'ddb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ddb_[a-z_0-9]+)\s*\(', text)))


@pytest.fixture(scope='module')
def built():
    if not os.path.exists(_lib.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.library_path()


def test_header_and_binding_agree():
    assert _header_functions() == sorted(_lib.SIGNATURES)


def test_library_exports_every_symbol(built):
    out = subprocess.run(['nm', '-D', '--defined-only', built], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r'\bT (ddb_[a-z_0-9]+)', out))
    assert set(_header_functions()) <= exported


def test_library_loads_and_reports_version(built):
    assert _lib.abi_version() == 2
    lib = _lib.load()
    assert lib.ddb_last_error() is not None


def test_sass_is_sm100a_with_bulk_tma(built):
    """The shipped cubin is sm_100a and the shared-memory kernel stages its tableau with bulk TMA (UBLKCP)."""
    out = subprocess.run(['cuobjdump', '-sass', built], capture_output=True, text=True, check=True).stdout
    assert 'sm_100a' in out
    assert 'UBLKCP' in out
    assert 'DFMA' in out


def test_row_variants_are_in_the_shipped_cubin(built):
    """Every instantiation of the row-per-thread kernel the dispatch table names (rowreg_kernel.cuh: row_variants) is in the
    library, with and without the in-solver generator -- including the wide variants (six / twelve warps, 151 columns)."""
    out = subprocess.run(['cuobjdump', '-elf', built], capture_output=True, text=True, check=True).stdout
    src = open(os.path.join(os.path.dirname(built), 'csrc', 'rowreg_kernel.cuh')).read()
    table = re.findall(r'DDB_ROW_VARIANT\((\d+), (\d+), (\d+), (\d+), GEN\)', src)
    assert len(table) >= 13 and ('101', '46', '12', '1') in table and ('151', '74', '5', '1') in table
    for nc, ts, w, minb in table:
        for gen in '01':
            assert 'simplex_rowreg_kernelILi%sELi%sELi%sELi%sELb%s' % (nc, ts, w, minb, gen) in out, (nc, ts, w, minb, gen)


def test_no_cpu_fallback_without_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip('device present')
    import numpy as np
    from deep_dantzig_b200 import solver
    with pytest.raises(_lib.DdbError):
        solver.solve_label_host(np.zeros((1, 4, 2)), np.ones((1, 4)), np.ones((1, 2)))
    with pytest.raises(_lib.DdbError):
        _lib.Context(0)
