"""The C-ABI library loads without a GPU and exports every symbol that
include/bsqgpu.h declares; host-only entry points behave."""
import os
import re

import numpy as np
import pytest

from biseqt_b200 import _ffi
from biseqt_b200._ffi import ffi
from oracle import lean

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = re.sub(r'/\*.*?\*/', '', open(os.path.join(ROOT, 'include', 'bsqgpu.h')).read(), flags=re.S)
    return sorted(set(re.findall(r'\b(bsq_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    if not os.path.exists(_ffi.LIBRARY):
        g.build()
    return _ffi.load()


def test_exports_every_declared_symbol(lib):
    names = _declared_functions()
    assert len(names) >= 14
    for n in names:
        assert getattr(lib, n) is not None, n


def test_version_and_no_gpu_behaviour(lib):
    assert b'sm_100a' in ffi.string(lib.bsq_version())
    if lib.bsq_device_count() == 0:
        # no CPU fallback: creating a context must fail loudly
        assert lib.bsq_create(0) == ffi.NULL
        assert b'no CUDA device' in ffi.string(lib.bsq_last_error(ffi.NULL))


def test_count_cells(lib):
    rng = np.random.default_rng(0)
    for _ in range(500):
        n, m = int(rng.integers(0, 60)), int(rng.integers(0, 60))
        a, b = sorted(int(v) for v in rng.integers(-m - 5, n + 6, size=2))
        assert lib.bsq_count_cells(1, n, m, a, b) == lean.count_cells(1, n, m, a, b)
        assert lib.bsq_count_cells(0, n, m, 0, 0) == (n + 1) * (m + 1)


def test_partition_pairs_lpt(lib):
    rng = np.random.default_rng(1)
    N = 1000
    o_len = rng.integers(10, 500, N).astype(np.int32)
    m_len = rng.integers(10, 500, N).astype(np.int32)
    shard = np.zeros(N, np.int32)
    load = np.zeros(4, np.int64)
    rc = lib.bsq_partition_pairs(0, N, ffi.cast('int32_t*', o_len.ctypes.data),
                                 ffi.cast('int32_t*', m_len.ctypes.data), ffi.NULL, ffi.NULL, 4,
                                 ffi.cast('int32_t*', shard.ctypes.data), ffi.cast('int64_t*', load.ctypes.data))
    assert rc == 0 and set(shard) == {0, 1, 2, 3}
    cells = (o_len.astype(np.int64) + 1) * (m_len + 1)
    for k in range(4):
        assert load[k] == cells[shard == k].sum()
    assert load.max() - load.min() <= cells.max()   # LPT balance bound


def test_int32_exact_predicate(lib):
    S = np.ascontiguousarray(np.where(np.eye(4, dtype=bool), 1.0, -3.0))
    prm = ffi.new('bsq_params*')
    prm.alphabet_len, prm.go, prm.ge = 4, -5.0, -2.0
    prm.subst = ffi.cast('double*', S.ctypes.data)
    assert lib.bsq_scores_are_int32_exact(prm, 150, 150) == 1
    assert lib.bsq_scores_are_int32_exact(prm, 2 ** 30, 2 ** 30) == 0
    prm.ge = -2.5
    assert lib.bsq_scores_are_int32_exact(prm, 150, 150) == 0
