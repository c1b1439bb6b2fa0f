"""TEST INFRASTRUCTURE -- cffi driver for oracle/lean_oracle.c (the CPU
restatement of pwlib in compact-state form).  Only tests/, smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module."""
import os
import subprocess
import numpy as np
from cffi import FFI

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liblean_oracle.so')

OK, NO_ALIGNMENT, EMPTY, INIT_FAIL, REF_WOULD_ABORT = 0, 1, 2, 3, 4

_ffi = FFI()
_ffi.cdef("""
int lean_align(int mode, int type, int L, const double *subst, double go,
               double ge, const uint8_t *origin, int n, const uint8_t *mutant,
               int m, int dmin, int dmax, double *score, int *end_x, int *end_y,
               int *start_x, int *start_y, char *transcript, int *tx_len,
               int64_t *ncells_out, double *hout);
int64_t lean_count_cells(int mode, int n, int m, int dmin, int dmax);
int lean_align_batch(int mode, int type, int L, const double *subst, double go,
                     double ge, int64_t npairs, const uint8_t *codes,
                     const int64_t *o_off, const int32_t *o_len,
                     const int64_t *m_off, const int32_t *m_len,
                     const int32_t *dmin, const int32_t *dmax, double *score,
                     int32_t *end_x, int32_t *end_y, int32_t *start_x,
                     int32_t *start_y, int32_t *status, int32_t *tx_len,
                     const int64_t *tx_off, char *tx_buf);
""")
_lib = None


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE, 'lean'])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = _ffi.dlopen(_SO)
    return _lib


def _u8(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint8))


def align(origin, mutant, subst, go, ge, mode=0, type=0, diag_range=None,
          want_table=False):
    """One problem on frame-relative letter arrays.  Returns a dict with
    status, score, end=(x,y), start=(x,y) (frame relative), transcript,
    ncells, and optionally table (STD only, (n+1)x(m+1))."""
    L = lib()
    o, t = _u8(origin), _u8(mutant)
    n, m = len(o), len(t)
    S = np.ascontiguousarray(np.asarray(subst, dtype=np.float64))
    nl = S.shape[0]
    dmin, dmax = (0, 0) if diag_range is None else diag_range
    score = _ffi.new('double*')
    ints = [_ffi.new('int*') for _ in range(5)]
    nc = _ffi.new('int64_t*')
    tx = _ffi.new('char[]', n + m + 2)
    table = None
    hptr = _ffi.NULL
    if want_table and mode == 0:
        table = np.zeros((n + 1, m + 1), dtype=np.float64)
        hptr = _ffi.cast('double*', table.ctypes.data)
    st = L.lean_align(mode, type, nl, _ffi.cast('double*', S.ctypes.data),
                      float(go), float(ge),
                      _ffi.cast('uint8_t*', o.ctypes.data), n,
                      _ffi.cast('uint8_t*', t.ctypes.data), m,
                      int(dmin), int(dmax), score, ints[0], ints[1], ints[2],
                      ints[3], tx, ints[4], nc, hptr)
    res = dict(status=st, score=None, end=None, start=None, transcript=None,
               ncells=nc[0], table=table)
    if st in (OK, EMPTY, REF_WOULD_ABORT):
        res['score'] = score[0]
        res['end'] = (ints[0][0], ints[1][0])
        res['start'] = (ints[2][0], ints[3][0])
        res['transcript'] = _ffi.string(tx).decode()
    return res


def count_cells(mode, n, m, dmin=0, dmax=0):
    return int(lib().lean_count_cells(mode, n, m, dmin, dmax))


def align_batch(mode, type, subst, go, ge, codes, o_off, o_len, m_off, m_len,
                dmin=None, dmax=None):
    """Packed batch (same layout as the C ABI of include/bsqgpu.h)."""
    L = lib()
    S = np.ascontiguousarray(np.asarray(subst, dtype=np.float64))
    codes = _u8(codes)
    o_off = np.ascontiguousarray(o_off, dtype=np.int64)
    m_off = np.ascontiguousarray(m_off, dtype=np.int64)
    o_len = np.ascontiguousarray(o_len, dtype=np.int32)
    m_len = np.ascontiguousarray(m_len, dtype=np.int32)
    N = len(o_len)
    cap = o_len.astype(np.int64) + m_len.astype(np.int64)
    tx_off = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(cap, out=tx_off[1:])
    tx_buf = np.zeros(max(int(tx_off[-1]), 1), dtype=np.uint8)
    out = dict(score=np.zeros(N), end_x=np.zeros(N, np.int32),
               end_y=np.zeros(N, np.int32), start_x=np.zeros(N, np.int32),
               start_y=np.zeros(N, np.int32), status=np.zeros(N, np.int32),
               tx_len=np.zeros(N, np.int32))

    def p(a, t):
        return _ffi.NULL if a is None else _ffi.cast(t, a.ctypes.data)
    if dmin is not None:
        dmin = np.ascontiguousarray(dmin, dtype=np.int32)
        dmax = np.ascontiguousarray(dmax, dtype=np.int32)
    L.lean_align_batch(mode, type, S.shape[0], p(S, 'double*'), float(go),
                       float(ge), N, p(codes, 'uint8_t*'),
                       p(o_off, 'int64_t*'), p(o_len, 'int32_t*'),
                       p(m_off, 'int64_t*'), p(m_len, 'int32_t*'),
                       p(dmin, 'int32_t*'), p(dmax, 'int32_t*'),
                       p(out['score'], 'double*'), p(out['end_x'], 'int32_t*'),
                       p(out['end_y'], 'int32_t*'),
                       p(out['start_x'], 'int32_t*'),
                       p(out['start_y'], 'int32_t*'),
                       p(out['status'], 'int32_t*'),
                       p(out['tx_len'], 'int32_t*'), p(tx_off, 'int64_t*'),
                       p(tx_buf, 'char*'))
    out['tx_off'] = tx_off
    out['tx_buf'] = tx_buf
    return out


def transcripts(out):
    """list of str from an align_batch()/solve_batch() style result."""
    buf, off, ln = out['tx_buf'], out['tx_off'], out['tx_len']
    return [bytes(buf[off[i]:off[i] + ln[i]]).decode() for i in range(len(ln))]
