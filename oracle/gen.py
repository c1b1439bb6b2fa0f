"""TEST INFRASTRUCTURE -- cffi driver for oracle/gen_oracle.c, the CPU restatement of the
device-side pair generator (bsq_generate_pairs).  Only tests/ may import this module."""
import os
import subprocess
from math import floor

import numpy as np
from cffi import FFI

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libgen_oracle.so')
_ffi = FFI()
_ffi.cdef("""
int gen_pair(uint64_t seed, int64_t pair, int L, int n, int max_offset, int fixed_mlen, uint32_t subst_thr,
             uint32_t go_thr, uint32_t ge_thr, uint8_t *origin, uint8_t *mutant, char *tx, int *ntx, int *off_out);
""")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, 'gen_oracle.c')):
            subprocess.check_call(['make', '-s', '-C', _HERE, 'gen'])
        _lib = _ffi.dlopen(_SO)
    return _lib


def gen_pair(seed, pair, L, n, max_offset, fixed_mlen, subst_prob, go_prob, ge_prob):
    """-> (origin uint8[n], mutant uint8[], planted transcript str, source offset)."""
    cap = 4 * n + 64 + fixed_mlen
    o = np.zeros(max(n, 1), np.uint8)
    m = np.zeros(cap, np.uint8)
    tx = _ffi.new('char[]', cap)
    ntx, off = _ffi.new('int*'), _ffi.new('int*')
    thr = [int(floor(p * 16777216.0)) for p in (subst_prob, go_prob, ge_prob)]
    ln = lib().gen_pair(int(seed), int(pair), L, n, max_offset, fixed_mlen, thr[0], thr[1], thr[2],
                        _ffi.cast('uint8_t*', o.ctypes.data), _ffi.cast('uint8_t*', m.ctypes.data), tx, ntx, off)
    return o[:n].copy(), m[:ln].copy(), _ffi.buffer(tx, ntx[0])[:].decode(), off[0]
