"""Per-strip timeline of the K3 (strip-pipelined) kernel; needs a profiling build:
   make -C biseqt_b200/csrc OBJ=$PWD/biseqt_b200/csrc/_build_prof OUT=$PWD/biseqt_b200/libbsqgpu_prof.so EXTRA_NVFLAGS=-DBSQ_K3_PROFILE
usage: python tools/k3prof.py c5_int"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import bench_workloads as bw  # noqa: E402
from biseqt_b200 import batch, _ffi  # noqa: E402


def main():
    kind = sys.argv[1]
    _ffi.ffi.cdef('int bsq_debug_k3_res(bsq_batch* b, void* out, int64_t bytes);')
    eng = batch.Engine(0, lib=_ffi.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'biseqt_b200', os.environ.get('BSQ_PROF_LIB', 'libbsqgpu_prof.so'))))
    w = bw.make(kind, 1, seed=5)
    rb = eng.upload(w['codes'], w['o_off'], w['o_len'], w['m_off'], w['m_len'], w['mode'], w['type'],
                    w['subst'], w['go'], w['ge'], dmin=w.get('dmin'), dmax=w.get('dmax'))
    rb.run(traceback=False)
    t = rb.run(traceback=False)
    print('fill_ms', t['fill_ms'])
    dt = np.dtype([('score', 'f8'), ('k1', 'i4'), ('k2', 'i4'), ('t_claim', 'i8'), ('t_ready', 'i8'),
                   ('t_end', 'i8'), ('wait_clk', 'i8'), ('polls', 'i4'), ('sm', 'i4')])
    K = [2, 4, 8][int(os.environ.get('BSQ_K3_CFG', '1'))]
    ns = (int(w['m_len'][0]) + 32 * K) // (32 * K)
    buf = np.zeros(ns, dtype=dt)
    rc = eng.lib.bsq_debug_k3_res(rb.handle, _ffi.ffi.cast('void*', buf.ctypes.data), buf.nbytes)
    assert rc == dt.itemsize, rc
    t0 = buf['t_claim'].min()
    lo, hi = ns // 4, ns // 4 + 12
    for i in range(lo, hi):
        r = buf[i]
        print('strip %6d sm %3d claim %9.1f ready %9.1f end %9.1f us | run %7.1f us wait %8.1f us polls %5d'
              % (i, r['sm'], (r['t_claim'] - t0) / 1e3, (r['t_ready'] - t0) / 1e3, (r['t_end'] - t0) / 1e3,
                 (r['t_end'] - r['t_ready']) / 1e3, r['wait_clk'] / 1965.0, r['polls']))
    mid = buf[ns // 8: 7 * ns // 8]
    print('median ready-to-ready lag %.2f us' % np.median(np.diff(mid['t_ready']) / 1e3))
    print('median run (ready->end) %.1f us, median wait %.1f us, median polls %d, rows/strip ~%d'
          % (np.median((mid['t_end'] - mid['t_ready']) / 1e3), np.median(mid['wait_clk'] / 1965.0),
             np.median(mid['polls']), 4097 + 32 * K + 31))
    rb.release()
    eng.close()


if __name__ == '__main__':
    main()
