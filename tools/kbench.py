"""Per-step kernel timings of one workload, device resident (A/B tool for kernel work).
usage: python tools/kbench.py c3_logodds 4096 [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import bench_workloads as bw  # noqa: E402
from biseqt_b200 import batch  # noqa: E402


def main():
    kind, npairs = sys.argv[1], int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    trace = (len(sys.argv) <= 4) or sys.argv[4] != 'notrace'
    from biseqt_b200 import _ffi
    eng = batch.Engine(0, lib=_ffi.load(os.environ['BSQ_LIB']) if os.environ.get('BSQ_LIB') else None)   # A/B builds
    w = bw.make_device(eng, kind, npairs, seed=3)
    rb = eng.upload(w['codes'], w['o_off'], w['o_len'], w['m_off'], w['m_len'], w['mode'], w['type'],
                    w['subst'], w['go'], w['ge'], dmin=w.get('dmin'), dmax=w.get('dmax'))
    fills, tots = [], []
    for _ in range(steps):
        t = rb.run(traceback=trace)
        fills.append(t['fill_ms'])
        tots.append(t['total_ms'])
    cells = t['cells']
    print('%s %d pairs cells %.3e dtype %s chunks %d' % (kind, npairs, cells, batch.DTYPE_NAMES[t['dtype']], t['chunks']))
    print('  fill_ms ', ' '.join('%.3f' % v for v in fills), ' -> best %.1f GCUPS' % (cells / min(fills) / 1e6))
    print('  total_ms', ' '.join('%.3f' % v for v in tots), ' -> best %.1f GCUPS' % (cells / min(tots) / 1e6))
    rb.release()
    eng.close()


if __name__ == '__main__':
    main()
