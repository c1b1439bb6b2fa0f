"""End-to-end time of bsq_align_batch (host buffers in, host buffers out) over batch sizes,
uniform and ragged read lengths, STD LOCAL with integer scores."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biseqt_b200 import batch

eng = batch.Engine(0)
rng = np.random.default_rng(1)
S = np.where(np.eye(4, dtype=bool), 1.0, -3.0)
for ragged in (False, True, 2):   # 2: origin and mutant lengths drawn independently (51 x 51 shapes)
    for N in (1, 100, 10000, 100000, 1000000):
        lens = rng.integers(100, 151, size=N) if ragged else np.full(N, 150)
        o_len = lens.astype(np.int32); m_len = (rng.integers(100, 151, size=N) if ragged == 2 else lens).astype(np.int32)
        tot = (o_len.astype(np.int64) + m_len).cumsum()
        o_off = np.concatenate([[0], tot[:-1]]).astype(np.int64)
        m_off = o_off + o_len
        codes = rng.integers(0, 4, size=int(tot[-1]) + 16, dtype=np.uint8)
        out = None
        if N >= 100000:   # big batches: pinned host buffers, the way a throughput caller would hold them
            pin = lambda a: (lambda p: (p.__setitem__(slice(None), a), p)[1])(eng.pinned_empty(len(a), a.dtype))
            codes, o_off, o_len, m_off, m_len = [pin(a) for a in (codes, o_off, o_len, m_off, m_len)]
            cap = int(o_len.sum(dtype=np.int64) + m_len.sum(dtype=np.int64))
            out = {'score': eng.pinned_empty(N, np.float64), 'tx_off': eng.pinned_empty(N + 1, np.int64),
                   'tx_buf': eng.pinned_empty(cap, np.uint8)}
            for k in ('end_x', 'end_y', 'start_x', 'start_y', 'status'):
                out[k] = eng.pinned_empty(N, np.int32)
        for _ in range(2):
            r = eng.align(codes, o_off, o_len, m_off, m_len, 0, 1, S, -5.0, -2.0, out=out)
        K = 5 if N >= 100000 else 20
        t0 = time.perf_counter()
        for _ in range(K):
            r = eng.align(codes, o_off, o_len, m_off, m_len, 0, 1, S, -5.0, -2.0, out=out)
        dt = (time.perf_counter() - t0) / K
        t = eng.timing()
        cells = float(((o_len.astype(np.int64) + 1) * (m_len + 1)).sum())
        print('%-7s N=%-8d %9.3f ms  %8.1f GCUPS  plan %.3f ms  device %.3f ms (fill %.3f)  dtype %s chunks %d%s' % (
            ('uniform', 'ragged', 'ragged2')[int(ragged)], N, dt * 1e3, cells / dt / 1e9, t['plan_ms'], t['total_ms'], t['fill_ms'],
            batch.DTYPE_NAMES[t['dtype']], t['chunks'], '  (pinned)' if out else ''))
eng.close()
