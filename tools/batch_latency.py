"""End-to-end time of bsq_align_batch (host buffers in, host buffers out) over batch sizes,
uniform and ragged read lengths, STD LOCAL with integer scores."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biseqt_b200 import batch

eng = batch.Engine(0)
rng = np.random.default_rng(1)
S = np.where(np.eye(4, dtype=bool), 1.0, -3.0)
for ragged in (False, True):
    for N in (1, 100, 10000, 100000, 1000000):
        lens = rng.integers(100, 151, size=N) if ragged else np.full(N, 150)
        o_len = lens.astype(np.int32); m_len = lens.astype(np.int32)
        tot = (o_len.astype(np.int64) + m_len).cumsum()
        o_off = np.concatenate([[0], tot[:-1]]).astype(np.int64)
        m_off = o_off + o_len
        codes = rng.integers(0, 4, size=int(tot[-1]) + 16, dtype=np.uint8)
        for _ in range(2):
            r = eng.align(codes, o_off, o_len, m_off, m_len, 0, 1, S, -5.0, -2.0)
        K = 5 if N >= 100000 else 20
        t0 = time.perf_counter()
        for _ in range(K):
            r = eng.align(codes, o_off, o_len, m_off, m_len, 0, 1, S, -5.0, -2.0)
        dt = (time.perf_counter() - t0) / K
        t = eng.timing()
        cells = float(((o_len.astype(np.int64) + 1) * (m_len + 1)).sum())
        print('%-7s N=%-8d %9.3f ms  %8.1f GCUPS  plan %.3f ms  dtype %s chunks %d' % (
            'ragged' if ragged else 'uniform', N, dt * 1e3, cells / dt / 1e9, t['plan_ms'], batch.DTYPE_NAMES[t['dtype']], t['chunks']))
eng.close()
