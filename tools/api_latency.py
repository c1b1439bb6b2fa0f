"""Latency of the drop-in one-pair-at-a-time API (Aligner.solve + traceback), the way the
reference's own tests and experiments call it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biseqt_b200.sequence import Alphabet, Sequence
from biseqt_b200.pw import Aligner, LOCAL, GLOBAL, STD_MODE

A = Alphabet('ACGT')
rng = np.random.default_rng(1)
for n in (100, 1000):
    S = Sequence(A, [int(c) for c in rng.integers(0, 4, n)])
    T = Sequence(A, [int(c) for c in rng.integers(0, 4, n)])
    for _ in range(3):
        with Aligner(S, T, alntype=LOCAL, match_score=1, mismatch_score=-1, go_score=-2, ge_score=-1) as al:
            al.solve(); al.traceback()
    t0 = time.perf_counter()
    K = 50
    for _ in range(K):
        with Aligner(S, T, alntype=LOCAL, match_score=1, mismatch_score=-1, go_score=-2, ge_score=-1) as al:
            al.solve(); al.traceback()
    print('%d x %d LOCAL: %.3f ms per Aligner (solve + traceback)' % (n, n, 1e3 * (time.perf_counter() - t0) / K))
    # the reference's default alignment type (configs[0]): big GLOBAL pairs take the strip pipeline (K3L)
    for _ in range(3):
        with Aligner(S, T, alntype=GLOBAL, match_score=1, mismatch_score=-1, go_score=-2, ge_score=-1) as al:
            al.solve(); al.traceback()
    t0 = time.perf_counter()
    for _ in range(K):
        with Aligner(S, T, alntype=GLOBAL, match_score=1, mismatch_score=-1, go_score=-2, ge_score=-1) as al:
            al.solve(); al.traceback()
    print('%d x %d GLOBAL: %.3f ms per Aligner (solve + traceback)' % (n, n, 1e3 * (time.perf_counter() - t0) / K))
