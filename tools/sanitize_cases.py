"""A small slice of the GPU parity suite sized for compute-sanitizer (10-100x slowdown):

    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py
    compute-sanitizer --tool synccheck python tools/sanitize_cases.py

Covers the kernels whose memory protocols are the delicate ones: K3's fence-free boundary-slab
hand-over between strips (cross-CTA, global memory), K1h's deferred pointer stores, the K1
multi-pass boundary column, K2's three phases, and the three traceback walkers.  Every result is
still checked against the oracle, so a sanitizer-clean run is also a parity run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from biseqt_b200 import batch   # noqa: E402
from tests import parity_cases as pc   # noqa: E402


def main():
    mk = lambda: batch.Engine(0)   # noqa: E731
    pc.run_forced_k3(mk, seed=95, count=12, wide_len=560)
    pc.run_duo_short_reads(mk, seed=87, npairs=9)
    eng = mk()
    try:
        pc.run_batches(eng, seed=21, npairs=6, lo=0, hi=60)
        pc.run_multipass_std(eng, seed=31, n_hi=40, widths=(520, 1100))
        pc.run_banded_interior(eng, seed=43, length=400, npairs=3)
        pc.run_very_wide_bands(eng, seed=91, length=1150, radius=540, npairs=1)
        pc.run_edge_cases(eng)
        pc.run_stats(eng, seed=97, npairs=20, hi=120)
    finally:
        eng.close()
    for which in ('warp', 'narrow', 'thread'):
        os.environ['BSQ_TRACE_KERNEL'] = which
        eng = mk()
        try:
            pc.run_batches(eng, seed=23, npairs=5, lo=0, hi=120)
            pc.run_banded_interior(eng, seed=47, length=700, npairs=2)
        finally:
            eng.close()
            del os.environ['BSQ_TRACE_KERNEL']
    print('sanitize_cases: all parity checks passed')


if __name__ == '__main__':
    main()
