"""The kernels' memory accesses under AddressSanitizer: the product's .cu sources compiled against
the SIMT emulator with -fsanitize=address (device buffers are plain malloc blocks there, so an
out-of-bounds read or write of any kernel is a heap-buffer-overflow report), driven through the C ABI
over a slice of the parity suite.  compute-sanitizer is closed on this GPU pool; this is the memcheck
substitute (it checks the indexing logic, not GPU timing: races are out of its reach).

    make -C tests/simt_emu asan
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
        python tools/asan_cases.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from biseqt_b200 import batch, _ffi   # noqa: E402
from tests import parity_cases as pc   # noqa: E402

lib = _ffi.load(os.path.join(ROOT, 'tests', 'simt_emu', '_build_asan', 'libbsqgpu_emu_asan.so'))
mk = lambda: batch.Engine(0, lib=lib)   # noqa: E731


def main():
    eng = mk()
    pc.run_random_singles(eng, seed=11, count=120)
    pc.run_batches(eng, seed=21, npairs=8, lo=0, hi=40)
    pc.run_multipass_std(eng, seed=31, n_hi=16, widths=(520,))
    pc.run_banded_interior(eng, seed=43, length=200, npairs=2)
    pc.run_wide_bands(eng, seed=41, length=120, radii=(40, 100, 300))
    pc.run_very_wide_bands(eng, seed=91, length=1100, radius=530)
    pc.run_wide_band_batch(eng, seed=141, lens=(300, 900, 520, 1200, 70, 800), radius=530)
    pc.run_edge_cases(eng)
    pc.run_stats(eng, seed=97, npairs=10, hi=60)
    pc.run_two_phase(eng, seed=61)
    pc.run_resident(eng, seed=65)
    eng.close()
    pc.run_forced_k3(mk, seed=95, count=40, wide_len=560)
    pc.run_duo_short_reads(mk, seed=87, npairs=9)
    pc.run_uniform_batches(mk, seed=83)
    pc.run_compact_io(mk, seed=111, npairs=9)
    pc.run_multi(mk, lambda n: batch.MultiEngine([0] * n, lib=lib), seed=121, npairs=11)
    pc.run_std_as_band(mk, seed=131, sizes=((300, 280),))
    pc.run_chunked(mk, seed=81)
    for which in ('warp', 'narrow', 'thread', 'pairthread'):
        os.environ['BSQ_TRACE_KERNEL'] = which
        e2 = mk()
        try:
            pc.run_batches(e2, seed=23, npairs=5, lo=0, hi=60)
            pc.run_banded_interior(e2, seed=47, length=200, npairs=2)
        finally:
            e2.close()
            del os.environ['BSQ_TRACE_KERNEL']
    print('asan_cases: all parity checks passed, no AddressSanitizer report')


if __name__ == '__main__':
    main()
