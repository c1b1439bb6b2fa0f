"""Kernel + scheduler logic on CPU: the product's CUDA sources compiled
against tests/simt_emu (fiber SIMT emulator) and driven through the real C ABI,
checked against the oracle.  These are NOT the parity tests proper (those are
tests/test_gpu_parity.py, -m gpu); they exist so that kernel bugs are found
without spending GPU minutes."""
import numpy as np
import pytest

from biseqt_b200 import batch
from tests import parity_cases as pc


def test_emu_golden(emu_engine, golden):
    pc.run_golden(emu_engine, golden, limit_random=150, limit_large=3)


def test_emu_random_singles(emu_engine):
    pc.run_random_singles(emu_engine, seed=11, count=250)


def test_emu_batches(emu_engine):
    pc.run_batches(emu_engine, seed=21, npairs=12, lo=0, hi=40)


def test_emu_multipass_std(emu_engine):
    pc.run_multipass_std(emu_engine, seed=31, n_hi=24, widths=(520, 1100))


def test_emu_wide_bands(emu_engine):
    pc.run_wide_bands(emu_engine, seed=41, length=120, radii=(40, 100, 300))


def test_emu_banded_interior(emu_engine):
    pc.run_banded_interior(emu_engine, seed=43, length=200, npairs=3)


def test_emu_wide_band_interior(emu_engine):
    pc.run_wide_bands(emu_engine, seed=45, length=520, radii=(460,))


@pytest.mark.parametrize('which', ['warp', 'narrow', 'pairthread'])
def test_emu_warp_traceback(emu_lib, monkeypatch, which):
    """Force the tile walker (normally only used for long pairs): a warp, or 8 lanes, per pair."""
    monkeypatch.setenv('BSQ_TRACE_KERNEL', which)
    eng = batch.Engine(0, lib=emu_lib)
    try:
        pc.run_batches(eng, seed=23, npairs=6, lo=0, hi=60)
        pc.run_multipass_std(eng, seed=33, n_hi=24, widths=(700,))
        pc.run_banded_interior(eng, seed=47, length=200, npairs=2)
        pc.run_wide_bands(eng, seed=49, length=300, radii=(100, 200))
    finally:
        eng.close()


def test_emu_very_wide_band(emu_engine):
    pc.run_very_wide_bands(emu_engine, seed=91, length=1100, radius=530)


def test_emu_wide_band_batch(emu_engine):
    pc.run_wide_band_batch(emu_engine, seed=141, lens=(300, 900, 520, 1200, 70, 800), radius=530)


def test_emu_forced_k3(emu_lib):
    pc.run_forced_k3(lambda: batch.Engine(0, lib=emu_lib), seed=95, count=120)


def test_emu_dtype_equivalence(emu_engine):
    pc.run_dtype_equivalence(emu_engine, seed=51, npairs=5, lo=5, hi=30)


def test_emu_two_phase(emu_engine):
    pc.run_two_phase(emu_engine, seed=61)


def test_emu_resident(emu_engine):
    pc.run_resident(emu_engine, seed=65)


def test_emu_edge_cases(emu_engine):
    pc.run_edge_cases(emu_engine)


def test_emu_compact_io(emu_lib):
    pc.run_compact_io(lambda: batch.Engine(0, lib=emu_lib), seed=111, npairs=13)


def test_emu_multi(emu_lib):
    pc.run_multi(lambda: batch.Engine(0, lib=emu_lib), lambda n: batch.MultiEngine([0] * n, lib=emu_lib), seed=121, npairs=17)


def test_emu_std_as_band(emu_lib):
    pc.run_std_as_band(lambda: batch.Engine(0, lib=emu_lib), seed=131, sizes=((300, 280), (256, 420)))
    pc.run_std_as_band_many(lambda: batch.Engine(0, lib=emu_lib), seed=133, npairs=12, n=520)


def test_emu_lifetime(emu_lib):
    pc.run_lifetime(lambda: batch.Engine(0, lib=emu_lib))


def test_emu_table_scores(emu_engine):
    pc.run_table_scores(emu_engine, seed=71)


def test_emu_chunked(emu_lib):
    pc.run_chunked(lambda: batch.Engine(0, lib=emu_lib), seed=81)


def test_emu_uniform_batches(emu_lib):
    pc.run_uniform_batches(lambda: batch.Engine(0, lib=emu_lib), seed=83)


def test_emu_duo_short_reads(emu_lib):
    pc.run_duo_short_reads(lambda: batch.Engine(0, lib=emu_lib), seed=87, npairs=11)


def test_emu_stats(emu_engine):
    pc.run_stats(emu_engine, seed=97)


def test_emu_big_ragged_plan(emu_lib):
    pc.run_big_ragged_plan(lambda: batch.Engine(0, lib=emu_lib), seed=99, npairs=66000, all_modes=False)


def test_emu_caller_transcript_buffer(emu_engine):
    """align(out={'tx_buf': ...}): the caller's buffer is used as it is when the transcripts fit;
    one that is too small (BSQ_ERR_CAPACITY) is replaced by a worst-case buffer."""
    rng = np.random.default_rng(5)
    N = 6
    o_len = np.full(N, 30, np.int32)
    m_len = np.full(N, 30, np.int32)
    tot = (o_len.astype(np.int64) + m_len).cumsum()
    o_off = np.concatenate([[0], tot[:-1]]).astype(np.int64)
    m_off = o_off + o_len
    codes = rng.integers(0, 4, size=int(tot[-1]) + 16, dtype=np.uint8)
    for i in range(N):
        codes[m_off[i]:m_off[i] + 30] = codes[o_off[i]:o_off[i] + 30]
    S = np.where(np.eye(4, dtype=bool), 1.0, -1.0)
    args = (codes, o_off, o_len, m_off, m_len, 0, 0, S, -1.0, -1.0)
    r0 = emu_engine.align(*args)
    small, big = {'tx_buf': np.empty(10, np.uint8)}, {'tx_buf': np.empty(1000, np.uint8)}
    r1 = emu_engine.align(*args, out=small)
    r2 = emu_engine.align(*args, out=big)
    assert r0.transcripts() == r1.transcripts() == r2.transcripts() == ['M' * 30] * N
    assert r2.tx_buf is big['tx_buf'] and r1.tx_buf is not small['tx_buf']


def test_emu_block_plan_apis(emu_lib):
    pc.run_block_plan_apis(lambda: batch.Engine(0, lib=emu_lib), seed=103)
