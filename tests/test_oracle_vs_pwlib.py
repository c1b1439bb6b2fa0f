"""The lean CPU oracle against the UNMODIFIED reference pwlib (oracle/_ref/pwlib.so)
on the bench's own workload generators, at the shapes of BASELINE.json's configs
(tests/test_oracle_golden.py pins it on small random problems and the golden
vectors; this file closes the gap up to the sizes the bench and the GPU parity
tests use).  Skipped where the reference library was not built (no
/root/reference and no prebuilt oracle/_ref).

Reference functions exercised: dptable_init/solve/traceback (pw.c:10-151) with
_std_find_optimal / _banded_find_optimal (_pw_internals.c:303-414)."""
import os

import numpy as np
import pytest

import bench_workloads as bw
from oracle import lean, pwref

pytestmark = pytest.mark.skipif(not pwref.available(), reason='oracle/_ref/pwlib.so not built')


def _pair(w, i):
    o = w['codes'][w['o_off'][i]:w['o_off'][i] + w['o_len'][i]]
    t = w['codes'][w['m_off'][i]:w['m_off'][i] + w['m_len'][i]]
    dr = None if w['mode'] == 0 else (int(w['dmin'][i]), int(w['dmax'][i]))
    return o, t, dr


def _compare(w, idx, capfd=None):
    e = lean.align_batch(w['mode'], w['type'], w['subst'], w['go'], w['ge'], w['codes'], w['o_off'],
                         w['o_len'], w['m_off'], w['m_len'], w.get('dmin'), w.get('dmax'))
    etx = lean.transcripts(e)
    subst = w['subst'].tolist()
    for i in idx:
        o, t, dr = _pair(w, i)
        st = int(e['status'][i])
        b = pwref.align(o, t, subst, w['go'], w['ge'], w['mode'], w['type'], dr, will_abort=(st == 4))
        assert b['status'] == st, 'pair %d' % i
        if st in (0, 2, 4):
            assert b['score'] == e['score'][i], 'pair %d' % i       # bit-exact doubles
            assert tuple(b['end']) == (int(e['end_x'][i]), int(e['end_y'][i])), 'pair %d' % i
        if st == 0:
            assert b['transcript'] == etx[i], 'pair %d' % i
            assert tuple(b['start']) == (int(e['start_x'][i]), int(e['start_y'][i])), 'pair %d' % i
    if capfd is not None:
        capfd.readouterr()   # pwlib printf()s when it clips a band


def test_c1_shape(capfd):
    """configs[0]: 1 kb x 1 kb STD GLOBAL, log-odds doubles."""
    for seed in (1, 3):
        _compare(bw.make('c1', 1, seed=seed), [0], capfd)


@pytest.mark.parametrize('kind', ['c2_int', 'c2_logodds'])
def test_c2_shape(kind, capfd):
    """configs[1]: 150 bp read pairs, STD LOCAL, both scorings."""
    _compare(bw.make(kind, 300, seed=2), range(300), capfd)


@pytest.mark.parametrize('kind', ['c3_int', 'c3_logodds'])
def test_c3_shape(kind, capfd):
    """configs[2]: ~10 kb pairs, B_OVERLAP, radius 64."""
    _compare(bw.make(kind, 4, seed=3), range(4), capfd)


def test_c3_blocal_shape(capfd):
    """The north-star workload: batched banded LOCAL alignment (_pw_internals.c:395-408)."""
    for kind in ('c3_blocal_int', 'c3_blocal_logodds'):
        _compare(bw.make(kind, 2, seed=3), range(2), capfd)


def test_c4_shape(capfd):
    """configs[3]: one ~50 kb pair, band_radius-derived band (pwlib holds ~1.5 GB for it)."""
    _compare(bw.make('c4_int', 1, seed=4), [0], capfd)


@pytest.mark.skipif(not os.environ.get('BSQ_TEST_C5S_PWLIB'),
                    reason='scaled config 5 on true pwlib needs ~18 GB and minutes: set BSQ_TEST_C5S_PWLIB=1')
def test_c5_scaled_shape(capfd):
    """configs[4] scaled to 65 kb x 65 kb, radius 2048."""
    _compare(bw.make('c5s_int', 1, seed=6), [0], capfd)


def test_lean_table_matches_reference_table():
    """lean's H export (what Aligner.table_scores() is tested against) against
    cells[i][j].choices[0].score of the reference (pw.py:278-285), every STD type."""
    rng = np.random.default_rng(77)
    for typ in range(7):
        for fam in range(3):
            n, m = int(rng.integers(1, 60)), int(rng.integers(1, 60))
            o = rng.integers(0, 4, n).astype(np.uint8)
            t = rng.integers(0, 4, m).astype(np.uint8)
            if fam == 0:
                S, go, ge = bw.int_scores(1, -1), -2.0, -1.0
            elif fam == 1:
                S, go, ge = bw.st.log_odds_scores(4, 0.1, 0.1, 0.2)
            else:
                S, go, ge = bw.int_scores(2, -3), 1.0, -2.0   # go > 0
            a = lean.align(o, t, S, go, ge, 0, typ, want_table=True)
            P = pwref.Problem(o, t, S.tolist(), go, ge, 0, typ)
            assert P.init() == 0
            try:
                P.solve()
                ref = np.array(P.table_scores())
            finally:
                P.free()
            assert np.array_equal(a['table'], ref), (typ, fam)
