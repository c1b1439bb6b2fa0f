# -*- coding: utf-8 -*-
"""The reference's own tests for this path (tests/test_pw.py of biseqt),
re-expressed in Python 3 against biseqt_b200.pw.  Each test runs once on the
CPU SIMT-emulator build (-m "not gpu") and once on the real B200 (-m gpu)."""
import numpy as np
import pytest

from biseqt_b200 import pw, _ffi
from biseqt_b200.sequence import Alphabet, Sequence
from biseqt_b200.stochastics import log_odds_scores, mutate_batch
from biseqt_b200.pw import Alignment, Aligner
from biseqt_b200.pw import STD_MODE, BANDED_MODE
from biseqt_b200.pw import GLOBAL, LOCAL, OVERLAP, B_OVERLAP, B_GLOBAL, B_LOCAL
from oracle import lean


@pytest.fixture(params=['emu', pytest.param('gpu', marks=pytest.mark.gpu)])
def backend(request, emu_lib):
    pw.set_library(emu_lib if request.param == 'emu' else None)
    yield request.param
    pw.set_library(None)


def test_constants_match_pwlib():
    # pwlib.h:30-65
    assert (STD_MODE, BANDED_MODE) == (0, 1)
    assert [pw.GLOBAL, pw.LOCAL, pw.START_ANCHORED, pw.END_ANCHORED, pw.OVERLAP,
            pw.START_ANCHORED_OVERLAP, pw.END_ANCHORED_OVERLAP] == list(range(7))
    assert (B_GLOBAL, B_LOCAL, B_OVERLAP) == (0, 1, 2)
    assert pw.ALN_TYPES[BANDED_MODE] == [B_GLOBAL, B_OVERLAP, B_LOCAL]


def test_projected_aln_len():  # tests/test_pw.py:10-18
    assert Alignment.projected_len('MMM', on='origin') == 3
    assert Alignment.projected_len('MMM', on='mutant') == 3
    assert Alignment.projected_len('SMS', on='origin') == 3
    assert Alignment.projected_len('DMS', on='origin') == 3
    assert Alignment.projected_len('DMS', on='mutant') == 2
    assert Alignment.projected_len('IMS', on='origin') == 2
    assert Alignment.projected_len('IMS', on='mutant') == 3


@pytest.mark.parametrize('alphabet', [Alphabet('ACGT'), Alphabet(['00', '01'])],
                         ids=['one letter alphabet', 'two letter alphabet'])
def test_alignment_std_basic(backend, alphabet):  # tests/test_pw.py:21-67
    S = alphabet.parse(alphabet[0] * 10)
    with pytest.raises(AssertionError):
        Alignment(S, S, 'MSSST')  # illegal character
    with pytest.raises(AssertionError):
        Alignment(S, S, 'M', origin_start=len(S))  # illegal starting point
    with pytest.raises(AssertionError):
        Alignment(S, S, 'MM', origin_start=len(S) - 1)  # transcript too long

    with Aligner(S, S) as aligner:
        aligner.solve()
        assert aligner.traceback().transcript == 'M' * len(S), \
            'default mode is standard global alignment'
        scores = aligner.table_scores()
        assert len(scores) == len(S) and all([len(row) == len(S) for row in scores])
        assert max(max(row) for row in scores) == scores[-1][-1]

    with Aligner(S, S[:len(S) // 2]) as aligner:
        aligner.solve()
        alignment = aligner.traceback()
        assert alignment.transcript.count('D') == len(S) // 2
        assert '-' * (len(S) // 2) in str(alignment)

    junk = alphabet.parse(alphabet[1] * len(S))
    origin, mutant = S + junk, junk + S
    alignment = Alignment(origin, mutant, 'M' * len(S), mutant_start=len(S))
    with Aligner(origin, mutant, alntype=LOCAL) as aligner:
        aligner.solve()
        assert alignment == aligner.traceback(), 'basic local alignment works'

    with Aligner(S, junk, alntype=LOCAL) as aligner:
        assert aligner.solve() is None and aligner.traceback() is None

    with Aligner(origin, mutant, alntype=OVERLAP) as aligner:
        aligner.solve()
        assert aligner.traceback().transcript == 'M' * len(S)


@pytest.mark.parametrize('alphabet', [Alphabet('ACGT'), Alphabet(['00', '01'])],
                         ids=['one letter alphabet', 'two letter alphabet'])
def test_alignment_banded_basic(backend, alphabet):  # tests/test_pw.py:70-92
    S = alphabet.parse(alphabet[0] * 10)
    with pytest.raises(AssertionError):
        Aligner(S, S, alnmode=BANDED_MODE, diag_range=(-len(S) - 1, 0))
    with pytest.raises(AssertionError):
        Aligner(S, S, alnmode=BANDED_MODE, diag_range=(0, len(S) + 1))

    with Aligner(S, S, alnmode=BANDED_MODE, diag_range=(0, 0)) as aligner:
        aligner.solve()
        assert aligner.traceback() == Alignment(S, S, 'M' * len(S))

    junk = alphabet.parse(alphabet[1] * len(S))
    origin, mutant = S + junk, junk + S
    alignment = Alignment(origin, mutant, 'M' * len(S), mutant_start=len(S))
    with Aligner(origin, mutant, alnmode=BANDED_MODE, alntype=B_OVERLAP,
                 diag_range=(-2 * len(S), 2 * len(S)), ge_score=-1) as aligner:
        aligner.solve()
        assert alignment == aligner.traceback()


def test_alignment_banded_memory(backend):  # tests/test_pw.py:95-103
    A = Alphabet('ACGT')
    L = int(1e6) if backend == 'gpu' else 3000   # the emulator is ~1e4x slower
    S = Sequence(A, np.zeros(L, np.uint8))
    T = Sequence(A, np.ones(L, np.uint8))
    with Aligner(S, T, alnmode=BANDED_MODE, diag_range=(0, 0)) as aligner:
        aligner.solve()
        assert aligner.traceback().transcript == 'S' * L


noise_levels = [1e-2, 1e-1, 2e-1, 3e-1, 4e-1]


def _mutated_pair(err, seed, pad=0):
    A = Alphabet('ACGT')
    rng = np.random.default_rng(seed)
    S = Sequence(A, rng.integers(0, 4, 100).astype(np.uint8))
    mut, mlen, txs = mutate_batch(rng, S.codes[None, :], 4, err, err, err, want_transcripts=True)
    T = Sequence(A, mut[0, :mlen[0]].copy())
    subst, go, ge = log_odds_scores(4, err, err, err)
    return A, S, T, txs[0], subst.tolist(), go, ge


@pytest.mark.parametrize('err', noise_levels, ids=['noise=%.1e' % l for l in noise_levels])
def test_alignment_std_global(backend, err):  # tests/test_pw.py:109-140
    A, S, T, tx, subst_scores, go_score, ge_score = _mutated_pair(err, seed=int(err * 1000))
    mutation_aln = Alignment(S, T, tx)
    mutation_score = mutation_aln.calculate_score(subst_scores, go_score, ge_score)
    aligner = Aligner(S, T, subst_scores=subst_scores, go_score=go_score,
                      ge_score=ge_score, alnmode=STD_MODE, alntype=GLOBAL)
    with aligner:
        reported_score = aligner.solve()
        assert round(reported_score, 3) >= round(mutation_score, 3)
        alignment = aligner.traceback()
        aln_score = alignment.calculate_score(subst_scores, go_score, ge_score)
        assert round(aln_score, 3) == round(reported_score, 3)
        assert round(aln_score, 3) == round(aligner.calculate_score(alignment), 3)
        assert Alignment.projected_len(alignment.transcript, on='origin') == len(S)
        assert Alignment.projected_len(alignment.transcript, on='mutant') == len(T)


def test_pw_truncate_to_matches():  # tests/test_pw.py:143-154
    A = Alphabet('ACGT')
    S = A.parse('A' * 10 + 'T' * 10 + 'A' * 10)
    T = A.parse('T' * 30)
    aln = Alignment(S, T, 'S' * 10 + 'M' * 10 + 'S' * 10)
    t = aln.truncate_to_match()
    assert t.transcript == 'M' * 10 and t.origin_start == 10 and t.mutant_start == 10


@pytest.mark.parametrize('err', noise_levels, ids=['noise=%.1e' % l for l in noise_levels])
def test_alignment_std_local(backend, err):  # tests/test_pw.py:157-185
    A, S, T, tx, subst_scores, go_score, ge_score = _mutated_pair(err, seed=77 + int(err * 1000))
    mutation_score = Alignment(S, T, tx).calculate_score(subst_scores, go_score, ge_score)
    T = A.parse('A' * 100) + T + A.parse('G' * 100)
    aligner = Aligner(S, T, subst_scores=subst_scores, go_score=go_score,
                      ge_score=ge_score, alnmode=STD_MODE, alntype=LOCAL)
    with aligner:
        reported_score = aligner.solve()
        assert round(reported_score, 3) >= round(mutation_score, 3)
        alignment = aligner.traceback()
        aln_score = alignment.calculate_score(subst_scores, go_score, ge_score)
        assert round(aln_score, 3) == round(reported_score, 3)
        assert Alignment.projected_len(alignment.transcript, on='origin') <= len(S)
        assert Alignment.projected_len(alignment.transcript, on='mutant') < len(T)


def test_pw_render_basic():  # tests/test_pw.py:188-238
    A = Alphabet('ACGT')
    S = A.parse('AACT')
    aln = Alignment(S, S, 'M' * len(S))
    assert aln.render_term(colored=False).count('\033') == 0
    assert aln.render_term(colored=True).count('\033') > 0
    with pytest.raises(AssertionError):
        aln.render_term(margin=-1)
    with pytest.raises(AssertionError):
        aln.render_term(term_width=5)
    aln = Alignment(S + S, S + S, 'M' * len(S), origin_start=len(S))
    assert '[%d]' % len(S) in aln.render_term(margin=0, colored=False)
    assert '[%d]' % (len(S) - 1) in aln.render_term(margin=1, colored=False)
    full_margin = aln.render_term(margin=30, colored=False)
    assert str(S) + '.' * len(S) in full_margin
    assert len(set(len(l) for l in full_margin.rstrip().split('\n'))) == 1
    aln = Alignment(S + S, A.parse('AGT'), 'MSDM', origin_start=len(S))
    with_del = aln.render_term(colored=False)
    assert 'AG-T' in with_del
    lines = with_del.rstrip().split('\n')
    assert lines[0].index('C') == lines[1].index('-')
    aln.render_term(colored=True)
    aln = Alignment(S + S, A.parse('AACGT'), 'MMMIM', origin_start=len(S))
    with_ins = aln.render_term(colored=False)
    assert 'AAC-T' in with_ins
    lines = with_ins.rstrip().split('\n')
    assert lines[0].index('-') == lines[1].index('G')


def test_pw_render_width():  # tests/test_pw.py:241-254
    A = Alphabet('ACGT')
    N = 100
    S = A.parse('A' * (2 * N))
    aln = Alignment(S, S, 'M' * N, origin_start=N)
    render = aln.render_term(margin=2 * N, colored=False, term_width=N // 2)
    line_lens = [len(l) for l in render.rstrip().split('\n')]
    assert all(length <= N // 2 for length in line_lens)
    assert any(length == N // 2 for length in line_lens)
    assert len(set(line_lens)) <= 2


def test_pw_render_longlet():  # tests/test_pw.py:257-261
    A = Alphabet(['00', '11'])
    aln = Alignment(A.parse('0011'), A.parse('11'), 'DM')
    assert '--11' in aln.render_term(colored=False)


def test_docstring_example(backend):  # biseqt/pw.py:9-21
    A = Alphabet('ACGT')
    with Aligner(A.parse('AAACGCGT'), A.parse('AACGCCTT')) as aligner:
        assert aligner.solve() == 6.0
        aln = aligner.traceback()
    assert aln.transcript == 'MMDMMMIDMI'
    assert str(aln) == 'origin[0]: AAACGC-GT-\nmutant[0]: AA-CGCC-TT\n'


def test_subframes_and_affine(backend):
    """origin_range / mutant_range only offset reads and the reported start
    (_pw_internals.c:226-229, pw.c:146-147); checked against the oracle on slices."""
    A = Alphabet('ACGT')
    rng = np.random.default_rng(5)
    for it in range(12):
        o = rng.integers(0, 4, 60).astype(np.uint8)
        t = rng.integers(0, 4, 70).astype(np.uint8)
        t[10:50] = o[5:45]
        a0, a1 = sorted(int(v) for v in rng.integers(0, 61, 2))
        b0, b1 = sorted(int(v) for v in rng.integers(0, 71, 2))
        typ = int(rng.choice([GLOBAL, LOCAL, OVERLAP, pw.START_ANCHORED, pw.END_ANCHORED]))
        kw = dict(match_score=2, mismatch_score=-3, go_score=-5, ge_score=-2)
        e = lean.align(o[a0:a1], t[b0:b1], np.where(np.eye(4, dtype=bool), 2.0, -3.0), -5, -2, 0, typ)
        with Aligner(Sequence(A, o), Sequence(A, t), origin_range=(a0, a1), mutant_range=(b0, b1),
                     alntype=typ, **kw) as al:
            score = al.solve()
            if e['status'] == 1:
                assert score is None
                continue
            assert score == e['score']
            aln = al.traceback()
            if e['status'] == 2 or not e['transcript']:
                assert aln is None
            else:
                assert aln.transcript == e['transcript']
                assert (aln.origin_start, aln.mutant_start) == (e['start'][0] + a0, e['start'][1] + b0)


def test_init_failure_and_unsupported(backend):
    A = Alphabet('ACGT')
    S, T = A.parse('A' * 10), A.parse('A' * 4)
    with pytest.raises(Exception):   # (n,m) not inside the band: dptable_init -> -1
        with Aligner(S, T, alnmode=BANDED_MODE, alntype=B_GLOBAL, diag_range=(0, 2)) as al:
            al.solve()
    with pytest.raises(NotImplementedError):
        Aligner(S, T, max_new_mins=3)
    with pytest.raises(AssertionError):
        Aligner(S, T, alnmode=BANDED_MODE, alntype=OVERLAP, diag_range=(0, 0))   # type not in ALN_TYPES[mode]


def test_very_wide_band_through_aligner(backend):
    """A band wider than 1024 diagonals (the strip-pipelined kernel) through the drop-in API,
    against the oracle: the shape of BASELINE.json's config 5, scaled down."""
    rng = np.random.default_rng(17)
    A = Alphabet('ACGT')
    n = 1250
    o = rng.integers(0, 4, size=(1, n), dtype=np.uint8)
    mut, mlen = mutate_batch(rng, o, 4, 0.05, 0.02, 0.1)
    S = Sequence(A, [int(c) for c in o[0]])
    T = Sequence(A, [int(c) for c in mut[0, :mlen[0]]])
    subst, go, ge = log_odds_scores(4, 0.05, 0.02, 0.1)
    dr = (-560, 560)
    with Aligner(S, T, alnmode=BANDED_MODE, alntype=B_GLOBAL, subst_scores=subst.tolist(),
                 go_score=go, ge_score=ge, diag_range=dr) as al:
        score = al.solve()
        aln = al.traceback()
    e = lean.align(o[0], mut[0, :mlen[0]], subst, go, ge, 1, 0, dr)
    assert e['status'] == 0
    assert score == e['score']
    assert aln.transcript == e['transcript']
    assert (aln.origin_start, aln.mutant_start) == tuple(e['start'])
