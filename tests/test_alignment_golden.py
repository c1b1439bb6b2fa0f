"""``biseqt_b200.pw.Alignment`` against outputs recorded from the reference's class
(``biseqt/pw.py:322-572``) by ``tests/golden/make_alignment_golden.py``: projected lengths, scores,
truncation to the matched core and the wrapped text rendering, 160 random alignments over three
alphabets (one with two-character letters)."""
import json
import os

import pytest

from biseqt_b200.sequence import Alphabet, Sequence
from biseqt_b200.pw import Alignment

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def cases():
    with open(os.path.join(HERE, 'golden', 'alignment_golden.json')) as f:
        return json.load(f)


def _make(c):
    A = Alphabet(c['letters'])
    return Alignment(Sequence(A, c['origin']), Sequence(A, c['mutant']), c['transcript'],
                     origin_start=c['origin_start'], mutant_start=c['mutant_start'])


def test_projected_len_and_score(cases):
    for c in cases:
        aln = _make(c)
        assert [Alignment.projected_len(c['transcript'], on='origin'),
                Alignment.projected_len(c['transcript'], on='mutant')] == c['plen']
        assert aln.calculate_score(c['subst'], c['go'], c['ge']) == c['score']   # same additions, same order


def test_truncate_to_match(cases):
    seen = set()
    for c in cases:
        aln = _make(c)
        if c['truncated'] == 'IndexError':
            with pytest.raises(IndexError):
                aln.truncate_to_match()
            seen.add('error')
            continue
        tr = aln.truncate_to_match()
        if c['truncated'] is None:
            assert tr is None
            seen.add('none')
        else:
            assert [tr.transcript, tr.origin_start, tr.mutant_start] == c['truncated']
            assert tr.transcript[0] == 'M' and tr.transcript[-1] == 'M'
            seen.add('cut')
    assert seen == {'error', 'none', 'cut'}


def test_render_term(cases):
    for c in cases:
        aln = _make(c)
        assert str(aln) == c['str']
        for r in c['renders']:
            width = float('inf') if r['width'] == 'inf' else r['width']
            assert aln.render_term(term_width=width, margin=r['margin'], colored=r['colored']) == r['text']


def test_constructor_assertions():
    A = Alphabet('ACGT')
    S, T = A.parse('ACGT'), A.parse('ACG')
    with pytest.raises(AssertionError):
        Alignment(S, T, '')
    with pytest.raises(AssertionError):
        Alignment(S, T, 'MMX')
    with pytest.raises(AssertionError):
        Alignment(S, T, 'MMMM')                 # runs past the end of the mutant
    with pytest.raises(AssertionError):
        Alignment(S, T, 'MM', origin_start=3)
    with pytest.raises(AssertionError):
        Alignment(S, 'ACG', 'MMM')
    a = Alignment(S, T, 'MMM')
    assert a == Alignment(S, T, 'MMM') and a != Alignment(S, T, 'MMD')
    with pytest.raises(AssertionError):
        a.render_term(term_width=10)
    with pytest.raises(AssertionError):
        a.render_term(margin=-1)
