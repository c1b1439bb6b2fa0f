"""Shared comparison helpers for the parity tests."""
import numpy as np

from biseqt_b200 import batch
from oracle import lean


def oracle_one(pr):
    return lean.align(pr['origin'], pr['mutant'], pr['subst'], pr['go'], pr['ge'],
                      pr['mode'], pr['type'], pr['diag_range'])


def engine_batch(eng, problems, **kw):
    """Run problems sharing mode/type/scores as ONE batch."""
    p0 = problems[0]
    codes, o_off, o_len, m_off, m_len = batch.pack_pairs([(p['origin'], p['mutant']) for p in problems])
    dmin = dmax = None
    if p0['mode'] == 1:
        dmin = [p['diag_range'][0] for p in problems]
        dmax = [p['diag_range'][1] for p in problems]
    return eng.align(codes, o_off, o_len, m_off, m_len, p0['mode'], p0['type'], p0['subst'],
                     p0['go'], p0['ge'], dmin=dmin, dmax=dmax, **kw)


def same_as(expected, r, i, what=''):
    """expected: dict(status, score, end, start, transcript) (oracle or golden)."""
    st = int(r.status[i])
    assert st == expected['status'], '%s: status %d != %d' % (what, st, expected['status'])
    if st in (0, 2, 4):
        # bit-exact: both sides add the same doubles in the same order
        assert float(r.score[i]) == expected['score'], '%s: score %r != %r' % (what, float(r.score[i]), expected['score'])
        assert (int(r.end_x[i]), int(r.end_y[i])) == tuple(expected['end']), what
    if st == 0:
        assert r.transcript(i) == expected['transcript'], '%s: transcript' % what
        assert (int(r.start_x[i]), int(r.start_y[i])) == tuple(expected['start']), what


def check_problems(eng, problems, **kw):
    """Each problem alone (batch of one) against the lean oracle."""
    for k, pr in enumerate(problems):
        exp = oracle_one(pr)
        r = engine_batch(eng, [pr], **kw)
        same_as(exp, r, 0, what='problem %d %r' % (k, {a: pr[a] for a in ('mode', 'type', 'go', 'ge', 'diag_range')}))


def golden_problem(g):
    return dict(origin=np.array(g['origin'], dtype=np.uint8), mutant=np.array(g['mutant'], dtype=np.uint8),
                subst=np.array(g['subst'], dtype=np.float64), go=g['go'], ge=g['ge'], mode=g['mode'],
                type=g['type'], diag_range=None if g['diag_range'] is None else tuple(g['diag_range']))


def golden_expected(g):
    """The reference reports absolute starts; golden problems use full frames so
    they coincide with the frame-relative ones the C ABI returns."""
    return dict(status=g['status'], score=g['score'], end=g['end'], start=g['start'],
                transcript=g['transcript'])
