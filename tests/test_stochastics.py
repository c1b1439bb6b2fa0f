"""biseqt_b200.stochastics: the device-side batched generator (bsq_generate_pairs) against its
CPU restatement byte for byte, and both generators against the distribution laws the
reference's own tests check (tests/test_stochastics.py:73-117 of the reference: a mutant's
transcript applies to the origin, substitution / gap rates match the process parameters,
log_odds_scores signs).  The CPU run drives the SIMT-emulator build; -m gpu the real library."""
import numpy as np
import pytest

from biseqt_b200 import batch, stochastics as st
from oracle import gen as gen_oracle


def _apply(origin, mutant, tx):
    """Check that transcript tx turns origin into mutant (reference test_stochastics.py:88-100)."""
    i = j = 0
    for op in tx:
        if op in 'MS':
            assert (origin[i] == mutant[j]) == (op == 'M')
            i += 1
            j += 1
        elif op == 'I':
            j += 1
        else:
            assert op == 'D'
            i += 1
    return i, j


def _check_generator(eng, npairs):
    cases = [dict(origin_len=150, subst_prob=0.02, go_prob=0.005, ge_prob=0.01, fixed_mutant_len=150),
             dict(origin_len=400, subst_prob=0.03, go_prob=0.06, ge_prob=0.12, max_offset=200),
             dict(origin_len=97, subst_prob=0.3, go_prob=0.2, ge_prob=0.5, L=3),
             dict(origin_len=60, subst_prob=0.0, go_prob=0.0, ge_prob=0.0, L=1),
             dict(origin_len=0, subst_prob=0.1, go_prob=0.1, ge_prob=0.2)]
    for ci, kw in enumerate(cases):
        L = kw.get('L', 4)
        g = st.generate_pairs(eng, npairs, seed=77 + ci, transcripts=True, **kw)
        n = kw['origin_len']
        assert (g['o_len'] == n).all() and len(g['codes']) == int(g['m_off'][-1] + g['m_len'][-1]) + 16 if npairs else True
        for i in range(npairs):
            o, m, tx, off = gen_oracle.gen_pair(77 + ci, i, L, n, kw.get('max_offset', 0), kw.get('fixed_mutant_len', 0),
                                                kw['subst_prob'], kw['go_prob'], kw['ge_prob'])
            go = g['codes'][g['o_off'][i]:g['o_off'][i] + g['o_len'][i]]
            gm = g['codes'][g['m_off'][i]:g['m_off'][i] + g['m_len'][i]]
            assert np.array_equal(go, o) and np.array_equal(gm, m), (ci, i)
            assert g['transcripts'][i] == tx and int(g['src_off'][i]) == off, (ci, i)
            if not kw.get('max_offset') and not kw.get('fixed_mutant_len'):
                assert _apply(o, m, tx) == (len(o), len(m))
    # a pair does not depend on the batch it is generated in
    a = st.generate_pairs(eng, 5, 80, subst_prob=0.1, go_prob=0.05, ge_prob=0.2, seed=3)
    b = st.generate_pairs(eng, 9, 80, subst_prob=0.1, go_prob=0.05, ge_prob=0.2, seed=3)
    assert np.array_equal(a['codes'][:a['m_off'][4] + a['m_len'][4]], b['codes'][:b['m_off'][4] + b['m_len'][4]])
    with pytest.raises(Exception):
        st.generate_pairs(eng, 3, 50, subst_prob=0.1, go_prob=0.3, ge_prob=0.2)   # go > ge (stochastics.py:140)


def test_emu_generator_matches_cpu_restatement(emu_engine):
    _check_generator(emu_engine, 7)


@pytest.mark.gpu
def test_gpu_generator_matches_cpu_restatement(gpu_engine):
    _check_generator(gpu_engine, 200)


def _rates(txs):
    s = ''.join(txs)
    n_copy = s.count('M') + s.count('S')
    gaps = sum(1 for tx in txs for k in range(len(tx)) if tx[k] in 'ID' and (k == 0 or tx[k - 1] != tx[k]))
    gap_ops = s.count('I') + s.count('D')
    return s.count('S') / max(n_copy, 1), gaps / max(n_copy + gaps, 1), 1 - gaps / max(gap_ops, 1)


def test_generator_distribution_laws():
    """Rates recovered from the planted transcripts match the process parameters (the reference
    checks the same on its own mutate(), tests/test_stochastics.py:101-117), for the CPU
    restatement of the device generator and for the host numpy generator."""
    subst, go, ge = 0.1, 0.05, 0.3
    txs = [gen_oracle.gen_pair(5, i, 4, 3000, 0, 0, subst, go, ge)[2] for i in range(40)]
    s, g, e = _rates(txs)
    assert abs(s - subst) < 0.01 and abs(g - go) < 0.006 and abs(e - ge) < 0.03
    rng = np.random.default_rng(4)
    o = st.rand_seqs(rng, 40, 3000, 4)
    mut, mlen, txs = st.mutate_batch(rng, o, 4, subst, go, ge, want_transcripts=True)
    s, g, e = _rates(txs)
    assert abs(s - subst) < 0.01 and abs(g - go) < 0.006 and abs(e - ge) < 0.03
    for i in range(40):
        assert _apply(o[i], mut[i, :mlen[i]], txs[i]) == (3000, int(mlen[i]))
    # letters are uniform (rand_seq, stochastics.py:28-41)
    letters = np.concatenate([gen_oracle.gen_pair(9, i, 4, 5000, 0, 0, 0.0, 0.0, 0.0)[0] for i in range(8)])
    assert np.abs(np.bincount(letters, minlength=4) / len(letters) - 0.25).max() < 0.01


def test_log_odds_scores():
    """stochastics.py:298-310 / reference tests: matches score positive, substitutions and gaps
    negative, go = log(go_prob) - log(ge_prob) (0 for a linear model)."""
    S, go, ge = st.log_odds_scores(4, 0.1, 0.05, 0.2)
    assert (np.diag(S) > 0).all() and (S[~np.eye(4, dtype=bool)] < 0).all()
    assert go < 0 and ge < 0 and abs(go - (np.log(0.05) - np.log(0.2))) < 1e-12
    assert st.log_odds_scores(4, 0.1, 0.1, 0.1)[1] == 0.0
