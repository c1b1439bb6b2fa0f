"""Synthetic workloads of BASELINE.json's configs:
Alphabet('ACGT'), uniform random origin (rand_seq), mutant with
MutationProcess.mutate semantics, numpy Generator(PCG64(seed)).  Shared by
bench.py and the GPU parity tests."""
import numpy as np

from biseqt_b200 import stochastics as st
from biseqt_b200 import band

STD_MODE, BANDED_MODE = 0, 1
LOCAL, GLOBAL = 1, 0
B_GLOBAL, B_LOCAL, B_OVERLAP = 0, 1, 2


def _pack_rows(o_rows, o_lens, m_rows, m_lens):
    """[origin_i | mutant_i] back to back in one uint8 array + offsets."""
    N = len(o_lens)
    o_lens = np.asarray(o_lens, dtype=np.int64)
    m_lens = np.asarray(m_lens, dtype=np.int64)
    tot = o_lens + m_lens
    base = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(tot, out=base[1:])
    codes = np.zeros(int(base[-1]) + 16, dtype=np.uint8)
    o_off = base[:-1].copy()
    m_off = o_off + o_lens
    if N and hasattr(o_rows, 'shape') and (o_lens == o_lens[0]).all() and (m_lens == m_lens[0]).all():
        no, nm = int(o_lens[0]), int(m_lens[0])
        view = codes[:N * (no + nm)].reshape(N, no + nm)
        view[:, :no] = o_rows[:, :no]
        view[:, no:] = m_rows[:, :nm]
    else:
        for i in range(N):
            codes[o_off[i]:o_off[i] + o_lens[i]] = o_rows[i][:o_lens[i]]
            codes[m_off[i]:m_off[i] + m_lens[i]] = m_rows[i][:m_lens[i]]
    return codes, o_off, o_lens.astype(np.int32), m_off, m_lens.astype(np.int32)


def int_scores(match, mismatch, L=4):
    return np.where(np.eye(L, dtype=bool), float(match), float(mismatch))


def make(kind, npairs, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    L = 4
    if kind in ('c2_int', 'c2_logodds', 'c2w_int'):
        # configs[1]: 150 bp read pairs, STD LOCAL, affine gaps (c2w_*: the same with 250 bp reads)
        n = 250 if kind == 'c2w_int' else 150
        o = st.rand_seqs(rng, npairs, n, L)
        mut, mlen = st.mutate_batch(rng, o, L, 0.02, 0.005, 0.01)
        m = np.empty((npairs, n), dtype=np.uint8)
        pad = st.rand_seqs(rng, npairs, n, L)
        w = min(mut.shape[1], n)
        m[:, :w] = mut[:, :w]
        if w < n:
            m[:, w:] = pad[:, w:]
        short = np.arange(n)[None, :] >= np.minimum(mlen, n)[:, None]   # trimmed / padded to n
        m[short] = pad[short]
        codes, o_off, o_len, m_off, m_len = _pack_rows(o, np.full(npairs, n), m, np.full(npairs, n))
        if kind in ('c2_int', 'c2w_int'):   # IgBlast-like, experiments/blot_ig_genotyping.py:351-355
            S, go, ge = int_scores(1, -3), -5.0, -2.0
        else:
            S, go, ge = st.log_odds_scores(L, 0.05, 0.05, 0.1)
        return dict(kind=kind, mode=STD_MODE, type=LOCAL, subst=S, go=go, ge=ge, codes=codes,
                    o_off=o_off, o_len=o_len, m_off=m_off, m_len=m_len,
                    desc='%d pairs %dx%d bp, STD LOCAL, %s' % (npairs, n, n, 'int +1/-3 go-5 ge-2' if kind != 'c2_logodds' else 'log-odds'))
    if kind in ('c3_int', 'c3_logodds', 'c4_int', 'c3_blocal_int', 'c3_blocal_logodds'):
        # configs[2] / [3]: long noisy reads, banded overlap alignment (c3_blocal_*: the same pairs
        # and bands as config 3 aligned B_LOCAL, the north star's "batched banded local alignment")
        if kind == 'c4_int':
            n, maxoff, gmax, sens = 50000, 25000, 0.15, 0.99
        else:
            n, maxoff = 10000, 5000
        chunk = st.rand_seqs(rng, npairs, n + maxoff, L)
        off = rng.integers(0, maxoff + 1, size=npairs)
        o = chunk[:, :n]
        src = np.stack([chunk[i, off[i]:off[i] + n] for i in range(npairs)])
        mut, mlen = st.mutate_batch(rng, src, L, 0.03, 0.06, 0.12)
        codes, o_off, o_len, m_off, m_len = _pack_rows(o, np.full(npairs, n), mut, mlen)
        if kind == 'c4_int':
            K = band.expected_overlap_len(n, mlen, off, gmax)
            r = band.band_radii(K, gmax, sens)
        else:
            r = np.full(npairs, 64)
        dmin = np.maximum(off - r, -mlen).astype(np.int32)
        dmax = np.minimum(off + r, n).astype(np.int32)
        if kind.endswith('logodds'):
            S, go, ge = st.log_odds_scores(L, 0.03, 0.06, 0.12)
        else:
            S, go, ge = int_scores(2, -3), 0.0, -2.0
        btype = B_LOCAL if 'blocal' in kind else B_OVERLAP
        return dict(kind=kind, mode=BANDED_MODE, type=btype, subst=S, go=go, ge=ge, codes=codes,
                    o_off=o_off, o_len=o_len, m_off=m_off, m_len=m_len, dmin=dmin, dmax=dmax,
                    desc='%d pairs ~%d bp, %s, radius %s, %s' % (
                        npairs, n, 'B_LOCAL' if btype == B_LOCAL else 'B_OVERLAP',
                        '64' if kind != 'c4_int' else 'band_radius(K,.15,.99)~%d' % int(np.median(r)),
                        'log-odds' if kind.endswith('logodds') else 'int +2/-3 go0 ge-2'))
    if kind in ('c5_int', 'c5_logodds', 'c5s_int', 'c5s_logodds'):
        # configs[4]: `npairs` pairs of 1 Mb x ~1 Mb, B_GLOBAL, radius 2048 (c5s_*: 64 kb, the
        # size the CPU oracle finishes in seconds).  The mutant is made piecewise (1000-letter
        # pieces mutated in lock step) because the mutation state machine is per letter.
        n = 1000000 if kind.startswith('c5_') else 65000
        piece = 1000
        rows = []
        for _ in range(npairs):
            o = st.rand_seqs(rng, n // piece, piece, L)
            mut, mlen = st.mutate_batch(rng, o, L, 0.05, 0.01, 0.02)
            rows.append((o.reshape(-1), np.concatenate([mut[i, :mlen[i]] for i in range(len(mlen))])))
        o_lens = [len(r[0]) for r in rows]
        m_lens = [len(r[1]) for r in rows]
        codes, o_off, o_len, m_off, m_len = _pack_rows([r[0] for r in rows], o_lens, [r[1] for r in rows], m_lens)
        if kind.endswith('logodds'):
            S, go, ge = st.log_odds_scores(L, 0.05, 0.01, 0.02)
        else:
            S, go, ge = int_scores(2, -3), -4.0, -2.0
        dmin = np.full(npairs, -2048, dtype=np.int32)
        dmax = np.full(npairs, 2048, dtype=np.int32)
        return dict(kind=kind, mode=BANDED_MODE, type=B_GLOBAL, subst=S, go=go, ge=ge, codes=codes,
                    o_off=o_off, o_len=o_len, m_off=m_off, m_len=m_len, dmin=dmin, dmax=dmax,
                    desc='%d pair(s) %d x ~%d bp, B_GLOBAL, radius 2048, %s' % (
                        npairs, n, n, 'log-odds' if kind.endswith('logodds') else 'int +2/-3 go-4 ge-2'))
    if kind in ('c1', 'c1n'):
        # configs[0]: one 1 kb x 1 kb STD GLOBAL pair, log-odds scores (go_score = 0); c1n: npairs of them
        n = 1000
        count = npairs if kind == 'c1n' else 1
        o = st.rand_seqs(rng, count, n, L)
        mut, mlen = st.mutate_batch(rng, o, L, 0.1, 0.1, 0.1)
        codes, o_off, o_len, m_off, m_len = _pack_rows(o, [n] * count, mut, mlen)
        S, go, ge = st.log_odds_scores(L, 0.1, 0.1, 0.1)
        return dict(kind=kind, mode=STD_MODE, type=GLOBAL, subst=S, go=go, ge=ge, codes=codes,
                    o_off=o_off, o_len=o_len, m_off=m_off, m_len=m_len,
                    desc='%d pair(s) 1 kb x ~1 kb STD GLOBAL log-odds' % count)
    raise ValueError(kind)


def make_device(eng, kind, npairs, seed):
    """The same workloads as :func:`make` for the batched configs (c2_*, c3_*, c4_*), generated on
    the device (``bsq_generate_pairs``: rand_seq + MutationProcess.mutate, csrc/bsq_gen.cuh) so
    that BASELINE.json's full sizes (10^6 / 10^5 / 10^4 pairs) take milliseconds instead of
    minutes.  Same distributions and layout, different random stream than the numpy generator."""
    L = 4
    if kind in ('c2_int', 'c2_logodds', 'c2w_int'):
        n = 250 if kind == 'c2w_int' else 150
        g = st.generate_pairs(eng, npairs, n, L, 0.02, 0.005, 0.01, seed=seed, fixed_mutant_len=n)
        if kind == 'c2_logodds':
            S, go, ge = st.log_odds_scores(L, 0.05, 0.05, 0.1)
        else:
            S, go, ge = int_scores(1, -3), -5.0, -2.0
        return dict(kind=kind, mode=STD_MODE, type=LOCAL, subst=S, go=go, ge=ge, codes=g['codes'], o_off=g['o_off'],
                    o_len=g['o_len'], m_off=g['m_off'], m_len=g['m_len'], back_to_back=True,
                    desc='%d pairs %dx%d bp, STD LOCAL, %s' % (npairs, n, n, 'int +1/-3 go-5 ge-2' if kind != 'c2_logodds' else 'log-odds'))
    if kind in ('c3_int', 'c3_logodds', 'c4_int', 'c3_blocal_int', 'c3_blocal_logodds'):
        if kind == 'c4_int':
            n, maxoff, gmax, sens = 50000, 25000, 0.15, 0.99
        else:
            n, maxoff = 10000, 5000
        g = st.generate_pairs(eng, npairs, n, L, 0.03, 0.06, 0.12, seed=seed, max_offset=maxoff)
        off, mlen = g['src_off'].astype(np.int64), g['m_len'].astype(np.int64)
        if kind == 'c4_int':
            r = band.band_radii(band.expected_overlap_len(n, mlen, off, gmax), gmax, sens)
        else:
            r = np.full(npairs, 64)
        dmin = np.maximum(off - r, -mlen).astype(np.int32)
        dmax = np.minimum(off + r, n).astype(np.int32)
        if kind.endswith('logodds'):
            S, go, ge = st.log_odds_scores(L, 0.03, 0.06, 0.12)
        else:
            S, go, ge = int_scores(2, -3), 0.0, -2.0
        btype = B_LOCAL if 'blocal' in kind else B_OVERLAP
        return dict(kind=kind, mode=BANDED_MODE, type=btype, subst=S, go=go, ge=ge, codes=g['codes'], o_off=g['o_off'],
                    o_len=g['o_len'], m_off=g['m_off'], m_len=g['m_len'], dmin=dmin, dmax=dmax, back_to_back=True,
                    desc='%d pairs ~%d bp, %s, radius %s, %s' % (
                        npairs, n, 'B_LOCAL' if btype == B_LOCAL else 'B_OVERLAP',
                        '64' if kind != 'c4_int' else 'band_radius(K,.15,.99)~%d' % int(np.median(r)),
                        'log-odds' if kind.endswith('logodds') else 'int +2/-3 go0 ge-2'))
    return make(kind, npairs, seed)


def count_cells(w):
    """DP cells the reference would fill (pw.c:59-61), vectorised closed form."""
    n = w['o_len'].astype(np.int64)
    m = w['m_len'].astype(np.int64)
    if w['mode'] == STD_MODE:
        return int(((n + 1) * (m + 1)).sum())
    from biseqt_b200 import _ffi
    lib = _ffi.load()
    return int(sum(lib.bsq_count_cells(1, int(n[i]), int(m[i]), int(w['dmin'][i]), int(w['dmax'][i]))
                   for i in range(len(n))))
