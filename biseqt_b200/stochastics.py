"""Batched synthetic data with the reference's distributions
(``biseqt/stochastics.py``): ``rand_seq`` (:28-41), ``MutationProcess.mutate``
(:143-201) and ``log_odds_scores`` (:298-310), vectorised over pairs because
the reference draws one ``np.random.choice`` per base.

Two forms: :func:`generate_pairs` runs on the device (``bsq_generate_pairs``,
csrc/bsq_gen.cuh; counter-based randomness, 10^5 pairs of 10 kb in milliseconds)
and is what the bench uses at BASELINE.json's full sizes; :func:`rand_seqs` /
:func:`mutate_batch` are the host (numpy, seeded ``Generator``) lock-step form
used for small cases.  The reference itself is unseeded."""
from math import log

import numpy as np


def rand_seqs(rng, npairs, length, L=4):
    """npairs uniform random sequences, (npairs, length) uint8."""
    return rng.integers(0, L, size=(npairs, length), dtype=np.uint8)


def log_odds_scores(L, subst_prob, go_prob, ge_prob):
    """Substitution matrix and (go, ge) scores of MutationProcess.log_odds_scores
    with a uniform null hypothesis (stochastics.py:298-310)."""
    assert 0 < go_prob <= ge_prob < 1 and 0 < subst_prob < 1
    match, each = 1 - subst_prob, subst_prob / (L - 1)
    S = [[log(1 - ge_prob) + log(match if i == j else each) - log(1. / L) for j in range(L)]
         for i in range(L)]
    return np.array(S, dtype=np.float64), log(go_prob) - log(ge_prob), log(ge_prob)


def mutate_batch(rng, origins, L, subst_prob, go_prob, ge_prob, want_transcripts=False):
    """MutationProcess.mutate on every row of ``origins`` (npairs, n) at once.

    The per-sequence state machine of stochastics.py:174-200 is advanced in
    lock step: every iteration is one pass of the reference's while loop for
    all sequences that have not finished.  Returns (mutants, lengths[, ops]):
    mutants is (npairs, cap) uint8 padded with zeros.
    """
    origins = np.ascontiguousarray(origins, dtype=np.uint8)
    N, n = origins.shape
    cap = int(n * (1 + 2 * go_prob / max(1e-9, 1 - ge_prob)) + 8 * np.sqrt(n) + 64)
    out = np.zeros((N, cap), dtype=np.uint8)
    olen = np.zeros(N, dtype=np.int64)
    pos = np.zeros(N, dtype=np.int64)
    op = np.zeros(N, dtype=np.int8)       # 0 none/M/S, 1 = I, 2 = D
    ops = [] if want_transcripts else None
    rows = np.arange(N)
    while True:
        act = pos < n
        if not act.any():
            break
        u = rng.random(N)
        v = rng.random(N)
        let = rng.integers(0, L, size=N, dtype=np.uint8)
        in_gap = act & (op != 0)
        extend = in_gap & (u < ge_prob)
        stop = in_gap & ~extend
        notgap = act & (op == 0)
        opengap = notgap & (u < go_prob)
        copy = notgap & ~opengap
        new_d = opengap & (v < 0.5)
        new_i = opengap & ~new_d
        ins = (extend & (op == 1)) | new_i
        dele = (extend & (op == 2)) | new_d
        # copy with substitution: a different letter with prob subst_prob
        src = origins[rows, np.minimum(pos, n - 1)]
        sub = copy & (v < subst_prob) & (L > 1)
        other = (src + 1 + (let % max(L - 1, 1))) % L
        emit = ins | copy
        val = np.where(ins, let, np.where(sub, other, src)).astype(np.uint8)
        if emit.any():
            idx = rows[emit]
            if olen[idx].max() >= cap:      # extremely unlikely; grow
                out = np.concatenate([out, np.zeros((N, cap), np.uint8)], axis=1)
                cap *= 2
            out[idx, olen[idx]] = val[emit]
            olen[idx] += 1
        pos[dele | copy] += 1
        op[new_i] = 1
        op[new_d] = 2
        op[stop] = 0
        if want_transcripts:
            code = np.zeros(N, dtype=np.uint8)
            code[copy & ~sub] = ord('M')
            code[sub] = ord('S')
            code[ins] = ord('I')
            code[dele] = ord('D')
            ops.append(code)
    if want_transcripts:
        M = np.stack(ops, axis=1)
        txs = [bytes(r[r != 0]).decode() for r in M]
        return out, olen, txs
    return out, olen


def generate_pairs(engine, npairs, origin_len, L=4, subst_prob=0.0, go_prob=0.0, ge_prob=0.0, seed=0,
                   max_offset=0, fixed_mutant_len=0, transcripts=False, pinned=False):
    """``npairs`` synthetic (origin, mutant) pairs generated on the device
    (``bsq_generate_pairs``): origin = uniform random letters (``rand_seq``), mutant =
    ``MutationProcess(subst_probs=subst_prob, go_prob=, ge_prob=).mutate`` of the origin, or of
    the chunk window ``off ~ U[0, max_offset]`` letters further on (overlapping long reads).
    Returns a dict with ``codes, o_off, o_len, m_off, m_len, src_off`` in the layout
    :meth:`biseqt_b200.batch.Engine.align` takes (+ ``transcripts``: list of planted edit
    transcripts).  ``pinned``: ``codes`` lives in page-locked host memory."""
    from ._ffi import ffi, BsqError
    npairs = int(npairs)
    gp = ffi.new('bsq_gen_params*')
    gp.seed, gp.alphabet_len, gp.origin_len = int(seed), int(L), int(origin_len)
    gp.max_offset, gp.fixed_mutant_len = int(max_offset), int(fixed_mutant_len)
    gp.subst_prob, gp.go_prob, gp.ge_prob = float(subst_prob), float(go_prob), float(ge_prob)
    o_off, m_off = np.empty(npairs, np.int64), np.empty(npairs, np.int64)
    o_len, m_len = np.empty(npairs, np.int32), np.empty(npairs, np.int32)
    src_off = np.empty(npairs, np.int32)
    ncodes = ffi.new('int64_t*')
    if fixed_mutant_len > 0:
        est = origin_len + fixed_mutant_len
    else:   # expected mutant length n (1 + go/(1-ge) / 2 ...) with slack; retried on BSQ_ERR_CAPACITY
        est = origin_len * (2 + go_prob / max(1e-9, 1 - ge_prob)) + 8 * np.sqrt(origin_len + 1) + 64
    cap = int(npairs * est) + 16
    tx_off = np.empty(npairs + 1, np.int64) if transcripts else None
    for attempt in range(2):
        codes = engine.pinned_empty(cap, np.uint8) if pinned else np.empty(cap, np.uint8)
        tx_cap = int(npairs * (2 * est)) if transcripts else 0
        tx_buf = np.empty(max(tx_cap, 1), np.uint8) if transcripts else None
        P = lambda a, t: ffi.NULL if a is None else ffi.cast(t, a.ctypes.data)   # noqa: E731
        rc = engine.lib.bsq_generate_pairs(
            engine.ctx, gp, npairs, P(codes, 'uint8_t*'), cap, ncodes, P(o_off, 'int64_t*'), P(o_len, 'int32_t*'),
            P(m_off, 'int64_t*'), P(m_len, 'int32_t*'), P(src_off, 'int32_t*'), P(tx_off, 'int64_t*'),
            P(tx_buf, 'char*'), tx_cap)
        if rc == -4 and attempt == 0:     # BSQ_ERR_CAPACITY: *ncodes holds the size needed
            cap = int(ncodes[0])
            est = max(est, (cap / max(npairs, 1)))
            continue
        if rc != 0:
            raise BsqError('bsq_generate_pairs failed (%d): %s' % (rc, engine._err()))
        break
    out = dict(codes=codes[:int(ncodes[0])] if npairs else codes[:0], o_off=o_off, o_len=o_len, m_off=m_off,
               m_len=m_len, src_off=src_off)
    if transcripts:
        out['transcripts'] = [bytes(tx_buf[tx_off[i]:tx_off[i + 1]]).decode() for i in range(npairs)]
    return out
