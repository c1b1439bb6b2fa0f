"""Batched synthetic data with the reference's distributions
(``biseqt/stochastics.py``): ``rand_seq`` (:28-41), ``MutationProcess.mutate``
(:143-201) and ``log_odds_scores`` (:298-310), vectorised over pairs because
the reference draws one ``np.random.choice`` per base.  Seeded
``numpy.random.Generator``; the reference itself is unseeded."""
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
