"""Seeded random alignment-problem generators shared by the oracle and GPU
parity tests (mirrors the differential probe described in SURVEY.md App. A)."""
import numpy as np

STD_TYPES = list(range(7))       # GLOBAL .. END_ANCHORED_OVERLAP (pwlib.h:39-53)
BANDED_TYPES = [0, 1, 2]         # B_GLOBAL, B_LOCAL, B_OVERLAP (pwlib.h:59-65)


def score_family(rng, L, family, gomode):
    """family: 0 = 0/1 match/mismatch, 1 = integer +/-, 2 = irrational doubles,
    3 = full random matrix.  gomode: -1 / 0 / +1 sign of go."""
    if family == 0:
        S = np.where(np.eye(L, dtype=bool), 1.0, 0.0)
        ge = float(rng.choice([0, -1]))
        go_mag = 1.0
    elif family == 1:
        ma, mi = int(rng.integers(1, 4)), -int(rng.integers(0, 4))
        S = np.where(np.eye(L, dtype=bool), float(ma), float(mi))
        ge = -float(rng.integers(0, 3))
        go_mag = float(rng.integers(1, 6))
    elif family == 2:
        e = float(rng.choice([0.02, 0.1, 0.3]))
        g = float(rng.choice([0.05, 0.1, 0.25]))
        S = np.where(np.eye(L, dtype=bool), np.log(1 - g) + np.log(1 - e) + np.log(L),
                     np.log(1 - g) + np.log(e / max(L - 1, 1)) + np.log(L))
        ge = float(np.log(g))
        go_mag = float(-np.log(rng.choice([0.3, 0.5, 0.9])))
    else:
        S = rng.normal(size=(L, L)).round(int(rng.integers(0, 3)))
        ge = float(np.round(rng.normal(), int(rng.integers(0, 3))))
        go_mag = float(abs(np.round(rng.normal(), 1))) + 0.5
    go = 0.0 if gomode == 0 else (-go_mag if gomode < 0 else go_mag)
    return np.ascontiguousarray(S, dtype=np.float64), float(go), float(ge)


def mutate(rng, seq, L, subst=0.1, go=0.05, ge=0.2):
    """MutationProcess.mutate semantics (stochastics.py:143-201) on an int
    array; returns (mutant, transcript)."""
    out, tx = [], []
    pos, op = 0, ''
    n = len(seq)
    while pos < n:
        if op in ('I', 'D'):
            if rng.random() < ge:
                if op == 'I':
                    out.append(int(rng.integers(L)))
                else:
                    pos += 1
            else:
                op = ''
        else:
            if rng.random() < go:
                if rng.random() < 0.5:
                    op = 'D'
                    pos += 1
                else:
                    op = 'I'
                    out.append(int(rng.integers(L)))
            else:
                if rng.random() < subst and L > 1:
                    c = (int(seq[pos]) + 1 + int(rng.integers(L - 1))) % L
                else:
                    c = int(seq[pos])
                out.append(c)
                op = 'M' if c == seq[pos] else 'S'
                pos += 1
        tx.append(op)
    return np.array(out, dtype=np.uint8), ''.join(tx)


def random_problem(rng, max_len=13, L=None, min_len=0):
    """One random small problem covering every mode/type/score family."""
    L = L or int(rng.choice([1, 2, 4, 4, 4, 7]))
    n = int(rng.integers(min_len, max_len + 1))
    o = rng.integers(0, L, size=n).astype(np.uint8)
    if rng.random() < 0.6:
        t, _ = mutate(rng, o, L, subst=float(rng.choice([0.05, 0.2])),
                      go=float(rng.choice([0.05, 0.15])), ge=0.3)
        if len(t) > max_len:
            t = t[:max_len]
        if len(t) < min_len:
            t = np.concatenate([t, rng.integers(0, L, size=min_len - len(t)).astype(np.uint8)])
    else:
        t = rng.integers(0, L, size=int(rng.integers(min_len, max_len + 1))).astype(np.uint8)
    m = len(t)
    S, go, ge = score_family(rng, L, int(rng.integers(0, 4)), int(rng.integers(-1, 2)))
    mode = int(rng.integers(0, 2))
    if mode == 0:
        typ, dr = int(rng.choice(STD_TYPES)), None
    else:
        typ = int(rng.choice(BANDED_TYPES))
        # possibly out of range on purpose: the reference clips to [-m, n]
        a = int(rng.integers(-m - 2, n + 3))
        b = int(rng.integers(-m - 2, n + 3))
        dr = (min(a, b), max(a, b))
        if typ == 0 and rng.random() < 0.7:   # make B_GLOBAL feasible most of the time
            dr = (min(dr[0], 0, n - m), max(dr[1], 0, n - m))
    return dict(origin=o, mutant=t, subst=S, go=go, ge=ge, mode=mode, type=typ,
                diag_range=dr, L=L)
