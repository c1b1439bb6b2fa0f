"""Generates tests/golden/pwlib_golden.json by running the UNMODIFIED reference
pwlib (oracle/_ref/pwlib.so, compiled from /root/reference/biseqt/pwlib by
oracle/Makefile) through oracle/pwref.py.  Run in the authoring container:

    make -C oracle && python tests/golden/make_golden.py

The GPU box has no /root/reference; it only reads the committed JSON.

Contents:
  * the reference's own documented / tested vectors for this path:
      - biseqt/pw.py:9-21 docstring (AAACGCGT vs AACGCCTT)
      - tests/test_pw.py:21-103 (std basic, banded basic, banded memory)
  * seeded random problems over every alnmode/alntype, score family and sign
    of go (tests/gen.py), with the reference's outputs recorded.
Problems on which the reference would PANICK/exit(1) (pw.c:132-134) are
recorded with status 4 and only score/end (traceback is not called).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lean, pwref  # noqa: E402
from tests.gen import random_problem  # noqa: E402


def run_ref(pr):
    pre = lean.align(pr['origin'], pr['mutant'], pr['subst'], pr['go'], pr['ge'],
                     pr['mode'], pr['type'], pr['diag_range'])
    r = pwref.align(pr['origin'], pr['mutant'], np.asarray(pr['subst']).tolist(), pr['go'],
                    pr['ge'], pr['mode'], pr['type'], pr['diag_range'],
                    will_abort=(pre['status'] == lean.REF_WOULD_ABORT))
    return r


def record(name, pr, r):
    return dict(name=name, origin=[int(c) for c in pr['origin']],
                mutant=[int(c) for c in pr['mutant']],
                subst=np.asarray(pr['subst'], dtype=float).tolist(), go=float(pr['go']),
                ge=float(pr['ge']), mode=int(pr['mode']), type=int(pr['type']),
                diag_range=None if pr['diag_range'] is None else [int(v) for v in pr['diag_range']],
                status=int(r['status']), score=r['score'],
                end=None if r['end'] is None else [int(v) for v in r['end']],
                start=None if r['start'] is None else [int(v) for v in r['start']],
                transcript=r['transcript'])


def eye(L, match=1.0, mismatch=0.0):
    return np.where(np.eye(L, dtype=bool), match, mismatch)


def named_cases():
    A = {'A': 0, 'C': 1, 'G': 2, 'T': 3}
    enc = lambda s: np.array([A[c] for c in s], dtype=np.uint8)  # noqa: E731
    S = np.zeros(10, np.uint8)
    junk = np.ones(10, np.uint8)
    cases = []

    def add(name, o, m, mode=0, typ=0, dr=None, go=0.0, ge=0.0, subst=None):
        cases.append((name, dict(origin=o, mutant=m, subst=eye(4) if subst is None else subst,
                                 go=go, ge=ge, mode=mode, type=typ, diag_range=dr)))
    add('pw.py:9-21 docstring', enc('AAACGCGT'), enc('AACGCCTT'))
    add('test_pw.py:33 S vs S global', S, S)
    add('test_pw.py:44 S vs S[:5]', S, S[:5])
    add('test_pw.py:55 local S+junk vs junk+S', np.concatenate([S, junk]), np.concatenate([junk, S]), typ=1)
    add('test_pw.py:60 local S vs junk (none)', S, junk, typ=1)
    add('test_pw.py:64 overlap', np.concatenate([S, junk]), np.concatenate([junk, S]), typ=4)
    add('test_pw.py:80 banded (0,0) default type', S, S, mode=1, typ=0, dr=(0, 0))
    add('test_pw.py:88 b_overlap (-20,20) ge=-1', np.concatenate([S, junk]),
        np.concatenate([junk, S]), mode=1, typ=2, dr=(-20, 20), ge=-1.0)
    # test_pw.py:95-103 at reduced length (the 1e6 version is a GPU property test)
    add('test_pw.py:95 banded memory (L=2000)', np.zeros(2000, np.uint8), np.ones(2000, np.uint8),
        mode=1, typ=0, dr=(0, 0))
    return cases


def main():
    out = []
    for name, pr in named_cases():
        out.append(record(name, pr, run_ref(pr)))
    rng = np.random.default_rng(20240607)
    for i in range(400):
        pr = random_problem(rng, max_len=24)
        out.append(record('random-%03d' % i, pr, run_ref(pr)))
    # a few larger affine/log-odds problems (multi-lane / multi-pass geometry)
    rng = np.random.default_rng(7)
    for i in range(12):
        pr = random_problem(rng, max_len=700 if i < 4 else 260, min_len=150, L=4)
        out.append(record('large-%02d' % i, pr, run_ref(pr)))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'pwlib_golden.json')
    with open(path, 'w') as f:
        json.dump(out, f, separators=(',', ':'))
    hist = {}
    for r in out:
        hist[r['status']] = hist.get(r['status'], 0) + 1
    print('wrote', path, len(out), 'cases; status histogram', hist, '%d bytes' % os.path.getsize(path))


if __name__ == '__main__':
    main()
