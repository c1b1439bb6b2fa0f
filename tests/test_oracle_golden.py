"""The lean CPU oracle against (a) the golden vectors recorded from the
compiled reference pwlib and (b), when oracle/_ref/pwlib.so is present, the
reference itself on fresh random problems."""
import os

import numpy as np
import pytest

from oracle import lean, pwref
from tests.gen import random_problem
from tests.util import golden_problem


def test_lean_matches_golden(golden):
    assert len(golden) > 400
    for g in golden:
        pr = golden_problem(g)
        a = lean.align(pr['origin'], pr['mutant'], pr['subst'], pr['go'], pr['ge'], pr['mode'],
                       pr['type'], pr['diag_range'])
        assert a['status'] == g['status'], g['name']
        if g['status'] in (0, 2, 4):
            assert a['score'] == g['score'], g['name']
            assert list(a['end']) == g['end'], g['name']
        if g['status'] == 0:
            assert a['transcript'] == g['transcript'], g['name']
            assert list(a['start']) == g['start'], g['name']


def test_reference_documented_vectors(golden):
    by = {g['name']: g for g in golden}
    d = by['pw.py:9-21 docstring']
    assert (d['transcript'], d['score']) == ('MMDMMMIDMI', 6.0)
    assert by['test_pw.py:33 S vs S global']['transcript'] == 'M' * 10
    assert by['test_pw.py:44 S vs S[:5]']['transcript'].count('D') == 5
    loc = by['test_pw.py:55 local S+junk vs junk+S']
    assert (loc['transcript'], loc['start']) == ('M' * 10, [0, 10])
    assert by['test_pw.py:60 local S vs junk (none)']['status'] == 1
    assert by['test_pw.py:64 overlap']['transcript'] == 'M' * 10
    assert by['test_pw.py:80 banded (0,0) default type']['transcript'] == 'M' * 10
    ov = by['test_pw.py:88 b_overlap (-20,20) ge=-1']
    assert (ov['transcript'], ov['start']) == ('M' * 10, [0, 10])
    assert by['test_pw.py:95 banded memory (L=2000)']['transcript'] == 'S' * 2000


@pytest.mark.skipif(not pwref.available(), reason='oracle/_ref/pwlib.so not built (no /root/reference here)')
def test_lean_matches_compiled_reference(capfd):
    rng = np.random.default_rng(12345)
    n_ok = 0
    for it in range(4000):
        pr = random_problem(rng)
        a = lean.align(pr['origin'], pr['mutant'], pr['subst'], pr['go'], pr['ge'], pr['mode'],
                       pr['type'], pr['diag_range'])
        b = pwref.align(pr['origin'], pr['mutant'], pr['subst'].tolist(), pr['go'], pr['ge'],
                        pr['mode'], pr['type'], pr['diag_range'],
                        will_abort=(a['status'] == lean.REF_WOULD_ABORT))
        assert a['status'] == b['status']
        if a['status'] in (0, 2, 4):
            assert a['score'] == b['score'] and tuple(a['end']) == tuple(b['end'])
        if a['status'] == 0:
            assert a['transcript'] == b['transcript'] and tuple(a['start']) == tuple(b['start'])
            n_ok += 1
    capfd.readouterr()  # pwlib printf()s when it clips a band
    assert n_ok > 2000


@pytest.mark.skipif(not pwref.available(), reason='oracle/_ref/pwlib.so not built')
def test_abort_prediction_is_exact():
    """status 4 <=> the reference really exit(1)s (pw.c:132-134); checked in a fork."""
    rng = np.random.default_rng(99)
    checked = 0
    for it in range(1500):
        pr = random_problem(rng)
        a = lean.align(pr['origin'], pr['mutant'], pr['subst'], pr['go'], pr['ge'], pr['mode'],
                       pr['type'], pr['diag_range'])
        if a['status'] != 4 and not (a['status'] == 0 and it % 25 == 0):
            continue
        pid = os.fork()
        if pid == 0:
            devnull = os.open(os.devnull, os.O_WRONLY)
            os.dup2(devnull, 1)
            pwref.align(pr['origin'], pr['mutant'], pr['subst'].tolist(), pr['go'], pr['ge'],
                        pr['mode'], pr['type'], pr['diag_range'])
            os._exit(0)
        _, st = os.waitpid(pid, 0)
        assert (os.WEXITSTATUS(st) == 1) == (a['status'] == 4)
        checked += 1
    assert checked > 100


def test_count_cells_matches_fill_loop():
    rng = np.random.default_rng(5)
    for _ in range(300):
        n, m = int(rng.integers(0, 40)), int(rng.integers(0, 40))
        a, b = sorted(int(v) for v in rng.integers(-m - 3, n + 4, size=2))
        r = lean.align(rng.integers(0, 4, n), rng.integers(0, 4, m), np.eye(4), 0, -1, 1, 2, (a, b))
        if r['status'] != 3:
            assert r['ncells'] == lean.count_cells(1, n, m, a, b)
