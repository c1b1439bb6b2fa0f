"""Records what the reference's host-side ``Alignment`` value class does, for tests/test_alignment_golden.py.

Run ONCE in the authoring container (needs /root/reference; the GPU box and the test suite only
read the JSON this writes):

    python tests/golden/make_alignment_golden.py

The reference class (biseqt/pw.py:322-572) is lifted out of its module with ``ast`` and executed
against this repo's ``Sequence`` / ``Alphabet`` (its module imports the cffi-built pwlib, which is
not installed here) with ``termcolor`` replaced by a three-line stand-in for the two colours it is asked for.
Nothing from the reference is copied into the repo: only inputs and the outputs it produced.
"""
import ast
import json
import os
import random
import re
import sys
from collections import namedtuple

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from biseqt_b200.sequence import Alphabet, Sequence  # noqa: E402

REF = '/root/reference/biseqt/pw.py'


class _Termcolor(object):
    """termcolor 1.1.0 (the release of the reference's time; not installed here): a colour prefix
    when a colour is given, and the reset sequence appended in every case -- which is why the
    reference strips RESET from uncoloured output (pw.py:567-570)."""
    RESET = '\033[0m'

    @staticmethod
    def colored(text, color=None, on_color=None):
        assert on_color is None
        if color is not None:
            text = '\033[%dm%s' % ({'green': 32, 'red': 31}[color], text)
        return text + _Termcolor.RESET


def reference_alignment_class():
    src = open(REF).read()
    tree = ast.parse(src)
    node = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'Alignment'][0]
    mod = ast.Module(body=[node], type_ignores=[])
    ns = {'Sequence': Sequence, 're': re, 'namedtuple': namedtuple, 'termcolor': _Termcolor}
    exec(compile(mod, REF, 'exec'), ns)
    return ns['Alignment']


def random_case(rng, letters):
    A = Alphabet(letters)
    L = len(A)
    n, m = rng.randint(1, 40), rng.randint(1, 40)
    o = [rng.randrange(L) for _ in range(n)]
    t = [rng.randrange(L) for _ in range(m)]
    os_, ms_ = rng.randrange(n), rng.randrange(m)
    i, j, ops = os_, ms_, []
    while len(ops) < 60:
        op = rng.choice('MMMSSID' if rng.random() < 0.8 else 'ID')
        if op in 'MS':
            if i >= n or j >= m:
                break
            op = 'M' if o[i] == t[j] else 'S'
            i, j = i + 1, j + 1
        elif op == 'D':
            if i >= n:
                break
            i += 1
        else:
            if j >= m:
                break
            j += 1
        ops.append(op)
        if rng.random() < 0.04:
            break
    if not ops:
        return None
    return dict(letters=letters, origin=o, mutant=t, transcript=''.join(ops), origin_start=os_, mutant_start=ms_)


def main():
    Ref = reference_alignment_class()
    rng = random.Random(20261018)
    cases = []
    while len(cases) < 160:
        c = random_case(rng, rng.choice(['ACGT', 'ACGT', 'AB', ['00', '01', '10', '11', '22']]))
        if c is None:
            continue
        A = Alphabet(c['letters'])
        S, T = Sequence(A, c['origin']), Sequence(A, c['mutant'])
        aln = Ref(S, T, c['transcript'], origin_start=c['origin_start'], mutant_start=c['mutant_start'])
        L = len(A)
        subst = [[rng.choice([-3., -2., -1.5, 1., 2., 2.25]) for _ in range(L)] for _ in range(L)]
        go, ge = rng.choice([0., -1., -4., 1.5]), rng.choice([-2., -0.5, -1.])
        c['subst'], c['go'], c['ge'] = subst, go, ge
        c['score'] = aln.calculate_score(subst, go, ge)
        c['plen'] = [Ref.projected_len(c['transcript'], on='origin'), Ref.projected_len(c['transcript'], on='mutant')]
        if 'M' in c['transcript']:
            tr = aln.truncate_to_match()
            c['truncated'] = None if tr is None else [tr.transcript, tr.origin_start, tr.mutant_start]
        else:
            c['truncated'] = 'IndexError'
        renders = []
        for width, margin, colored in ((float('inf'), 0, False), (30, 0, True), (37, 5, False), (64, 50, True), (120, 3, True)):
            renders.append(dict(width=('inf' if width == float('inf') else width), margin=margin, colored=colored,
                                text=aln.render_term(term_width=width, margin=margin, colored=colored)))
        c['renders'] = renders
        c['str'] = str(aln)
        cases.append(c)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'alignment_golden.json')
    with open(out, 'w') as f:
        json.dump(cases, f)
    print('wrote %d cases to %s' % (len(cases), out))


if __name__ == '__main__':
    main()
