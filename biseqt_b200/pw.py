# -*- coding: utf-8 -*-
"""Drop-in for ``biseqt.pw`` (reference ``biseqt/pw.py``) on top of the
B200 alignment engine: same ``Aligner`` context manager, ``solve()``,
``traceback()``, ``table_scores()``, ``Alignment`` and the same integer
constants, but the DP fill and traceback run in ``libbsqgpu.so``.

    >>> from biseqt_b200.sequence import Alphabet
    >>> from biseqt_b200.pw import Aligner
    >>> A = Alphabet('ACGT')
    >>> S, T = A.parse('AAACGCGT'), A.parse('AACGCCTT')
    >>> with Aligner(S, T) as aligner:
    ...     score = aligner.solve()
    ...     aln = aligner.traceback()
    >>> print(aln)
    origin[0]: AAACGC-GT-
    mutant[0]: AA-CGCC-TT

One ``Aligner`` is a GPU batch of one pair; for throughput use
:class:`biseqt_b200.batch.Engine` directly.
"""

import numpy as np

from .sequence import Sequence
from . import batch as _batch
from ._ffi import CONST as _C

# alignment modes (pwlib.h:30-33)
STD_MODE = _C['BSQ_STD_MODE']
BANDED_MODE = _C['BSQ_BANDED_MODE']
# standard alignment types (pwlib.h:39-53)
GLOBAL = _C['BSQ_GLOBAL']
LOCAL = _C['BSQ_LOCAL']
START_ANCHORED = _C['BSQ_START_ANCHORED']
END_ANCHORED = _C['BSQ_END_ANCHORED']
OVERLAP = _C['BSQ_OVERLAP']
START_ANCHORED_OVERLAP = _C['BSQ_START_ANCHORED_OVERLAP']
END_ANCHORED_OVERLAP = _C['BSQ_END_ANCHORED_OVERLAP']
# banded alignment types (pwlib.h:59-65)
B_GLOBAL = _C['BSQ_B_GLOBAL']
B_LOCAL = _C['BSQ_B_LOCAL']
B_OVERLAP = _C['BSQ_B_OVERLAP']

ALN_TYPES = {
    STD_MODE: [GLOBAL, LOCAL, START_ANCHORED, END_ANCHORED, OVERLAP,
               START_ANCHORED_OVERLAP, END_ANCHORED_OVERLAP],
    BANDED_MODE: [B_GLOBAL, B_OVERLAP, B_LOCAL],
}

_engine = None
_engine_lib = None


def set_library(lib):
    """Test hook: make Aligner use an already dlopen'ed library object (the
    SIMT-emulator build).  The product never calls this."""
    global _engine, _engine_lib
    if _engine is not None:
        _engine.close()
    _engine, _engine_lib = None, lib


def get_engine():
    """The process-wide :class:`Engine` Aligner objects share (device 0)."""
    global _engine
    if _engine is None:
        _engine = _batch.Engine(0, lib=_engine_lib)
    return _engine


class Aligner(object):
    """Context that solves one pairwise alignment problem; keyword arguments,
    defaults, assertions and return conventions follow ``biseqt/pw.py:119-319``.

    ``__enter__`` validates the problem the way ``dptable_init`` does (raises
    ``Exception('Failed to initialize the DP table.')`` where the reference
    returns -1); ``solve()`` fills the table on the GPU and keeps the
    traceback pointers in HBM; ``traceback()`` walks them; ``__exit__`` frees.

    Deviations from the reference, all where its behaviour is undefined or fatal
    (INTEGRATION.md section 3):

    * ``max_new_mins > 0`` raises ``NotImplementedError`` (the reference reads
      uninitialised fields there, ``_pw_internals.c:236-243``);
    * alphabets of more than 32 letters raise ``ValueError`` here (the kernels stage the
      score table in shared memory; the reference has no limit);
    * ``|score| * (n + m) >= 2**31`` raises ``NotImplementedError`` from ``solve()`` (the
      reference's ``-INT_MAX`` clamp of gap candidates could bind there);
    * ``traceback()`` returns ``None`` for an empty optimal alignment where the reference
      fails ``assert alignment != ffi.NULL`` (``pw.py:299``), and returns the transcript
      where the reference ``exit(1)``s (a path of indels only from (0,0), ``pw.c:132-134``);
    * ``table_scores()`` re-solves the problem instead of reading the solved table and raises
      ``IndexError`` for ranges beyond the frame (the reference reads out of bounds); banded
      tables are refused like in the reference (``pw.py:280-281``).
    """

    def __init__(self, origin, mutant, **kw):
        self.min_score = kw.get('min_score', float('-inf'))
        self.alnmode = kw.get('alnmode', STD_MODE)
        self.alntype = kw.get('alntype', GLOBAL)
        assert self.alnmode in [STD_MODE, BANDED_MODE]
        assert self.alntype in ALN_TYPES[self.alnmode]

        assert isinstance(origin, Sequence) and isinstance(mutant, Sequence)
        assert origin.alphabet == mutant.alphabet
        self.origin, self.mutant = origin, mutant
        self.alphabet = origin.alphabet
        if len(self.alphabet) > 32:
            raise ValueError('alphabets of more than 32 letters are not supported by the GPU kernels '
                             '(got %d letters)' % len(self.alphabet))

        origin_range = kw.get('origin_range', (0, len(self.origin)))
        mutant_range = kw.get('mutant_range', (0, len(self.mutant)))
        assert 0 <= origin_range[0] <= origin_range[1] <= len(self.origin)
        assert 0 <= mutant_range[0] <= mutant_range[1] <= len(self.mutant)
        self.origin_range, self.mutant_range = origin_range, mutant_range

        self.go_score = kw.get('go_score', 0)
        self.ge_score = kw.get('ge_score', 0)
        L = len(self.alphabet)
        subst_scores = kw.get('subst_scores', None)
        if subst_scores is None:
            mismatch = kw.get('mismatch_score', 0)
            match = kw.get('match_score', 1)
            subst_scores = [
                [match if i == j else mismatch for i in range(L)]
                for j in range(L)
            ]
        assert isinstance(subst_scores, list) and len(subst_scores) == L
        self.subst_scores = subst_scores

        self.max_new_mins = kw.get('max_new_mins', -1)
        if self.max_new_mins > 0:
            # the reference reads uninitialised memory on this path
            # (_pw_internals.c:236-243, 282-289); refuse instead of guessing
            raise NotImplementedError('max_new_mins > 0 is not supported')
        self.diag_range = kw.get('diag_range', None)
        self.min_diag = self.max_diag = None
        if self.alnmode == BANDED_MODE:
            self.min_diag, self.max_diag = kw['diag_range']
            assert -len(mutant) <= self.min_diag <= \
                self.max_diag <= len(origin)
        self.opt = None
        self._solved = None
        self._result = None
        self._entered = False

    # frame-relative letter arrays (pwlib offsets reads by origin_range.i,
    # _pw_internals.c:226-229)
    def _frame(self):
        o = self.origin.codes[self.origin_range[0]:self.origin_range[1]]
        m = self.mutant.codes[self.mutant_range[0]:self.mutant_range[1]]
        return o, m

    def __enter__(self):
        n = self.origin_range[1] - self.origin_range[0]
        m = self.mutant_range[1] - self.mutant_range[0]
        if self.alnmode == BANDED_MODE:
            # _banded_table_init_dims, _pw_internals.c:29-49
            dmax, dmin = min(self.max_diag, n), max(self.min_diag, -m)
            dend = n - m
            if self.alntype == B_GLOBAL and \
                    (dend > dmax or dend < dmin or dmax * dmin > 0):
                raise Exception('Failed to initialize the DP table.')
            if 1 + dmax - dmin < 0:
                raise Exception('Failed to initialize the DP table.')
        get_engine()   # fail loudly here if there is no GPU / no library
        self._entered = True
        return self

    def __exit__(self, *args):
        if self._solved is not None:
            self._solved.release()
        self._solved = None
        self._entered = False

    def solve(self):
        """Fills the DP table and reports the optimal score, or None."""
        eng = get_engine()
        if self._solved is not None:
            self._solved.release()
            self._solved = None
        self._result = None
        o, m = self._frame()
        codes = np.concatenate([o, m]) if len(o) + len(m) else np.zeros(1, np.uint8)
        kw = {}
        if self.alnmode == BANDED_MODE:
            kw = dict(dmin=[self.min_diag], dmax=[self.max_diag])
        sb = eng.solve(codes, [0], [len(o)], [len(o)], [len(m)], self.alnmode,
                       self.alntype, self.subst_scores, self.go_score,
                       self.ge_score, **kw)
        self._solved = sb
        st = int(sb.status[0])
        if st == _batch.INIT_FAIL:
            raise Exception('Failed to initialize the DP table.')
        if st == _batch.UNSUPPORTED:
            raise NotImplementedError(
                'problem outside what this build supports (score range beyond +/-2^31)')
        if st != _batch.OK:
            self.opt = None
            return None
        self.opt = (int(sb.end_x[0]), int(sb.end_y[0]))
        score = float(sb.score[0])
        if score < self.min_score:
            self.opt = None
            return None
        return score

    def table_scores(self):
        """2D list of scores for i in origin_range, j in mutant_range (the
        reference indexes cells by the *absolute* ranges, pw.py:282-284)."""
        if self.alnmode != STD_MODE:
            raise NotImplementedError
        o, m = self._frame()
        H = get_engine().table_scores(o, m, self.alntype, self.subst_scores,
                                      self.go_score, self.ge_score)
        return [[float(H[i][j]) for j in range(*self.mutant_range)]
                for i in range(*self.origin_range)]

    def traceback(self):
        """Traces back the optimal alignment found by :func:`solve`."""
        if self.opt is None or self._solved is None:
            return None
        if self._result is None:
            self._result = self._solved.traceback()
        r = self._result
        st = int(r.status[0])
        transcript = r.transcript(0)
        if st == _batch.EMPTY or not transcript:
            # reference: dptable_traceback returns NULL and pw.py:299 asserts;
            # an empty alignment is reported as "none found" here
            return None
        return Alignment(self.origin, self.mutant, transcript,
                         score=float(r.score[0]),
                         origin_start=int(r.start_x[0]) + self.origin_range[0],
                         mutant_start=int(r.start_y[0]) + self.mutant_range[0])

    def calculate_score(self, alignment):
        return alignment.calculate_score(self.subst_scores, self.go_score,
                                         self.ge_score)


_PAINT = {'M': '\033[32m', 'S': '\033[31m'}    # matches green, substitutions red
_RESET = '\033[0m'


class Alignment(object):
    """What ``Aligner.traceback()`` returns: a transcript over ``MSID`` anchored at a start position
    on each sequence.  A host-side value object with the interface of the reference's
    (``biseqt/pw.py:322-572``: same constructor, same assertions, same text output -- checked against
    outputs recorded from the reference, ``tests/test_alignment_golden.py``); the batched
    equivalents of its three computations run on the device (``BatchResult.stats``).

    One rendering detail is kept on purpose: with ``colored=True`` EVERY column ends in the reset
    sequence, coloured or not, as termcolor 1.1 (the reference's dependency) emits it.
    """

    _ADVANCES = {'origin': 'MSD', 'mutant': 'MSI'}

    def __init__(self, origin, mutant, transcript, score=None,
                 origin_start=0, mutant_start=0):
        assert isinstance(origin, Sequence) and isinstance(mutant, Sequence)
        assert origin.alphabet == mutant.alphabet
        transcript = str(transcript)
        assert len(transcript) > 0 and not transcript.strip('MSID')
        assert origin_start >= 0 and mutant_start >= 0
        assert origin_start + self.projected_len(transcript, on='origin') <= len(origin)
        assert mutant_start + self.projected_len(transcript, on='mutant') <= len(mutant)
        self.alphabet = origin.alphabet
        self.origin, self.mutant = origin, mutant
        self.origin_start, self.mutant_start = origin_start, mutant_start
        self.transcript = transcript
        self.score = score

    def _key(self):
        return (self.origin, self.mutant, self.transcript, self.origin_start, self.mutant_start)

    def __eq__(self, other):
        assert isinstance(other, Alignment)
        return self._key() == other._key()

    def __ne__(self, other):
        return not self == other

    __hash__ = None

    def __str__(self):
        return self.render_term(term_width=float('inf'), margin=0, colored=False)

    @classmethod
    def projected_len(cls, transcript, on='origin'):
        """Letters of ``origin`` (``mutant``) the transcript consumes (pw.py:367-389)."""
        assert on in cls._ADVANCES
        return sum(transcript.count(op) for op in cls._ADVANCES[on])

    def calculate_score(self, subst_scores, go_score, ge_score):
        """Score of the transcript under the given scores (pw.py:391-428): substitution scores
        letter by letter, ``go_score + len * ge_score`` per maximal run of one gap kind; runs are
        accumulated in transcript order, each summed on its own, like the reference does."""
        ops = np.frombuffer(self.transcript.encode('ascii'), dtype=np.uint8)
        on_o = (ops != ord('I')).astype(np.int64)
        on_m = (ops != ord('D')).astype(np.int64)
        at_o = self.origin_start + np.cumsum(on_o) - on_o      # position on origin BEFORE each op
        at_m = self.mutant_start + np.cumsum(on_m) - on_m
        cuts = np.flatnonzero(np.diff(ops)) + 1                 # run boundaries
        total = 0.
        for a, b in zip(np.concatenate(([0], cuts)), np.concatenate((cuts, [len(ops)]))):
            if chr(ops[a]) in 'MS':
                run = 0
                for k in range(a, b):
                    run += subst_scores[self.origin[int(at_o[k])]][self.mutant[int(at_m[k])]]
                total += run
            else:
                total += go_score + ge_score * int(b - a)
        return total

    def truncate_to_match(self):
        """The alignment between its first and its last match (pw.py:430-448); ``None`` when fewer
        than two matches are left.  (No match at all: ``IndexError``, as in the reference.)"""
        first, last = self.transcript.find('M'), self.transcript.rfind('M')
        if first < 0:
            raise IndexError('no match in transcript')
        if first >= last:
            return None
        skipped = self.transcript[:first]
        return Alignment(self.origin, self.mutant, self.transcript[first:last + 1],
                         origin_start=self.origin_start + self.projected_len(skipped, on='origin'),
                         mutant_start=self.mutant_start + self.projected_len(skipped, on='mutant'))

    def _columns(self, margin):
        """The rendered columns in order: ``(op or None, position on origin, position on mutant)``;
        ``None`` columns are the context shown around the alignment."""
        width = self.alphabet._letlen
        i, j = self.origin_start, self.mutant_start
        lead = min(margin, max(i, j) * width)
        for k in range(lead):
            yield None, i - lead + k, j - lead + k
        for op in self.transcript:
            yield op, i, j
            i += op != 'I'
            j += op != 'D'
        trail = min(margin, max(len(self.origin) - i, len(self.mutant) - j) * width)
        for k in range(trail):
            yield None, i + k, j + k

    def render_term(self, term_width=120, margin=0, colored=True):
        """Two-line text rendering, wrapped at ``term_width`` (pw.py:450-572): every line pair
        opens with the positions it starts at, gaps show as ``-``, context beyond either sequence
        as ``.``; a column is only ever broken off BEFORE it is written, when the pair already
        reaches ``term_width``."""
        assert term_width >= 30
        assert margin >= 0
        width = self.alphabet._letlen
        seqs = (self.origin, self.mutant)
        shown = ('MSD', 'MSI')
        pairs = []                   # finished line pairs
        rows, used = None, 0         # the open pair: two lists of pieces, visible length so far

        def close():
            texts = [''.join(r) for r in rows]
            size = max(len(t) for t in texts)
            pairs.append('%s\n%s\n' % tuple(t.rjust(size) for t in texts))

        for op, i, j in self._columns(margin):
            if rows is None or used >= term_width:
                if rows is not None:
                    close()
                rows = [['origin[%d]: ' % i], ['mutant[%d]: ' % j]]
                used = max(len(rows[0][0]), len(rows[1][0]))
                assert used <= term_width, \
                    'Alignment preamble does not fit in width %d' % term_width
            for row, seq, at, ops in zip(rows, seqs, (i, j), shown):
                if op is None:
                    text = self.alphabet[seq[at]] if 0 <= at < len(seq) else '.' * width
                else:
                    text = self.alphabet[seq[at]] if op in ops else '-' * width
                if colored:
                    text = _PAINT.get(op, '') + text + _RESET
                row.append(text)
            used += width
        if rows is None:             # cannot happen: a transcript is never empty
            rows = [['origin[%d]: ' % self.origin_start], ['mutant[%d]: ' % self.mutant_start]]
        close()
        return ''.join(pairs)
