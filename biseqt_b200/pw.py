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
import re
from collections import namedtuple

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


_ANSI = {'green': '\033[32m', 'red': '\033[31m'}
_ANSI_RESET = '\033[0m'


def _colored(text, color=None):
    # the reference uses termcolor.colored (pw.py:523-524); same escape codes
    if color is None:
        return text
    return _ANSI[color] + text + _ANSI_RESET


class Alignment(object):
    """A pairwise alignment (``biseqt/pw.py:322-572``): transcript over
    ``MSID`` plus start positions; pure host-side value object."""

    def __init__(self, origin, mutant, transcript, score=None,
                 origin_start=0, mutant_start=0):
        assert isinstance(origin, Sequence) and isinstance(mutant, Sequence)
        assert origin.alphabet == mutant.alphabet
        self.alphabet = origin.alphabet
        assert all(c in 'MSID' for c in transcript)
        assert len(transcript) > 0
        origin_end = origin_start + self.projected_len(transcript, on='origin')
        mutant_end = mutant_start + self.projected_len(transcript, on='mutant')
        assert 0 <= origin_start and origin_end <= len(origin)
        assert 0 <= mutant_start and mutant_end <= len(mutant)
        self.transcript = str(transcript)
        self.origin, self.mutant = origin, mutant
        self.origin_start, self.mutant_start = origin_start, mutant_start
        self.score = score

    def __str__(self):
        return self.render_term(term_width=float('+inf'), margin=0, colored=0)

    def __eq__(self, other):
        assert isinstance(other, Alignment)
        return other.origin == self.origin and \
            other.mutant == self.mutant and \
            other.transcript == self.transcript and \
            other.origin_start == self.origin_start and \
            other.mutant_start == self.mutant_start

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    @classmethod
    def projected_len(cls, transcript, on='origin'):
        assert on in ['origin', 'mutant']
        ops = 'MSD' if on == 'origin' else 'MSI'
        return sum(int(op in ops) for op in transcript)

    def calculate_score(self, subst_scores, go_score, ge_score):
        score = 0.
        i, j = self.origin_start, self.mutant_start
        for match in re.finditer(r'(.)\1*', self.transcript):
            run = match.group(0)
            op, num = run[0], len(run)
            if op in 'MS':
                score += sum(
                    subst_scores[self.origin[i + k]][self.mutant[j + k]]
                    for k in range(num)
                )
                i, j = i + num, j + num
            else:
                assert op in 'ID'
                score += go_score + ge_score * num
                if op == 'I':
                    j = j + num
                else:
                    i = i + num
        return score

    def truncate_to_match(self):
        tx_start = 0
        origin_start, mutant_start = self.origin_start, self.mutant_start
        tx_end = len(self.transcript) - 1
        while self.transcript[tx_start] != 'M':
            if self.transcript[tx_start] in 'DS':
                origin_start += 1
            if self.transcript[tx_start] in 'IS':
                mutant_start += 1
            tx_start += 1
        while self.transcript[tx_end] != 'M':
            tx_end -= 1
        if tx_start < tx_end:
            tx = self.transcript[tx_start:tx_end + 1]
            return Alignment(self.origin, self.mutant, tx,
                             origin_start=origin_start,
                             mutant_start=mutant_start)
        else:
            return None

    def render_term(self, term_width=120, margin=0, colored=True):
        assert term_width >= 30
        assert margin >= 0
        letlen = self.alphabet._letlen
        Carriage = namedtuple('carriage',
                              ['pos', 'o_idx', 'm_idx', 'o_line', 'm_line'])
        term_color = {'M': 'green', 'S': 'red'}

        def start_line(o_idx, m_idx):
            o_line = 'origin[%d]: ' % o_idx
            m_line = 'mutant[%d]: ' % m_idx
            pos = max(len(o_line), len(m_line))
            assert pos <= term_width, \
                'Alignment preamble does not fit in width %d' % term_width
            return Carriage(pos=pos, o_idx=o_idx, m_idx=m_idx,
                            o_line=o_line, m_line=m_line)

        def carriage_flush(carriage):
            line_len = max(len(carriage.o_line), len(carriage.m_line))
            return '%s\n%s\n' % (carriage.o_line.rjust(line_len),
                                 carriage.m_line.rjust(line_len))

        def carriage_fwd(carriage, op=None):
            gap = '.' * letlen if op is None else '-' * letlen
            o_contents, m_contents = gap, gap
            if op is None:
                if 0 <= carriage.o_idx < len(self.origin):
                    o_contents = self.alphabet[self.origin[carriage.o_idx]]
                if 0 <= carriage.m_idx < len(self.mutant):
                    m_contents = self.alphabet[self.mutant[carriage.m_idx]]
            else:
                assert op in 'MSID'
                if op in 'MSD':
                    o_contents = self.alphabet[self.origin[carriage.o_idx]]
                if op in 'MSI':
                    m_contents = self.alphabet[self.mutant[carriage.m_idx]]
            length = len(o_contents)
            assert length == len(m_contents)
            color = term_color.get(op) if colored else None
            o_contents = _colored(o_contents, color)
            m_contents = _colored(m_contents, color)
            output = ''
            if carriage.pos >= term_width:
                output += carriage_flush(carriage)
                carriage = start_line(carriage.o_idx, carriage.m_idx)
            return output, Carriage(
                pos=carriage.pos + length,
                o_idx=carriage.o_idx + int(op is None or op in 'MSD'),
                m_idx=carriage.m_idx + int(op is None or op in 'MSI'),
                o_line=carriage.o_line + o_contents,
                m_line=carriage.m_line + m_contents
            )

        def pre_margin(o_idx, m_idx):
            margin_len = min(margin, max(o_idx, m_idx) * letlen)
            carriage = start_line(o_idx - margin_len, m_idx - margin_len)
            output = ''
            for i in range(margin_len):
                out, carriage = carriage_fwd(carriage, op=None)
                output += out
            return output, carriage

        def post_margin(carriage):
            output = ''
            margin_len = min(margin,
                             max((len(self.origin) - carriage.o_idx) * letlen,
                                 (len(self.mutant) - carriage.m_idx) * letlen))
            for i in range(margin_len):
                out, carriage = carriage_fwd(carriage, op=None)
                output += out
            return output + carriage_flush(carriage)

        output, carriage = pre_margin(self.origin_start, self.mutant_start)
        for op in self.transcript:
            out, carriage = carriage_fwd(carriage, op=op)
            output += out
        output += post_margin(carriage)
        return output
