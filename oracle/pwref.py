"""TEST INFRASTRUCTURE -- py3 cffi driver for the UNMODIFIED reference pwlib.

Loads ``oracle/_ref/pwlib.so`` (compiled by ``oracle/Makefile`` from the
sources under /root/reference/biseqt/pwlib, never copied into this repo) the
same way the reference's own ``biseqt/pw.py:45-69`` does (cffi ABI mode,
``ffi.dlopen`` + ``ffi.cdef``) and builds the C problem structs exactly like
``biseqt/pw.py:203-245``.  The struct declarations below are the binding-side
restatement of the reference ABI (``pwlib.h:15-242``); nothing else of the
reference is reproduced here.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
import os
from cffi import FFI

_HERE = os.path.dirname(os.path.abspath(__file__))

# status classes, same numbering as include/bsqgpu.h
OK, NO_ALIGNMENT, EMPTY, INIT_FAIL, REF_WOULD_ABORT = 0, 1, 2, 3, 4

STD_MODE, BANDED_MODE = 0, 1

_CDEF = """
typedef struct { int i; int j; } intpair;
typedef enum { STD_MODE, BANDED_MODE } alnmode;
typedef enum { GLOBAL, LOCAL, START_ANCHORED, END_ANCHORED, OVERLAP,
               START_ANCHORED_OVERLAP, END_ANCHORED_OVERLAP } std_alntype;
typedef enum { B_GLOBAL, B_LOCAL, B_OVERLAP } banded_alntype;
typedef struct { double **subst_scores; double gap_open_score;
                 double gap_extend_score; } alnscores;
typedef struct { int* origin; int* mutant; intpair origin_range;
                 intpair mutant_range; } alnframe;
typedef struct { std_alntype type; } std_alnparams;
typedef struct { banded_alntype type; int dmin; int dmax; } banded_alnparams;
typedef struct { alnframe *frame; alnscores *scores; int max_new_mins;
                 alnmode mode;
                 union { std_alnparams* std_params;
                         banded_alnparams* banded_params; }; } alnprob;
typedef struct alnchoice { char op; double score; struct alnchoice *base;
                           int mins_cd; int cur_min; } alnchoice;
typedef struct { int num_choices; struct alnchoice *choices; } dpcell;
typedef struct { dpcell** cells; int num_rows; int* row_lens;
                 alnprob* prob; } dptable;
typedef struct alignment { int origin_idx; int mutant_idx; double score;
                           char* transcript; } alignment;
int dptable_init(dptable* T);
void dptable_free(dptable* T);
alignment* dptable_traceback(dptable* T, intpair end);
intpair dptable_solve(dptable* T);
void free(void*);
"""

_ffi = None
_libs = {}


def available(variant='O2'):
    return os.path.exists(_so_path(variant))


def _so_path(variant):
    name = 'pwlib.so' if variant == 'O2' else 'pwlib_g.so'
    return os.path.join(_HERE, '_ref', name)


def load(variant='O2'):
    """variant 'O2' = -O2 build, 'g' = the reference's own flags (-g, no -O)."""
    global _ffi
    if _ffi is None:
        _ffi = FFI()
        _ffi.cdef(_CDEF)
    if variant not in _libs:
        _libs[variant] = _ffi.dlopen(_so_path(variant))
    return _ffi, _libs[variant]


class Problem(object):
    """One reference alignment problem, structs laid out as pw.py:203-245."""

    def __init__(self, origin, mutant, subst, go, ge, mode=0, type=0,
                 diag_range=None, origin_range=None, mutant_range=None,
                 max_new_mins=-1, variant='O2'):
        ffi, lib = load(variant)
        self.ffi, self.lib = ffi, lib
        L = len(subst)
        self._rows = [ffi.new('double[]', [float(v) for v in subst[i]])
                      for i in range(L)]
        self._subst = ffi.new('double *[]', self._rows)
        self._scores = ffi.new('alnscores*', {
            'subst_scores': self._subst, 'gap_open_score': float(go),
            'gap_extend_score': float(ge)})
        o = [int(c) for c in origin]
        m = [int(c) for c in mutant]
        # ffi.new('int[]', []) would be a zero-length array: keep >= 1 element
        self._o = ffi.new('int[]', o if o else [0])
        self._m = ffi.new('int[]', m if m else [0])
        self.origin_range = tuple(origin_range or (0, len(o)))
        self.mutant_range = tuple(mutant_range or (0, len(m)))
        self._frame = ffi.new('alnframe*', {
            'origin': self._o, 'mutant': self._m,
            'origin_range': self.origin_range,
            'mutant_range': self.mutant_range})
        self.mode = mode
        if mode == STD_MODE:
            self._params = ffi.new('std_alnparams*', {'type': type})
            key = 'std_params'
        else:
            self._params = ffi.new('banded_alnparams*', {
                'type': type, 'dmin': int(diag_range[0]),
                'dmax': int(diag_range[1])})
            key = 'banded_params'
        self._prob = ffi.new('alnprob*', {
            'frame': self._frame, 'scores': self._scores, 'mode': mode,
            'max_new_mins': max_new_mins, key: self._params})
        self._table = ffi.new('dptable*', {
            'prob': self._prob, 'cells': ffi.NULL, 'num_rows': -1,
            'row_lens': ffi.NULL})
        self.inited = False

    def init(self):
        rc = self.lib.dptable_init(self._table)
        self.inited = (rc == 0)
        return rc

    def solve(self):
        opt = self.lib.dptable_solve(self._table)
        self.opt = opt
        if opt.i == -1 or opt.j == -1:
            return None
        return self._table.cells[opt.i][opt.j].choices[0].score

    def end_xy(self):
        """End cell in frame-relative (x, y); _xy_from_cellpos, _pw_internals.c:100-114."""
        i, j = self.opt.i, self.opt.j
        if self.mode == STD_MODE:
            return i, j
        d = i + self._params.dmin
        return j + max(d, 0), j - min(d, 0)

    def traceback(self):
        aln = self.lib.dptable_traceback(self._table, self.opt)
        if aln == self.ffi.NULL:
            return None
        tx = self.ffi.string(aln.transcript).decode()
        out = (aln.score, aln.origin_idx, aln.mutant_idx, tx)
        self.lib.free(aln.transcript)
        self.lib.free(aln)
        return out

    def table_scores(self):
        """choices[0].score of every STD cell, (n+1) x (m+1); pw.py:278-285."""
        n = self.origin_range[1] - self.origin_range[0]
        m = self.mutant_range[1] - self.mutant_range[0]
        T = self._table
        return [[T.cells[i][j].choices[0].score for j in range(m + 1)]
                for i in range(n + 1)]

    def masks(self):
        """4-bit tied-max mask (B=1,D=2,I=4,M=8) of every STD cell."""
        n = self.origin_range[1] - self.origin_range[0]
        m = self.mutant_range[1] - self.mutant_range[0]
        T = self._table
        bit = {b'B': 1, b'D': 2, b'I': 4, b'M': 8, b'S': 8}
        out = []
        for i in range(n + 1):
            row = []
            for j in range(m + 1):
                c = T.cells[i][j]
                row.append(sum(bit[c.choices[k].op] for k in range(c.num_choices)))
            out.append(row)
        return out

    def free(self):
        if self.inited:
            self.lib.dptable_free(self._table)
            self.inited = False


def align(origin, mutant, subst, go, ge, mode=0, type=0, diag_range=None,
          origin_range=None, mutant_range=None, will_abort=False,
          variant='O2'):
    """init + solve + traceback + free through the reference.

    Returns dict(status, score, end=(x,y) frame-relative, start=(o,m) ABSOLUTE
    as the reference reports them (pw.c:146-147), transcript).  ``will_abort``
    must be True for problems on which the reference would PANICK/exit(1)
    (pw.c:132-134) -- the traceback call is skipped for those.
    """
    P = Problem(origin, mutant, subst, go, ge, mode, type, diag_range,
                origin_range, mutant_range, variant=variant)
    res = dict(status=OK, score=None, end=None, start=None, transcript=None)
    if P.init() != 0:
        res['status'] = INIT_FAIL
        return res
    try:
        score = P.solve()
        if score is None:
            res['status'] = NO_ALIGNMENT
            return res
        res['score'] = score
        res['end'] = P.end_xy()
        if will_abort:
            res['status'] = REF_WOULD_ABORT
            return res
        tb = P.traceback()
        if tb is None:
            res['status'] = EMPTY
            return res
        res['start'] = (tb[1], tb[2])
        res['transcript'] = tb[3]
        return res
    finally:
        P.free()
