"""Batch entry point over the C ABI (``include/bsqgpu.h``).

The reference API is one ``Aligner`` per pair; a GPU needs whole batches, so
this module takes packed numpy buffers (all letters in one uint8 array plus
per-pair offsets/lengths) and returns arrays.  ``biseqt_b200.pw.Aligner`` is a
batch of one on top of this.
"""
import numpy as np

from . import _ffi
from ._ffi import ffi, CONST, BsqError

STATUS_NAMES = {
    CONST['BSQ_OK']: 'OK', CONST['BSQ_NO_ALIGNMENT']: 'NO_ALIGNMENT',
    CONST['BSQ_EMPTY']: 'EMPTY', CONST['BSQ_INIT_FAIL']: 'INIT_FAIL',
    CONST['BSQ_REF_WOULD_ABORT']: 'REF_WOULD_ABORT', CONST['BSQ_UNSUPPORTED']: 'UNSUPPORTED',
}
OK, NO_ALIGNMENT, EMPTY, INIT_FAIL, REF_WOULD_ABORT, UNSUPPORTED = 0, 1, 2, 3, 4, 5
DTYPE_F64, DTYPE_I32, DTYPE_AUTO = 0, 1, -1
DTYPE_I16X2 = 2
DTYPE_NAMES = {0: 'f64', 1: 'i32', 2: 'i16x2'}
ERR_CAPACITY = -4   # BSQ_ERR_CAPACITY (include/bsqgpu.h)


def _ptr(a, ctype):
    return ffi.NULL if a is None else ffi.cast(ctype, a.ctypes.data)


def _c(a, dt):
    if a is None:
        return None
    a = np.asarray(a)
    if a.dtype != dt or not a.flags['C_CONTIGUOUS']:
        a = np.ascontiguousarray(a, dtype=dt)
    return a


F_CODES_2BIT, F_TX_2BIT = CONST['BSQ_F_CODES_2BIT'], CONST['BSQ_F_TX_2BIT']
_OPS = np.frombuffer(b'MSID', dtype=np.uint8)


def pack_codes_2bit(codes):
    """One letter per byte -> four letters per byte (``BSQ_F_CODES_2BIT``): letter j in bits
    2*(j&3).. of byte j>>2.  Alphabets of at most 4 letters."""
    c = np.ascontiguousarray(codes, dtype=np.uint8)
    pad = (-len(c)) % 4
    if pad:
        c = np.concatenate([c, np.zeros(pad, np.uint8)])
    q = c.reshape(-1, 4)
    return (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)


def unpack_ops_2bit(buf, lo, hi):
    """Ops [lo, hi) of a ``BSQ_F_TX_2BIT`` transcript buffer as b'MSID' bytes."""
    if hi <= lo:
        return b''
    b = np.asarray(buf[lo >> 2:(hi + 3) >> 2], dtype=np.uint8)
    q = np.empty((len(b), 4), np.uint8)
    for k in range(4):
        q[:, k] = (b >> (2 * k)) & 3
    return _OPS[q.reshape(-1)[(lo & 3):(lo & 3) + (hi - lo)]].tobytes()


class BatchResult(object):
    """Arrays of a finished batch; transcripts are packed in ``tx_buf`` (one byte per op, or
    four ops per byte when the batch was run with ``tx_2bit``; ``tx_off`` counts ops either way)."""

    def __init__(self, score, end_x, end_y, start_x, start_y, status, tx_off, tx_buf, tx_2bit=False,
                 tx_lens=None):
        self.score, self.end_x, self.end_y = score, end_x, end_y
        self.start_x, self.start_y, self.status = start_x, start_y, status
        self.tx_off, self.tx_buf, self.tx_2bit = tx_off, tx_buf, tx_2bit
        self.tx_lens = tx_lens    # MultiEngine: transcript i = ops tx_off[i] .. tx_off[i] + tx_lens[i]

    def __len__(self):
        return len(self.score)

    def transcript(self, i):
        if self.tx_off is None:
            return None
        lo = int(self.tx_off[i])
        hi = lo + int(self.tx_lens[i]) if self.tx_lens is not None else int(self.tx_off[i + 1])
        if self.tx_2bit:
            return unpack_ops_2bit(self.tx_buf, lo, hi).decode()
        return bytes(self.tx_buf[lo:hi]).decode()

    def transcripts(self):
        return [self.transcript(i) for i in range(len(self))]

    @property
    def tx_len(self):
        if self.tx_lens is not None:
            return np.asarray(self.tx_lens, dtype=np.int32)
        return np.diff(self.tx_off).astype(np.int32)


class Engine(object):
    """One GPU context (``bsq_ctx``).  Single-owner, not thread-safe."""

    def __init__(self, device=0, lib=None):
        self.lib = lib if lib is not None else _ffi.load()
        self.ctx = self.lib.bsq_create(int(device))
        if self.ctx == ffi.NULL:
            raise BsqError('bsq_create(%d) failed: %s' % (
                device, ffi.string(self.lib.bsq_last_error(ffi.NULL)).decode()))
        self.device = device
        self._keep = None

    def close(self):
        if self.ctx is not None and self.ctx != ffi.NULL:
            self.lib.bsq_destroy(self.ctx)
        self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _err(self):
        ctx = self.ctx if self.ctx is not None else ffi.NULL   # closed engine: the thread's last error
        return ffi.string(self.lib.bsq_last_error(ctx)).decode() or 'engine is closed'

    def reload_hooks(self):
        """Re-read the BSQ_* environment hooks (they are read once, when the engine is created)."""
        self.lib.bsq_reload_hooks(self.ctx)

    def pinned_empty(self, n, dtype):
        """1-D numpy array in page-locked host memory (full-speed H2D/D2H)."""
        dtype = np.dtype(dtype)
        nbytes = int(n) * dtype.itemsize
        p = self.lib.bsq_host_alloc(max(nbytes, 1))
        if p == ffi.NULL:
            raise BsqError('bsq_host_alloc(%d) failed' % nbytes)
        p = ffi.gc(p, self.lib.bsq_host_free, size=nbytes)
        arr = np.frombuffer(ffi.buffer(p, max(nbytes, 1)), dtype=dtype, count=int(n))
        return arr

    def _params(self, mode, type, subst, go, ge, dtype=DTYPE_AUTO, max_new_mins=-1, flags=0):
        S = np.ascontiguousarray(np.asarray(subst, dtype=np.float64))
        assert S.ndim == 2 and S.shape[0] == S.shape[1], 'subst must be L x L'
        prm = ffi.new('bsq_params*')
        prm.mode, prm.type = int(mode), int(type)
        prm.alphabet_len = S.shape[0]
        prm.max_new_mins = int(max_new_mins)
        prm.go, prm.ge = float(go), float(ge)
        prm.subst = ffi.cast('double*', S.ctypes.data)
        prm.dtype = int(dtype)
        prm.flags = int(flags)
        return prm, S

    def align(self, codes, o_off, o_len, m_off, m_len, mode, type, subst, go, ge,
              dmin=None, dmax=None, dtype=DTYPE_AUTO, transcripts=True, out=None,
              max_new_mins=-1, codes_2bit=False, ncodes=None, tx_2bit=False):
        """dptable_init + solve + traceback + free for a whole batch
        (``bsq_align_batch``).  Returns :class:`BatchResult`.  ``out`` may hold
        preallocated (e.g. pinned) arrays under the result's field names.

        Compact forms of what crosses PCIe (alphabets of <= 4 letters): ``codes_2bit`` --
        ``codes`` holds four letters per byte (:func:`pack_codes_2bit`; ``ncodes`` = number of
        letters, default 4 * len(codes)); ``tx_2bit`` -- transcripts come back four ops per byte;
        ``o_off = m_off = None`` -- the pairs lie back to back in ``codes``
        (origin_0 | mutant_0 | origin_1 | ...), no offsets are uploaded."""
        codes = _c(codes, np.uint8)
        o_off, m_off = _c(o_off, np.int64), _c(m_off, np.int64)
        o_len, m_len = _c(o_len, np.int32), _c(m_len, np.int32)
        dmin, dmax = _c(dmin, np.int32), _c(dmax, np.int32)
        N = len(o_len)
        flags = (F_CODES_2BIT if codes_2bit else 0) | (F_TX_2BIT if tx_2bit else 0)
        prm, S = self._params(mode, type, subst, go, ge, dtype, max_new_mins, flags)
        nletters = int(ncodes) if ncodes is not None else (4 * len(codes) if codes_2bit else len(codes))
        out = out or {}

        def arr(name, dt, n):
            a = out.get(name)
            if a is None:
                a = np.empty(n, dtype=dt)
            assert a.dtype == dt and len(a) >= n
            return a
        score = arr('score', np.float64, N)
        end_x, end_y = arr('end_x', np.int32, N), arr('end_y', np.int32, N)
        start_x, start_y = arr('start_x', np.int32, N), arr('start_y', np.int32, N)
        status = arr('status', np.int32, N)
        tx_off = tx_buf = None
        cap = 0
        def worst_case():   # a transcript has at most n + m ops
            return int(o_len.sum(dtype=np.int64) + m_len.sum(dtype=np.int64))
        if transcripts:
            tx_off = arr('tx_off', np.int64, N + 1)
            tx_buf = out.get('tx_buf')
            if tx_buf is not None:
                # the caller's buffer as it is: summing two million lengths costs a millisecond,
                # and the library reports BSQ_ERR_CAPACITY if the transcripts do not fit
                assert tx_buf.dtype == np.uint8
                cap = len(tx_buf)
            else:
                cap = worst_case() // 4 + 64 if tx_2bit else worst_case()
                tx_buf = np.empty(max(cap, 1), dtype=np.uint8)
        total = ffi.new('int64_t*')

        def call():
            return self.lib.bsq_align_batch(
                self.ctx, prm, N, _ptr(codes, 'uint8_t*'), nletters,
                _ptr(o_off, 'int64_t*'), _ptr(o_len, 'int32_t*'),
                _ptr(m_off, 'int64_t*'), _ptr(m_len, 'int32_t*'),
                _ptr(dmin, 'int32_t*'), _ptr(dmax, 'int32_t*'),
                _ptr(score, 'double*'), _ptr(end_x, 'int32_t*'), _ptr(end_y, 'int32_t*'),
                _ptr(start_x, 'int32_t*'), _ptr(start_y, 'int32_t*'), _ptr(status, 'int32_t*'),
                _ptr(tx_off, 'int64_t*'), _ptr(tx_buf, 'char*'), cap, total)
        rc = call()
        if rc == ERR_CAPACITY and transcripts:   # caller's buffer too small: use our own
            full = worst_case() // 4 + 64 if tx_2bit else worst_case()   # (summing a million lengths costs 1 ms)
            if cap < full:
                cap = full
                tx_buf = np.empty(max(cap, 1), dtype=np.uint8)
                rc = call()
        if rc != 0:
            raise BsqError('bsq_align_batch failed (%d): %s' % (rc, self._err()))
        return BatchResult(score[:N], end_x[:N], end_y[:N], start_x[:N], start_y[:N],
                           status[:N], None if tx_off is None else tx_off[:N + 1], tx_buf, tx_2bit=tx_2bit)

    def solve(self, codes, o_off, o_len, m_off, m_len, mode, type, subst, go, ge,
              dmin=None, dmax=None, dtype=DTYPE_AUTO, max_new_mins=-1, codes_2bit=False, ncodes=None,
              tx_2bit=False):
        """dptable_init + dptable_solve; tables stay in HBM (``bsq_solve_batch``)."""
        codes = _c(codes, np.uint8)
        o_off, m_off = _c(o_off, np.int64), _c(m_off, np.int64)
        o_len, m_len = _c(o_len, np.int32), _c(m_len, np.int32)
        dmin, dmax = _c(dmin, np.int32), _c(dmax, np.int32)
        N = len(o_len)
        flags = (F_CODES_2BIT if codes_2bit else 0) | (F_TX_2BIT if tx_2bit else 0)
        prm, S = self._params(mode, type, subst, go, ge, dtype, max_new_mins, flags)
        nletters = int(ncodes) if ncodes is not None else (4 * len(codes) if codes_2bit else len(codes))
        score = np.empty(N, np.float64)
        end_x, end_y = np.empty(N, np.int32), np.empty(N, np.int32)
        status = np.empty(N, np.int32)
        hb = ffi.new('bsq_batch**')
        rc = self.lib.bsq_solve_batch(
            self.ctx, prm, N, _ptr(codes, 'uint8_t*'), nletters,
            _ptr(o_off, 'int64_t*'), _ptr(o_len, 'int32_t*'),
            _ptr(m_off, 'int64_t*'), _ptr(m_len, 'int32_t*'),
            _ptr(dmin, 'int32_t*'), _ptr(dmax, 'int32_t*'),
            _ptr(score, 'double*'), _ptr(end_x, 'int32_t*'), _ptr(end_y, 'int32_t*'),
            _ptr(status, 'int32_t*'), hb)
        if rc != 0:
            raise BsqError('bsq_solve_batch failed (%d): %s' % (rc, self._err()))
        cap = int(o_len.sum(dtype=np.int64) + m_len.sum(dtype=np.int64))
        return SolvedBatch(self, hb[0], score, end_x, end_y, status, cap, tx_2bit)

    def upload(self, codes, o_off, o_len, m_off, m_len, mode, type, subst, go, ge,
               dmin=None, dmax=None, dtype=DTYPE_AUTO, keep_tables=False, max_new_mins=-1,
               codes_2bit=False, ncodes=None, tx_2bit=False):
        """Plan + H2D once (``bsq_batch_upload``); returns a :class:`ResidentBatch`
        whose ``run()`` executes with inputs and outputs resident in HBM."""
        codes = _c(codes, np.uint8)
        o_off, m_off = _c(o_off, np.int64), _c(m_off, np.int64)
        o_len, m_len = _c(o_len, np.int32), _c(m_len, np.int32)
        dmin, dmax = _c(dmin, np.int32), _c(dmax, np.int32)
        N = len(o_len)
        flags = (F_CODES_2BIT if codes_2bit else 0) | (F_TX_2BIT if tx_2bit else 0)
        prm, S = self._params(mode, type, subst, go, ge, dtype, max_new_mins, flags)
        nletters = int(ncodes) if ncodes is not None else (4 * len(codes) if codes_2bit else len(codes))
        hb = ffi.new('bsq_batch**')
        rc = self.lib.bsq_batch_upload(
            self.ctx, prm, N, _ptr(codes, 'uint8_t*'), nletters,
            _ptr(o_off, 'int64_t*'), _ptr(o_len, 'int32_t*'),
            _ptr(m_off, 'int64_t*'), _ptr(m_len, 'int32_t*'),
            _ptr(dmin, 'int32_t*'), _ptr(dmax, 'int32_t*'), int(bool(keep_tables)), hb)
        if rc != 0:
            raise BsqError('bsq_batch_upload failed (%d): %s' % (rc, self._err()))
        cap = int(o_len.sum(dtype=np.int64) + m_len.sum(dtype=np.int64))
        return ResidentBatch(self, hb[0], N, cap, tx_2bit)

    def microbench(self, which):
        """Measured issue rate (thread-ops/s, whole GPU) of one pipe; ``which`` is a
        BSQ_MB_* name such as ``'DADD'``."""
        out = ffi.new('double*')
        rc = self.lib.bsq_microbench(self.ctx, CONST['BSQ_MB_' + which], out)
        if rc != 0:
            raise BsqError('bsq_microbench failed (%d): %s' % (rc, self._err()))
        return out[0]

    def table_scores(self, origin, mutant, type, subst, go, ge):
        """H of every STD cell, (n+1) x (m+1) (``bsq_table_scores``)."""
        o, t = _c(origin, np.uint8), _c(mutant, np.uint8)
        prm, S = self._params(0, type, subst, go, ge)
        out = np.empty((len(o) + 1, len(t) + 1), dtype=np.float64)
        rc = self.lib.bsq_table_scores(self.ctx, prm, _ptr(o, 'uint8_t*'), len(o),
                                       _ptr(t, 'uint8_t*'), len(t), _ptr(out, 'double*'))
        if rc != 0:
            raise BsqError('bsq_table_scores failed (%d): %s' % (rc, self._err()))
        return out

    def timing(self):
        t = ffi.new('bsq_timing*')
        self.lib.bsq_get_timing(self.ctx, t)
        return {k: getattr(t, k) for k in ('fill_ms', 'traceback_ms', 'total_ms', 'cells',
                                           'fill_launches', 'h2d_bytes', 'd2h_bytes',
                                           'ptr_bytes', 'dtype', 'chunks', 'plan_ms', 'call_ms')}


class MultiEngine(object):
    """ONE batch over several GPUs of a box (``bsq_multi``): the batch is cut into contiguous
    ranges of equal work, one host thread + context per GPU aligns its range and writes straight
    into disjoint slices of the result arrays.  No collective, no peer traffic.
    ``devices``: CUDA ordinals (an ordinal may repeat), or an int n for ordinals 0 .. n-1."""

    def __init__(self, devices, lib=None):
        self.lib = lib if lib is not None else _ffi.load()
        devs = list(range(devices)) if isinstance(devices, int) else [int(d) for d in devices]
        arr = ffi.new('int32_t[]', devs)
        self.me = self.lib.bsq_multi_create(len(devs), arr)
        if self.me == ffi.NULL:
            raise BsqError('bsq_multi_create(%r) failed: %s' % (
                devs, ffi.string(self.lib.bsq_multi_last_error(ffi.NULL)).decode()))
        self.devices = devs

    def close(self):
        if self.me is not None and self.me != ffi.NULL:
            self.lib.bsq_multi_destroy(self.me)
        self.me = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def pinned_empty(self, n, dtype):
        dtype = np.dtype(dtype)
        nbytes = int(n) * dtype.itemsize
        p = self.lib.bsq_host_alloc(max(nbytes, 1))
        if p == ffi.NULL:
            raise BsqError('bsq_host_alloc(%d) failed' % nbytes)
        p = ffi.gc(p, self.lib.bsq_host_free, size=nbytes)
        return np.frombuffer(ffi.buffer(p, max(nbytes, 1)), dtype=dtype, count=int(n))

    def tx_capacity(self, o_len, m_len, tx_2bit=False):
        """Bytes ``tx_buf`` must hold (the worst case, every GPU packs its own region)."""
        ops = int(np.asarray(o_len, np.int64).sum() + np.asarray(m_len, np.int64).sum()) + 64 * len(self.devices)
        return ops // 4 + 16 if tx_2bit else ops

    def align(self, codes, o_off, o_len, m_off, m_len, mode, type, subst, go, ge, dmin=None, dmax=None,
              dtype=DTYPE_AUTO, transcripts=True, out=None, codes_2bit=False, ncodes=None, tx_2bit=False):
        """``Engine.align`` for one batch sharded over the engine's GPUs (``bsq_multi_align_batch``)."""
        codes = _c(codes, np.uint8)
        o_off, m_off = _c(o_off, np.int64), _c(m_off, np.int64)
        o_len, m_len = _c(o_len, np.int32), _c(m_len, np.int32)
        dmin, dmax = _c(dmin, np.int32), _c(dmax, np.int32)
        N = len(o_len)
        S = np.ascontiguousarray(np.asarray(subst, dtype=np.float64))
        prm = ffi.new('bsq_params*')
        prm.mode, prm.type, prm.alphabet_len, prm.max_new_mins = int(mode), int(type), S.shape[0], -1
        prm.go, prm.ge, prm.subst, prm.dtype = float(go), float(ge), ffi.cast('double*', S.ctypes.data), int(dtype)
        prm.flags = (F_CODES_2BIT if codes_2bit else 0) | (F_TX_2BIT if tx_2bit else 0)
        nletters = int(ncodes) if ncodes is not None else (4 * len(codes) if codes_2bit else len(codes))
        out = out or {}

        def arr(name, dt, n):
            a = out.get(name)
            if a is None:
                a = np.empty(n, dtype=dt)
            assert a.dtype == dt and len(a) >= n
            return a
        score = arr('score', np.float64, N)
        ints = {k: arr(k, np.int32, N) for k in ('end_x', 'end_y', 'start_x', 'start_y', 'status')}
        tx_off = tx_len = tx_buf = None
        cap = 0
        if transcripts:
            tx_off, tx_len = arr('tx_off', np.int64, N + 1), arr('tx_len', np.int32, N)
            cap = self.tx_capacity(o_len, m_len, tx_2bit)
            tx_buf = out.get('tx_buf')
            if tx_buf is None or len(tx_buf) < cap:
                tx_buf = np.empty(max(cap, 1), np.uint8)
            cap = len(tx_buf)
        rc = self.lib.bsq_multi_align_batch(
            self.me, prm, N, _ptr(codes, 'uint8_t*'), nletters, _ptr(o_off, 'int64_t*'), _ptr(o_len, 'int32_t*'),
            _ptr(m_off, 'int64_t*'), _ptr(m_len, 'int32_t*'), _ptr(dmin, 'int32_t*'), _ptr(dmax, 'int32_t*'),
            _ptr(score, 'double*'), _ptr(ints['end_x'], 'int32_t*'), _ptr(ints['end_y'], 'int32_t*'),
            _ptr(ints['start_x'], 'int32_t*'), _ptr(ints['start_y'], 'int32_t*'), _ptr(ints['status'], 'int32_t*'),
            _ptr(tx_off, 'int64_t*'), _ptr(tx_len, 'int32_t*'), _ptr(tx_buf, 'char*'), cap)
        if rc != 0:
            raise BsqError('bsq_multi_align_batch failed (%d): %s' % (
                rc, ffi.string(self.lib.bsq_multi_last_error(self.me)).decode()))
        return BatchResult(score[:N], ints['end_x'][:N], ints['end_y'][:N], ints['start_x'][:N], ints['start_y'][:N],
                           ints['status'][:N], None if tx_off is None else tx_off[:N + 1], tx_buf, tx_2bit=tx_2bit,
                           tx_lens=None if tx_len is None else tx_len[:N])

    def shards(self):
        """[(first pair, count, device, timing dict)] of the last :meth:`align`."""
        res = []
        for k in range(len(self.devices)):
            first, count, dev = ffi.new('int64_t*'), ffi.new('int64_t*'), ffi.new('int32_t*')
            t = ffi.new('bsq_timing*')
            self.lib.bsq_multi_get_shard(self.me, k, first, count, dev, t)
            res.append((first[0], count[0], dev[0], {f: getattr(t, f) for f in (
                'fill_ms', 'total_ms', 'cells', 'fill_launches', 'h2d_bytes', 'd2h_bytes', 'chunks', 'plan_ms', 'call_ms', 'dtype')}))
        return res


STATS_DTYPE = np.dtype([('n_match', 'i4'), ('n_subst', 'i4'), ('n_ins', 'i4'), ('n_del', 'i4'),
                        ('trunc_lo', 'i4'), ('trunc_hi', 'i4'), ('trunc_start_x', 'i4'),
                        ('trunc_start_y', 'i4'), ('rescore', 'f8')])


def _batch_stats(engine, handle, n):
    """``bsq_batch_stats``: op counts, Alignment.calculate_score and truncate_to_match for
    every pair of a traced resident batch, computed on the device (pw.py:367-448)."""
    out = np.zeros(n, dtype=STATS_DTYPE)
    rc = engine.lib.bsq_batch_stats(handle, ffi.cast('bsq_aln_stats*', out.ctypes.data))
    if rc != 0:
        raise BsqError('bsq_batch_stats failed (%d): %s' % (rc, engine._err()))
    return out


class SolvedBatch(object):
    """A solved batch whose pointer tables are still resident (``bsq_batch``)."""

    def __init__(self, engine, handle, score, end_x, end_y, status, cap, tx_2bit=False):
        self.engine, self.handle = engine, handle
        self.score, self.end_x, self.end_y, self.status = score, end_x, end_y, status
        self._cap, self._tx_2bit = (cap // 4 + 64 if tx_2bit else cap), tx_2bit

    def traceback(self):
        N = len(self.score)
        start_x, start_y = np.empty(N, np.int32), np.empty(N, np.int32)
        status = np.empty(N, np.int32)
        tx_off = np.empty(N + 1, np.int64)
        tx_buf = np.empty(max(self._cap, 1), np.uint8)
        total = ffi.new('int64_t*')
        rc = self.engine.lib.bsq_traceback_batch(
            self.handle, _ptr(start_x, 'int32_t*'), _ptr(start_y, 'int32_t*'),
            _ptr(status, 'int32_t*'), _ptr(tx_off, 'int64_t*'), _ptr(tx_buf, 'char*'),
            self._cap, total)
        if rc != 0:
            raise BsqError('bsq_traceback_batch failed (%d): %s' % (rc, self.engine._err()))
        return BatchResult(self.score, self.end_x, self.end_y, start_x, start_y, status,
                           tx_off, tx_buf, tx_2bit=self._tx_2bit)

    def stats(self):
        """Device-side Alignment post-processing (after :meth:`traceback`)."""
        return _batch_stats(self.engine, self.handle, len(self.score))

    def release(self):
        if self.handle is not None:
            self.engine.lib.bsq_release(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class ResidentBatch(object):
    """Inputs uploaded once; ``run()`` = fill + traceback + packing on the device."""

    def __init__(self, engine, handle, n, cap, tx_2bit=False):
        self.engine, self.handle, self.n = engine, handle, n
        self._cap, self._tx_2bit = (cap // 4 + 64 if tx_2bit else cap), tx_2bit

    def run(self, traceback=True):
        rc = self.engine.lib.bsq_batch_run(self.handle, int(bool(traceback)))
        if rc != 0:
            raise BsqError('bsq_batch_run failed (%d): %s' % (rc, self.engine._err()))
        return self.engine.timing()

    def fetch(self, transcripts=True):
        N = self.n
        score = np.empty(N, np.float64)
        ints = [np.empty(N, np.int32) for _ in range(5)]
        tx_off = np.empty(N + 1, np.int64) if transcripts else None
        tx_buf = np.empty(max(self._cap, 1), np.uint8) if transcripts else None
        total = ffi.new('int64_t*')
        rc = self.engine.lib.bsq_batch_fetch(
            self.handle, _ptr(score, 'double*'), *[_ptr(a, 'int32_t*') for a in ints],
            _ptr(tx_off, 'int64_t*'), _ptr(tx_buf, 'char*'), self._cap, total)
        if rc != 0:
            raise BsqError('bsq_batch_fetch failed (%d): %s' % (rc, self.engine._err()))
        return BatchResult(score, ints[0], ints[1], ints[2], ints[3], ints[4], tx_off, tx_buf,
                           tx_2bit=self._tx_2bit and transcripts)

    def stats(self):
        """Device-side Alignment post-processing (after ``run(traceback=True)``)."""
        return _batch_stats(self.engine, self.handle, self.n)

    def release(self):
        if self.handle is not None:
            self.engine.lib.bsq_release(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def pack_pairs(pairs):
    """[(origin_codes, mutant_codes), ...] -> (codes, o_off, o_len, m_off, m_len)."""
    o_len = np.array([len(p[0]) for p in pairs], dtype=np.int32)
    m_len = np.array([len(p[1]) for p in pairs], dtype=np.int32)
    tot = int(o_len.sum() + m_len.sum())
    codes = np.zeros(max(tot, 1), dtype=np.uint8)
    o_off = np.zeros(len(pairs), dtype=np.int64)
    m_off = np.zeros(len(pairs), dtype=np.int64)
    pos = 0
    for i, (o, t) in enumerate(pairs):
        o_off[i] = pos
        codes[pos:pos + len(o)] = o
        pos += len(o)
        m_off[i] = pos
        codes[pos:pos + len(t)] = t
        pos += len(t)
    return codes, o_off, o_len, m_off, m_len
