"""Parity test bodies shared by the CPU-emulator run (-m "not gpu") and the
real B200 run (-m gpu).  Every function takes an Engine; sizes are passed in."""
import os

import numpy as np

from biseqt_b200 import batch
from oracle import lean
from tests.gen import random_problem, score_family, mutate
from tests.util import (oracle_one, engine_batch, same_as, check_problems, golden_problem,
                        golden_expected)


def run_golden(eng, golden, limit_random=None, limit_large=None):
    nr = nl = 0
    for g in golden:
        if g['name'].startswith('random-'):
            nr += 1
            if limit_random is not None and nr > limit_random:
                continue
        if g['name'].startswith('large-'):
            nl += 1
            if limit_large is not None and nl > limit_large:
                continue
        r = engine_batch(eng, [golden_problem(g)])
        same_as(golden_expected(g), r, 0, what=g['name'])


def run_random_singles(eng, seed, count, max_len=13):
    rng = np.random.default_rng(seed)
    check_problems(eng, [random_problem(rng, max_len=max_len) for _ in range(count)])


def make_batch(rng, npairs, mode, typ, lo, hi, L=4, family=None, gomode=None, band=None):
    """npairs problems sharing scores/mode/type, heterogeneous lengths."""
    S, go, ge = score_family(rng, L, int(rng.integers(0, 4)) if family is None else family,
                             int(rng.integers(-1, 2)) if gomode is None else gomode)
    probs = []
    for _ in range(npairs):
        n = int(rng.integers(lo, hi + 1))
        o = rng.integers(0, L, size=n).astype(np.uint8)
        if rng.random() < 0.7:
            t, _ = mutate(rng, o, L, subst=0.1, go=0.05, ge=0.3)
        else:
            t = rng.integers(0, L, size=int(rng.integers(lo, hi + 1))).astype(np.uint8)
        m = len(t)
        dr = None
        if mode == 1:
            if band is not None:
                c = int(rng.integers(-3, 4))
                dr = (c - band, c + band)
            else:
                a, b = sorted(int(v) for v in rng.integers(-m - 2, n + 3, size=2))
                dr = (a, b)
            if typ == 0 and rng.random() < 0.8:
                dr = (min(dr[0], 0, n - m), max(dr[1], 0, n - m))
        probs.append(dict(origin=o, mutant=t, subst=S, go=go, ge=ge, mode=mode, type=typ,
                          diag_range=dr, L=L))
    return probs


def check_batch(eng, probs, **kw):
    r = engine_batch(eng, probs, **kw)
    for i, pr in enumerate(probs):
        same_as(oracle_one(pr), r, i, what='pair %d of batch (mode %d type %d)' % (i, pr['mode'], pr['type']))
    return r


def run_batches(eng, seed, npairs, lo, hi):
    rng = np.random.default_rng(seed)
    for mode, types in ((0, range(7)), (1, range(3))):
        for typ in types:
            check_batch(eng, make_batch(rng, npairs, mode, typ, lo, hi))


def run_multipass_std(eng, seed, n_hi=40, widths=(520, 700, 1100)):
    """Tables wider than the widest single-pass geometry (32x16 columns)."""
    rng = np.random.default_rng(seed)
    for w in widths:
        for typ in (0, 1, 4):
            S, go, ge = score_family(rng, 4, 1, -1)
            o = rng.integers(0, 4, size=int(rng.integers(n_hi // 2, n_hi))).astype(np.uint8)
            t = rng.integers(0, 4, size=w).astype(np.uint8)
            pos = int(rng.integers(0, w - len(o)))
            t[pos:pos + len(o)] = o   # plant the origin so local/overlap find something
            pr = dict(origin=o, mutant=t, subst=S, go=go, ge=ge, mode=0, type=typ, diag_range=None)
            check_batch(eng, [pr])


def run_wide_bands(eng, seed, length=300, radii=(40, 100, 200, 500)):
    rng = np.random.default_rng(seed)
    for r in radii:
        for typ in (0, 1, 2):
            probs = make_batch(rng, 2, 1, typ, length, length + 20, family=2, gomode=-1, band=r)
            check_batch(eng, probs)


def run_banded_interior(eng, seed, length=260, npairs=4):
    """Banded problems long enough to reach the ungated interior phase, with
    shifted bands (overlap offsets), odd/even widths, all types and go signs."""
    rng = np.random.default_rng(seed)
    for typ in (0, 1, 2):
        for gomode in (-1, 0, 1):
            for L in (4, 7):
                S, go, ge = score_family(rng, L, int(rng.integers(1, 4)), gomode)
                probs = []
                for _ in range(npairs):
                    n = int(rng.integers(length // 2, length))
                    off = 0 if typ == 0 else int(rng.integers(-length // 3, length // 3))
                    chunk = rng.integers(0, L, size=2 * length).astype(np.uint8)
                    o = chunk[max(0, -off):max(0, -off) + n]
                    src = chunk[max(0, off):max(0, off) + n]
                    t, _ = mutate(rng, src, L, subst=0.05, go=0.03, ge=0.3)
                    m = len(t)
                    r1, r2 = int(rng.integers(1, 45)), int(rng.integers(1, 45))
                    dr = (off - r1, off + r2)
                    if typ == 0:
                        dr = (min(dr[0], 0, n - m), max(dr[1], 0, n - m))
                    probs.append(dict(origin=o, mutant=t, subst=S, go=go, ge=ge, mode=1, type=typ,
                                      diag_range=dr, L=L))
                check_batch(eng, probs)


def run_very_wide_bands(eng, seed, length=1300, radius=560, npairs=1):
    """Bands wider than the 1024 diagonals of the widest K2 geometry go to K3 (strips)."""
    rng = np.random.default_rng(seed)
    for typ in (0, 1, 2):
        probs = make_batch(rng, npairs, 1, typ, length, length + 40, family=int(rng.integers(1, 3)),
                           gomode=int(rng.integers(-1, 2)), band=radius)
        check_batch(eng, probs)


def run_wide_band_batch(eng, seed, lens=(3000, 9000, 5200, 12000, 700, 8000), radius=600):
    """Several wide-band pairs of different lengths in one batch: their strips are claimed from ONE
    strip-major work list (strip j of every pair before strip j+1 of any; pairs run out of strips at
    different levels), every pair against the oracle."""
    rng = np.random.default_rng(seed)
    for typ, fam, gomode in ((0, 1, -1), (1, 2, 0), (2, 1, 1)):
        S, go, ge = score_family(rng, 4, fam, gomode)
        probs = []
        for n in lens:
            o = rng.integers(0, 4, size=n).astype(np.uint8)
            t, _ = mutate(rng, o, 4, subst=0.08, go=0.02, ge=0.3)
            c = int(rng.integers(-5, 6))
            dr = (c - radius, c + radius)
            if typ == 0:
                dr = (min(dr[0], 0, n - len(t)), max(dr[1], 0, n - len(t)))
            probs.append(dict(origin=o, mutant=t, subst=S, go=go, ge=ge, mode=1, type=typ, diag_range=dr, L=4))
        check_batch(eng, probs)


def run_forced_k3(make_engine, seed, count=150, cfgs=(0, 1, 2), wide_len=700):
    """The strip kernel on every banded problem, all three strip geometries."""
    os.environ['BSQ_FORCE_K3'] = '1'
    try:
        for cfg in cfgs:
            os.environ['BSQ_K3_CFG'] = str(cfg)
            eng = make_engine()
            try:
                rng = np.random.default_rng(seed + cfg)
                probs = []
                while len(probs) < count:
                    pr = random_problem(rng, max_len=40)
                    if pr['mode'] == 1:
                        probs.append(pr)
                check_problems(eng, probs)
                for typ in (0, 1, 2):
                    check_batch(eng, make_batch(rng, 9, 1, typ, 0, 150))
                run_banded_interior(eng, seed=seed + 7 + cfg, length=330, npairs=2)
                # bands wide enough (W = 501 >= 31 (R + K) + R + K) for the strip kernel's ungated interior steps
                run_very_wide_bands(eng, seed=seed + 11 + cfg, length=wide_len, radius=250, npairs=2)
                # a narrow band long enough for the traceback's vector-loaded 1024-letter windows, for tiles
                # requested ahead across dozens of strips, and for a transcript packed in two segments
                run_very_wide_bands(eng, seed=seed + 17 + cfg, length=4300, radius=24, npairs=2)
            finally:
                eng.close()
    finally:
        del os.environ['BSQ_FORCE_K3']
        os.environ.pop('BSQ_K3_CFG', None)


def run_uniform_batches(make_engine, seed):
    """Batches in which every pair has the same shape take the planner's fast path (pair
    records generated on the device); forced here for small batches, with and without chunking."""
    os.environ['BSQ_UNIFORM_MIN'] = '2'
    os.environ['BSQ_ARENA_CAP_BYTES'] = str(1 << 17)
    try:
        eng = make_engine()
    finally:
        del os.environ['BSQ_ARENA_CAP_BYTES']
    eng2 = make_engine()    # no arena cap: the two-phase form keeps every table resident
    try:
        rng = np.random.default_rng(seed)
        for mode, typ, n, m, dr in ((0, 1, 60, 60, None), (0, 0, 33, 47, None), (0, 4, 200, 170, None),
                                    (1, 2, 150, 140, (-12, 20)), (1, 0, 90, 95, (-9, 7)), (1, 1, 80, 80, (3, 30))):
            S, go, ge = score_family(rng, 4, int(rng.integers(1, 3)), int(rng.integers(-1, 2)))
            probs = []
            for i in range(37):
                o = rng.integers(0, 4, size=n).astype(np.uint8)
                t = rng.integers(0, 4, size=m).astype(np.uint8)
                k = min(n, m)
                if i % 5:
                    t[:k] = o[:k]
                    t[rng.integers(0, m, size=m // 10)] = rng.integers(0, 4, size=m // 10)
                probs.append(dict(origin=o, mutant=t, subst=S, go=go, ge=ge, mode=mode, type=typ, diag_range=dr))
            check_batch(eng, probs)
            # resident + two-phase forms of the same batch
            codes, o_off, o_len, m_off, m_len = batch.pack_pairs([(p['origin'], p['mutant']) for p in probs])
            kw = dict(dmin=[dr[0]] * len(probs), dmax=[dr[1]] * len(probs)) if dr else {}
            sb = eng2.solve(codes, o_off, o_len, m_off, m_len, mode, typ, S, go, ge, **kw)
            r = sb.traceback()
            for i, pr in enumerate(probs):
                same_as(oracle_one(pr), r, i, what='uniform two-phase %d' % i)
            sb.release()
    finally:
        eng.close()
        eng2.close()
        del os.environ['BSQ_UNIFORM_MIN']


def run_duo_short_reads(make_engine, seed, npairs=37):
    """STD LOCAL, small integer scores, one shape for the whole batch: the packed 16-bit
    kernel (two problems per lane, K1h).  Odd pair counts, edge shapes, go == 0 and go < 0,
    alphabets of 2..4 letters; the scalar I32 kernel (BSQ_NO_DUO) must give the same bytes."""
    os.environ['BSQ_UNIFORM_MIN'] = '2'
    try:
        eng = make_engine()
        try:
            rng = np.random.default_rng(seed)
            shapes = [(150, 150), (151, 151), (40, 151), (97, 33), (1, 1), (5, 0), (0, 7), (64, 152 - 1), (150, 120),
                      (250, 250), (100, 152), (255, 255), (200, 230), (150, 150)]
            for si, (n, m) in enumerate(shapes):
                for gomode in (-1, 0):
                    L = int(rng.choice([2, 3, 4, 4]))
                    ma, mi = int(rng.integers(1, 4)), -int(rng.integers(0, 4))
                    S = np.where(np.eye(L, dtype=bool), float(ma), float(mi))
                    ge = -float(rng.integers(0, 3))
                    go = 0.0 if gomode == 0 else -float(rng.integers(1, 6))
                    if si == len(shapes) - 1:   # the largest scores the 16-bit kernel accepts: 304 * 13 < 4000; wider tables use the 16-lane geometry
                        S = np.where(np.eye(L, dtype=bool), 13.0, -13.0)
                        ge, go = -7.0, (0.0 if gomode == 0 else -6.0)
                    probs = []
                    for i in range(npairs):
                        o = rng.integers(0, L, size=n).astype(np.uint8)
                        t = rng.integers(0, L, size=m).astype(np.uint8)
                        k = min(n, m)
                        if i % 4 and k:
                            sh = int(rng.integers(0, max(1, k // 3)))
                            t[sh:k] = o[:k - sh]
                            if m >= 10:
                                t[rng.integers(0, m, size=m // 10)] = rng.integers(0, L, size=m // 10)
                        probs.append(dict(origin=o, mutant=t, subst=S, go=go, ge=ge, mode=0, type=1, diag_range=None))
                    r = check_batch(eng, probs)
                    assert eng.timing()['dtype'] == batch.DTYPE_I16X2, (n, m, gomode)
                    os.environ['BSQ_NO_DUO'] = '1'
                    eng.reload_hooks()
                    try:
                        r1 = engine_batch(eng, probs)
                    finally:
                        del os.environ['BSQ_NO_DUO']
                        eng.reload_hooks()
                    assert eng.timing()['dtype'] == batch.DTYPE_I32
                    assert np.array_equal(r.score, r1.score) and np.array_equal(r.tx_off, r1.tx_off)
                    assert np.array_equal(r.tx_buf[:r.tx_off[-1]], r1.tx_buf[:r1.tx_off[-1]])
            # ragged batches: a handful of shapes, repeated an odd or even number of times, plus
            # pairs the packed kernel cannot take (two passes); neighbours in processing order share a
            # lane group.  Trials 3.. run the planner's multi-threaded passes (BSQ_PAR_GRAIN).
            for trial in range(6):
                if trial >= 3:
                    os.environ['BSQ_PAR_GRAIN'] = str(trial - 1)
                    eng.reload_hooks()
                L = 4
                S = np.where(np.eye(L, dtype=bool), 2.0, -3.0)
                go, ge = (-4.0, -1.0) if trial != 1 else (0.0, -2.0)
                shapes = [(int(rng.integers(20, 151)), int(rng.integers(20, 151))) for _ in range(5)]
                shapes += [(30, 200), (60, 400)]      # 16-lane geometry / needs two passes: scalar kernel
                probs = []
                for (n, m) in shapes:
                    for r in range(int(rng.integers(1, 6 if trial < 3 else 12))):
                        o = rng.integers(0, L, size=n).astype(np.uint8)
                        t = rng.integers(0, L, size=m).astype(np.uint8)
                        k = min(n, m)
                        t[:k] = o[:k]
                        t[rng.integers(0, m, size=max(1, m // 8))] = rng.integers(0, L, size=max(1, m // 8))
                        probs.append(dict(origin=o, mutant=t, subst=S, go=go, ge=ge, mode=0, type=1, diag_range=None))
                perm = rng.permutation(len(probs))
                probs = [probs[i] for i in perm]
                r = check_batch(eng, probs)
                assert eng.timing()['dtype'] == batch.DTYPE_I16X2
                assert eng.timing()['chunks'] >= 2     # packed chunks and scalar ones
                os.environ['BSQ_NO_DUO'] = '1'
                eng.reload_hooks()
                try:
                    r1 = engine_batch(eng, probs)
                finally:
                    del os.environ['BSQ_NO_DUO']
                    eng.reload_hooks()
                assert np.array_equal(r.score, r1.score) and np.array_equal(r.tx_off, r1.tx_off)
                assert np.array_equal(r.tx_buf[:r.tx_off[-1]], r1.tx_buf[:r1.tx_off[-1]])
            os.environ.pop('BSQ_PAR_GRAIN', None)
            eng.reload_hooks()
            # ragged batches that stay in caller order (one geometry, tables within 2x): every pair
            # its own shape, neighbours (2q, 2q+1) share a lane group, the shorter one padded
            for trial in range(8):
                blocked = trial >= 4     # the block-wise in-order planner (BSQ_PLAN_BLOCK: small blocks)
                trial %= 4
                if blocked:
                    os.environ['BSQ_PLAN_BLOCK'] = str(3 + trial)
                    os.environ['BSQ_PAR_GRAIN'] = '7'
                    os.environ['BSQ_ARENA_CAP_BYTES'] = '600000'   # several chunks (read when the engine is created)
                    eng.close()
                    eng = make_engine()
                L = 4
                S = np.where(np.eye(L, dtype=bool), 1.0, -3.0) if trial < 3 else np.where(np.eye(L, dtype=bool), 2.0, 0.0)
                go, ge = [(-5.0, -2.0), (0.0, -1.0), (-1.0, 0.0), (0.0, 0.0)][trial]
                lo, hi = [(90, 152), (160, 256), (90, 152), (100, 140)][trial]
                probs = []
                for i in range((2 * npairs if blocked else npairs) + trial):
                    n, m = int(rng.integers(lo, hi)), int(rng.integers(lo, hi))
                    o = rng.integers(0, L, size=n).astype(np.uint8)
                    t = rng.integers(0, L, size=m).astype(np.uint8)
                    k = min(n, m)
                    if i % 3:
                        t[:k] = o[:k]
                        t[rng.integers(0, m, size=max(1, m // 8))] = rng.integers(0, L, size=max(1, m // 8))
                    probs.append(dict(origin=o, mutant=t, subst=S, go=go, ge=ge, mode=0, type=1, diag_range=None))
                r = check_batch(eng, probs)
                assert eng.timing()['dtype'] == batch.DTYPE_I16X2
                assert (eng.timing()['chunks'] > 1) if blocked else (eng.timing()['chunks'] == 1)
                os.environ['BSQ_NO_DUO'] = '1'     # the scalar kernel on the same (in-order) plan
                eng.reload_hooks()
                try:
                    r1 = engine_batch(eng, probs)
                finally:
                    del os.environ['BSQ_NO_DUO']
                    eng.reload_hooks()
                assert eng.timing()['dtype'] == batch.DTYPE_I32
                assert np.array_equal(r.score, r1.score) and np.array_equal(r.tx_off, r1.tx_off)
                assert np.array_equal(r.tx_buf[:r.tx_off[-1]], r1.tx_buf[:r1.tx_off[-1]])
                for k in ('BSQ_PLAN_BLOCK', 'BSQ_PAR_GRAIN', 'BSQ_ARENA_CAP_BYTES'):
                    os.environ.pop(k, None)
                if not blocked:
                    eng.reload_hooks()
        finally:
            for k in ('BSQ_PLAN_BLOCK', 'BSQ_PAR_GRAIN', 'BSQ_ARENA_CAP_BYTES'):
                os.environ.pop(k, None)
            eng.close()
    finally:
        del os.environ['BSQ_UNIFORM_MIN']


def _py_stats(pr, tx, sx, sy):
    """Alignment.projected_len / calculate_score / truncate_to_match (pw.py:367-448) in plain
    Python, the way biseqt_b200.pw.Alignment (and the reference) compute them."""
    import re
    S, go, ge = pr['subst'], pr['go'], pr['ge']
    o, m = pr['origin'], pr['mutant']
    score, i, j = 0., sx, sy
    for match in re.finditer(r'(.)\1*', tx):
        run = match.group(0)
        op, num = run[0], len(run)
        if op in 'MS':
            score += sum(S[o[i + k]][m[j + k]] for k in range(num))
            i, j = i + num, j + num
        else:
            score += go + ge * num
            if op == 'I':
                j += num
            else:
                i += num
    trunc = None
    if 'M' in tx:
        a, b = tx.index('M'), tx.rindex('M')
        if a < b:
            ox = sx + sum(c in 'DS' for c in tx[:a])
            oy = sy + sum(c in 'IS' for c in tx[:a])
            trunc = (a, b, ox, oy)
    return {c: tx.count(c) for c in 'MSID'}, score, trunc


def run_stats(eng, seed, npairs=14, hi=70):
    """Device-side Alignment post-processing against the Python definitions."""
    rng = np.random.default_rng(seed)
    for mode, typ in ((0, 1), (0, 0), (0, 4), (1, 2), (1, 1)):
        probs = make_batch(rng, npairs, mode, typ, 5, hi)
        p0 = probs[0]
        codes, o_off, o_len, m_off, m_len = batch.pack_pairs([(p['origin'], p['mutant']) for p in probs])
        kw = {}
        if mode == 1:
            kw = dict(dmin=[p['diag_range'][0] for p in probs], dmax=[p['diag_range'][1] for p in probs])
        rb = eng.upload(codes, o_off, o_len, m_off, m_len, mode, typ, p0['subst'], p0['go'], p0['ge'], **kw)
        rb.run()
        r = rb.fetch()
        st = rb.stats()
        for i, pr in enumerate(probs):
            if int(r.status[i]) not in (0, 4) or not r.transcript(i):
                assert st['trunc_lo'][i] == -1 and st['n_match'][i] == 0
                continue
            cnt, score, trunc = _py_stats(pr, r.transcript(i), int(r.start_x[i]), int(r.start_y[i]))
            assert (st['n_match'][i], st['n_subst'][i], st['n_ins'][i], st['n_del'][i]) == \
                (cnt['M'], cnt['S'], cnt['I'], cnt['D']), i
            assert st['rescore'][i] == score, (i, st['rescore'][i], score)
            if trunc is None:
                assert st['trunc_lo'][i] == -1
            else:
                assert (st['trunc_lo'][i], st['trunc_hi'][i], st['trunc_start_x'][i], st['trunc_start_y'][i]) == trunc
        rb.release()


def run_big_ragged_plan(make_engine, seed, npairs=70000, all_modes=True):
    """The planner's parallel passes (counting sort, chunk scan) need > 65536 pairs to run on several
    threads: tiny ragged problems, several chunks, against the batch oracle and against the
    serial scan."""
    os.environ['BSQ_ARENA_CAP_BYTES'] = str(3 << 20)
    try:
        eng = make_engine()
    finally:
        del os.environ['BSQ_ARENA_CAP_BYTES']
    try:
        rng = np.random.default_rng(seed)
        o_len = rng.integers(0, 7, size=npairs).astype(np.int32)
        m_len = rng.integers(0, 7, size=npairs).astype(np.int32)
        tot = (o_len.astype(np.int64) + m_len).cumsum()
        o_off = np.concatenate([[0], tot[:-1]]).astype(np.int64)
        m_off = o_off + o_len
        codes = rng.integers(0, 4, size=int(tot[-1]) + 16, dtype=np.uint8)
        S = np.where(np.eye(4, dtype=bool), 2.0, -1.0)
        cases = ((0, 0, {}), (0, 1, {}), (1, 2, dict(dmin=np.full(npairs, -2, np.int32), dmax=np.full(npairs, 3, np.int32))))
        for mode, typ, kw in (cases if all_modes else cases[1:2]):
            r = eng.align(codes, o_off, o_len, m_off, m_len, mode, typ, S, -2.0, -1.0, **kw)
            assert eng.timing()['chunks'] > 2
            e = lean.align_batch(mode, typ, S, -2.0, -1.0, codes, o_off, o_len, m_off, m_len, kw.get('dmin'), kw.get('dmax'))
            assert np.array_equal(r.status, e['status'])
            ok = (e['status'] == 0) | (e['status'] == 4)
            assert np.array_equal(r.score[ok], e['score'][ok])
            assert np.array_equal(r.start_x[ok], e['start_x'][ok]) and np.array_equal(r.start_y[ok], e['start_y'][ok])
            assert np.array_equal(r.tx_len, e['tx_len'])
            os.environ['BSQ_SERIAL_SCAN'] = '1'
            eng.reload_hooks()
            try:
                r1 = eng.align(codes, o_off, o_len, m_off, m_len, mode, typ, S, -2.0, -1.0, **kw)
            finally:
                del os.environ['BSQ_SERIAL_SCAN']
                eng.reload_hooks()
            assert np.array_equal(r.tx_off, r1.tx_off) and np.array_equal(r.tx_buf[:r.tx_off[-1]], r1.tx_buf[:r1.tx_off[-1]])
            assert np.array_equal(r.status, r1.status) and np.array_equal(r.score, r1.score)
    finally:
        eng.close()


def run_ragged_reads(make_engine, seed, npairs=30000, lo=100, hi=151):
    """Ragged short reads through the caller-order planner (blocks of 4096 pairs, ordered by
    origin length inside a block, K1h duos of unequal shapes) against the batch oracle, and the
    same bytes from the sorting planner (BSQ_ALWAYS_SORT) and the scalar kernel (BSQ_NO_DUO)."""
    os.environ['BSQ_ARENA_CAP_BYTES'] = str(200 << 20)    # several chunks
    try:
        eng = make_engine()
    finally:
        del os.environ['BSQ_ARENA_CAP_BYTES']
    try:
        rng = np.random.default_rng(seed)
        o_len = rng.integers(lo, hi, size=npairs).astype(np.int32)
        m_len = rng.integers(lo, hi, size=npairs).astype(np.int32)
        tot = (o_len.astype(np.int64) + m_len).cumsum()
        o_off = np.concatenate([[0], tot[:-1]]).astype(np.int64)
        m_off = o_off + o_len
        codes = rng.integers(0, 4, size=int(tot[-1]) + 16, dtype=np.uint8)
        for i in range(0, npairs, 3):   # two thirds of the pairs are related reads
            k = min(o_len[i], m_len[i])
            codes[m_off[i]:m_off[i] + k] = codes[o_off[i]:o_off[i] + k]
            hits = rng.integers(0, k, size=k // 12)
            codes[m_off[i] + hits] = rng.integers(0, 4, size=len(hits))
        S = np.where(np.eye(4, dtype=bool), 1.0, -3.0)
        r = eng.align(codes, o_off, o_len, m_off, m_len, 0, 1, S, -5.0, -2.0)
        t = eng.timing()
        assert t['dtype'] == batch.DTYPE_I16X2 and t['chunks'] > 1
        e = lean.align_batch(0, 1, S, -5.0, -2.0, codes, o_off, o_len, m_off, m_len, None, None)
        assert np.array_equal(r.status, e['status'])
        assert np.array_equal(r.score, e['score'])
        assert np.array_equal(r.end_x, e['end_x']) and np.array_equal(r.end_y, e['end_y'])
        assert np.array_equal(r.start_x, e['start_x']) and np.array_equal(r.start_y, e['start_y'])
        assert r.transcripts() == lean.transcripts(e)
        for var in ('BSQ_ALWAYS_SORT', 'BSQ_NO_DUO'):
            os.environ[var] = '1'
            eng.reload_hooks()
            try:
                r1 = eng.align(codes, o_off, o_len, m_off, m_len, 0, 1, S, -5.0, -2.0)
            finally:
                del os.environ[var]
                eng.reload_hooks()
            assert np.array_equal(r.tx_off, r1.tx_off) and np.array_equal(r.tx_buf[:r.tx_off[-1]], r1.tx_buf[:r1.tx_off[-1]])
            assert np.array_equal(r.status, r1.status) and np.array_equal(r.score, r1.score)
    finally:
        eng.close()


def run_dtype_equivalence(eng, seed, npairs=12, lo=5, hi=60):
    """Integer scores: the I32 fast path and the F64 path give identical results."""
    rng = np.random.default_rng(seed)
    for mode, typ in ((0, 0), (0, 1), (0, 4), (1, 0), (1, 1), (1, 2)):
        for gomode in (-1, 0, 1):
            probs = make_batch(rng, npairs, mode, typ, lo, hi, family=1, gomode=gomode)
            a = check_batch(eng, probs, dtype=batch.DTYPE_F64)
            assert eng.timing()['dtype'] == batch.DTYPE_F64
            b = check_batch(eng, probs, dtype=batch.DTYPE_AUTO)
            assert eng.timing()['dtype'] == batch.DTYPE_I32
            assert np.array_equal(a.status, b.status) and np.array_equal(a.tx_buf[:a.tx_off[-1]], b.tx_buf[:b.tx_off[-1]])


def run_two_phase(eng, seed):
    rng = np.random.default_rng(seed)
    for mode, typ in ((0, 1), (1, 2)):
        probs = make_batch(rng, 9, mode, typ, 10, 50)
        p0 = probs[0]
        codes, o_off, o_len, m_off, m_len = batch.pack_pairs([(p['origin'], p['mutant']) for p in probs])
        kw = {}
        if mode == 1:
            kw = dict(dmin=[p['diag_range'][0] for p in probs], dmax=[p['diag_range'][1] for p in probs])
        sb = eng.solve(codes, o_off, o_len, m_off, m_len, mode, typ, p0['subst'], p0['go'], p0['ge'], **kw)
        for i, pr in enumerate(probs):
            e = oracle_one(pr)
            if e['status'] in (0, 2, 4):
                assert sb.score[i] == e['score'] and (sb.end_x[i], sb.end_y[i]) == tuple(e['end'])
        r = sb.traceback()
        for i, pr in enumerate(probs):
            same_as(oracle_one(pr), r, i, what='two-phase %d' % i)
        sb.release()


def run_resident(eng, seed):
    """upload once, run twice, fetch: same results as the one-shot call."""
    rng = np.random.default_rng(seed)
    for mode, typ in ((0, 1), (1, 2), (1, 0)):
        probs = make_batch(rng, 11, mode, typ, 10, 60)
        p0 = probs[0]
        codes, o_off, o_len, m_off, m_len = batch.pack_pairs([(p['origin'], p['mutant']) for p in probs])
        kw = {}
        if mode == 1:
            kw = dict(dmin=[p['diag_range'][0] for p in probs], dmax=[p['diag_range'][1] for p in probs])
        rb = eng.upload(codes, o_off, o_len, m_off, m_len, mode, typ, p0['subst'], p0['go'], p0['ge'], **kw)
        rb.run()
        t = rb.run()
        assert t['cells'] == sum(oracle_one(p)['ncells'] for p in probs if oracle_one(p)['status'] != 3)
        r = rb.fetch()
        for i, pr in enumerate(probs):
            same_as(oracle_one(pr), r, i, what='resident %d' % i)
        rb.release()


def run_block_plan_apis(make_engine, seed, npairs=14):
    """The caller-order block planner behind the two-phase (solve / traceback) and the resident
    (upload / run / fetch) entry points: ragged short reads, every table resident at once."""
    os.environ['BSQ_PLAN_BLOCK'] = '3'
    os.environ['BSQ_UNIFORM_MIN'] = '2'
    try:
        eng = make_engine()
        try:
            rng = np.random.default_rng(seed)
            S = np.where(np.eye(4, dtype=bool), 2.0, -3.0)
            probs = []
            for i in range(npairs):
                n, m = int(rng.integers(100, 141)), int(rng.integers(100, 141))
                o = rng.integers(0, 4, size=n).astype(np.uint8)
                t = rng.integers(0, 4, size=m).astype(np.uint8)
                k = min(n, m)
                t[:k] = o[:k]
                t[rng.integers(0, m, size=m // 9)] = rng.integers(0, 4, size=m // 9)
                probs.append(dict(origin=o, mutant=t, subst=S, go=-4.0, ge=-1.0, mode=0, type=1, diag_range=None))
            codes, o_off, o_len, m_off, m_len = batch.pack_pairs([(p['origin'], p['mutant']) for p in probs])
            sb = eng.solve(codes, o_off, o_len, m_off, m_len, 0, 1, S, -4.0, -1.0)
            assert eng.timing()['dtype'] == batch.DTYPE_I16X2
            r = sb.traceback()
            for i, pr in enumerate(probs):
                same_as(oracle_one(pr), r, i, what='two-phase, block plan %d' % i)
            sb.release()
            rb = eng.upload(codes, o_off, o_len, m_off, m_len, 0, 1, S, -4.0, -1.0)
            rb.run()
            rb.run()
            r = rb.fetch()
            for i, pr in enumerate(probs):
                same_as(oracle_one(pr), r, i, what='resident, block plan %d' % i)
            rb.release()
            # degenerate shapes (empty origin and / or mutant) next to ordinary ones in one lane group
            probs = []
            for i in range(3 * npairs):
                n, m = int(rng.integers(0, 7)), int(rng.integers(0, 7))
                o = rng.integers(0, 4, size=n).astype(np.uint8)
                t = rng.integers(0, 4, size=m).astype(np.uint8)
                if i % 2:
                    k = min(n, m)
                    t[:k] = o[:k]
                probs.append(dict(origin=o, mutant=t, subst=S, go=-4.0 if i < npairs else 0.0, ge=-1.0, mode=0, type=1, diag_range=None))
            for part in (probs[:npairs], probs[npairs:]):
                check_batch(eng, part)
                assert eng.timing()['dtype'] == batch.DTYPE_I16X2
        finally:
            eng.close()
    finally:
        del os.environ['BSQ_PLAN_BLOCK']
        del os.environ['BSQ_UNIFORM_MIN']


def run_edge_cases(eng):
    E = np.zeros(0, np.uint8)
    A4 = np.array([0, 1, 2, 3, 0, 1], np.uint8)
    for mode, typ, dr in ((0, 0, None), (0, 1, None), (0, 4, None), (1, 0, (0, 0)), (1, 2, (-2, 2)), (1, 1, (-1, 3))):
        for o, t in ((E, E), (A4, E), (E, A4), (A4[:1], A4[:1])):
            pr = dict(origin=o, mutant=t, subst=np.eye(4), go=-1.0, ge=-1.0, mode=mode, type=typ, diag_range=dr)
            check_batch(eng, [pr])
    # empty batch
    r = eng.align(np.zeros(1, np.uint8), np.zeros(0, np.int64), np.zeros(0, np.int32),
                  np.zeros(0, np.int64), np.zeros(0, np.int32), 0, 0, np.eye(4), 0, 0)
    assert len(r) == 0
    # alphabets of 1 and 20 letters
    rng = np.random.default_rng(3)
    for L in (1, 20):
        for mode, typ in ((0, 1), (1, 2)):
            check_batch(eng, make_batch(rng, 5, mode, typ, 5, 40, L=L))
    run_bad_codes(eng)


def run_bad_codes(eng):
    """A letter code >= alphabet_len is refused (BSQ_ERR_ARG) on every entry point instead of
    indexing the shared-memory score table out of bounds; unreferenced bytes of codes[] may hold
    anything (e.g. separators)."""
    import pytest
    from biseqt_b200._ffi import BsqError
    o = np.array([0, 1, 2, 3, 250, 1, 2], np.uint8)
    t = np.array([0, 1, 2, 3, 1, 2], np.uint8)
    for mode, typ, dr in ((0, 1, None), (1, 2, (-3, 3))):
        pr = dict(origin=o, mutant=t, subst=np.eye(4), go=-1.0, ge=-1.0, mode=mode, type=typ, diag_range=dr)
        with pytest.raises(BsqError, match='alphabet_len'):
            engine_batch(eng, [pr])
    codes = np.concatenate([o, t])
    args = ([0], [7], [7], [6], 0, 1, np.eye(4), -1.0, -1.0)
    with pytest.raises(BsqError, match='alphabet_len'):
        eng.solve(codes, *args)
    with pytest.raises(BsqError, match='alphabet_len'):
        eng.upload(codes, *args)
    # the same bytes with the bad letter outside every frame: fine
    sep = np.concatenate([o[:4], [255], t]).astype(np.uint8)
    r = eng.align(sep, [0], [4], [5], [6], 0, 1, np.eye(4), -1.0, -1.0)
    assert int(r.status[0]) == 0 and r.transcript(0) == 'MMMM'
    # and the engine still works afterwards
    check_batch(eng, [dict(origin=t, mutant=t, subst=np.eye(4), go=-1.0, ge=-1.0, mode=0, type=0, diag_range=None)])


def _same_results(r, q, what):
    assert np.array_equal(r.status, q.status), what
    assert np.array_equal(r.score, q.score), what
    for k in ('end_x', 'end_y', 'start_x', 'start_y'):
        ok = (r.status == 0)
        assert np.array_equal(getattr(r, k)[ok], getattr(q, k)[ok]), (what, k)
    assert np.array_equal(r.tx_off, q.tx_off), what
    assert r.transcripts() == q.transcripts(), what


def run_compact_io(make_engine, seed, npairs=41):
    """BSQ_F_CODES_2BIT (four letters per byte in), BSQ_F_TX_2BIT (four ops per byte out) and
    implicit back-to-back offsets against the plain one-byte forms of the same batches (which the
    other tests pin to the oracle): ragged and uniform STD, banded, chunked (per-chunk shipping of
    2-bit words), the two-phase and resident entry points."""
    rng = np.random.default_rng(seed)
    os.environ['BSQ_UNIFORM_MIN'] = '2'
    os.environ['BSQ_ARENA_CAP_BYTES'] = str(1 << 17)
    try:
        engines = [make_engine()]
    finally:
        del os.environ['BSQ_ARENA_CAP_BYTES']
    try:
        engines.append(make_engine())
    finally:
        del os.environ['BSQ_UNIFORM_MIN']
    engines.append(make_engine())
    try:
        for eng in engines:
            for mode, typ, uniform in ((0, 1, True), (0, 1, False), (0, 0, False), (1, 2, False), (1, 1, True)):
                S, go, ge = score_family(rng, 4, int(rng.integers(1, 3)), int(rng.integers(-1, 1)))
                pairs, dmin, dmax = [], [], []
                for i in range(npairs):
                    n = 90 if uniform else int(rng.integers(0, 160))
                    m = 90 if uniform else int(rng.integers(0, 160))
                    o = rng.integers(0, 4, size=n).astype(np.uint8)
                    t = rng.integers(0, 4, size=m).astype(np.uint8)
                    k = min(n, m)
                    if i % 4:
                        t[:k] = o[:k]
                        if m > 8:
                            t[rng.integers(0, m, size=m // 8)] = rng.integers(0, 4, size=m // 8)
                    pairs.append((o, t))
                    dmin.append(-7)
                    dmax.append(9)
                codes, o_off, o_len, m_off, m_len = batch.pack_pairs(pairs)
                kw = dict(dmin=dmin, dmax=dmax) if mode == 1 else {}
                base = eng.align(codes, o_off, o_len, m_off, m_len, mode, typ, S, go, ge, **kw)
                packed = batch.pack_codes_2bit(codes)
                what = 'mode %d type %d uniform %s' % (mode, typ, uniform)
                _same_results(base, eng.align(packed, o_off, o_len, m_off, m_len, mode, typ, S, go, ge, codes_2bit=True,
                                              ncodes=len(codes), **kw), what + ' 2-bit in')
                _same_results(base, eng.align(codes, o_off, o_len, m_off, m_len, mode, typ, S, go, ge, tx_2bit=True, **kw),
                              what + ' 2-bit out')
                _same_results(base, eng.align(packed, None, o_len, None, m_len, mode, typ, S, go, ge, codes_2bit=True,
                                              ncodes=len(codes), tx_2bit=True, **kw), what + ' 2-bit in/out, no offsets')
                _same_results(base, eng.align(codes, None, o_len, None, m_len, mode, typ, S, go, ge, **kw), what + ' no offsets')
        eng = engines[-1]
        sb = eng.solve(packed, None, o_len, None, m_len, mode, typ, S, go, ge, codes_2bit=True, ncodes=len(codes), tx_2bit=True, **kw)
        _same_results(base, sb.traceback(), 'two-phase 2-bit')
        sb.release()
        rb = eng.upload(packed, o_off, o_len, m_off, m_len, mode, typ, S, go, ge, codes_2bit=True, ncodes=len(codes), tx_2bit=True, **kw)
        rb.run()
        _same_results(base, rb.fetch(), 'resident 2-bit')
        rb.release()
        # a 3-letter alphabet in 2-bit form: code 3 is out of range and must be refused
        import pytest
        from biseqt_b200._ffi import BsqError
        with pytest.raises(BsqError, match='alphabet_len'):
            eng.align(batch.pack_codes_2bit(np.array([0, 1, 3, 2, 1, 0, 1, 2], np.uint8)), [0], [4], [4], [4], 0, 1,
                      np.eye(3), -1.0, -1.0, codes_2bit=True, ncodes=8)
        with pytest.raises(BsqError, match='4 letters'):
            eng.align(packed, o_off, o_len, m_off, m_len, 0, 1, np.eye(5), -1.0, -1.0, codes_2bit=True, ncodes=len(codes))
    finally:
        for eng in engines:
            eng.close()


def run_multi(make_engine, make_multi, seed, npairs=60):
    """bsq_multi: one batch cut over several contexts (on one GPU here: ordinals may repeat; the
    8-GPU box runs the same code with distinct ordinals) equals the single-context result byte for
    byte -- ragged and uniform batches, banded, explicit and back-to-back offsets, compact forms,
    fewer pairs than shards."""
    rng = np.random.default_rng(seed)
    eng = make_engine()
    try:
        for nshards in (1, 2, 3):
            me = make_multi(nshards)
            try:
                for mode, typ, uniform, n_pairs in ((0, 1, False, npairs), (0, 1, True, npairs), (1, 2, False, npairs),
                                                    (1, 0, True, 7), (0, 0, False, 2), (0, 4, False, 1)):
                    S, go, ge = score_family(rng, 4, int(rng.integers(1, 3)), int(rng.integers(-1, 1)))
                    pairs, dmin, dmax = [], [], []
                    for i in range(n_pairs):
                        n = 70 if uniform else int(rng.integers(0, 200))
                        m = 64 if uniform else int(rng.integers(0, 200))
                        o = rng.integers(0, 4, size=n).astype(np.uint8)
                        t = rng.integers(0, 4, size=m).astype(np.uint8)
                        k = min(n, m)
                        if i % 4:
                            t[:k] = o[:k]
                            if m > 8:
                                t[rng.integers(0, m, size=m // 8)] = rng.integers(0, 4, size=m // 8)
                        pairs.append((o, t))
                        dmin.append(min(-5, n - m))
                        dmax.append(max(6, n - m))
                    codes, o_off, o_len, m_off, m_len = batch.pack_pairs(pairs)
                    kw = dict(dmin=dmin, dmax=dmax) if mode == 1 else {}
                    base = eng.align(codes, o_off, o_len, m_off, m_len, mode, typ, S, go, ge, **kw)
                    packed = batch.pack_codes_2bit(codes)
                    what = '%d shards, mode %d type %d uniform %s' % (nshards, mode, typ, uniform)
                    for variant, r in (
                            ('plain', me.align(codes, o_off, o_len, m_off, m_len, mode, typ, S, go, ge, **kw)),
                            ('no offsets', me.align(codes, None, o_len, None, m_len, mode, typ, S, go, ge, **kw)),
                            ('compact', me.align(packed, None, o_len, None, m_len, mode, typ, S, go, ge, codes_2bit=True,
                                                 ncodes=len(codes), tx_2bit=True, **kw)),
                            ('2-bit in, offsets', me.align(packed, o_off, o_len, m_off, m_len, mode, typ, S, go, ge,
                                                           codes_2bit=True, ncodes=len(codes), **kw))):
                        assert np.array_equal(base.status, r.status), (what, variant)
                        assert np.array_equal(base.score, r.score), (what, variant)
                        ok = base.status == 0
                        for k in ('end_x', 'end_y', 'start_x', 'start_y'):
                            assert np.array_equal(getattr(base, k)[ok], getattr(r, k)[ok]), (what, variant, k)
                        assert np.array_equal(base.tx_len, r.tx_len), (what, variant)
                        assert base.transcripts() == r.transcripts(), (what, variant)
                    sh = me.shards()
                    assert sum(c for _, c, _, _ in sh) == n_pairs and sh[0][0] == 0
                    assert all(sh[k][0] + sh[k][1] == sh[k + 1][0] for k in range(nshards - 1))
            finally:
                me.close()
    finally:
        eng.close()


def run_std_as_band(make_engine, seed, sizes=((300, 280), (256, 700), (1000, 1040))):
    """One to a few big STD GLOBAL pairs are planned as full-width B_GLOBAL bands and take the
    strip pipeline (K3L) -- BASELINE.json configs[0]'s shape; same answer as the oracle's STD
    GLOBAL and as K1 (BSQ_NO_STD_BAND), all score families and signs of go."""
    rng = np.random.default_rng(seed)
    eng = make_engine()
    os.environ['BSQ_NO_STD_BAND'] = '1'
    try:
        eng_k1 = make_engine()
    finally:
        del os.environ['BSQ_NO_STD_BAND']
    try:
        for (n, m) in sizes:
            for fam, gomode in ((1, -1), (2, 0), (1, 1), (2, -1)):
                S, go, ge = score_family(rng, 4, fam, gomode)
                probs = []
                for i in range(1 + (n < 500)):
                    o = rng.integers(0, 4, size=n).astype(np.uint8)
                    t, _ = mutate(rng, o, 4, subst=0.1, go=0.05, ge=0.3)
                    t = t[:m] if len(t) >= m else np.concatenate([t, rng.integers(0, 4, size=m - len(t)).astype(np.uint8)])
                    probs.append(dict(origin=o, mutant=t, subst=S, go=go, ge=ge, mode=0, type=0, diag_range=None))
                r = check_batch(eng, probs)
                r1 = engine_batch(eng_k1, probs)
                assert np.array_equal(r.score, r1.score) and r.transcripts() == r1.transcripts()
        # two-phase form (what biseqt_b200.pw.Aligner issues) and a degenerate partner that keeps the batch on K1
        pr = probs[0]
        codes, o_off, o_len, m_off, m_len = batch.pack_pairs([(pr['origin'], pr['mutant'])])
        sb = eng.solve(codes, o_off, o_len, m_off, m_len, 0, 0, pr['subst'], pr['go'], pr['ge'])
        same_as(oracle_one(pr), sb.traceback(), 0, what='two-phase STD GLOBAL as band')
        sb.release()
        small = dict(pr, origin=pr['origin'][:40], mutant=pr['mutant'][:50])
        check_batch(eng, [pr, small])
    finally:
        eng.close()
        eng_k1.close()


def run_std_as_band_many(make_engine, seed, npairs=300, n=1000):
    """Hundreds of big STD GLOBAL pairs in one batch, planned as full-width bands (16 strips each,
    one strip-major work list) against the same batch on K1 (which the other tests pin to the oracle)."""
    rng = np.random.default_rng(seed)
    S, go, ge = score_family(rng, 4, 2, -1)
    probs = []
    for i in range(npairs):
        o = rng.integers(0, 4, size=int(rng.integers(n // 2, n + 1))).astype(np.uint8)
        t, _ = mutate(rng, o, 4, subst=0.1, go=0.05, ge=0.3)
        if len(t) < 256:
            t = np.concatenate([t, rng.integers(0, 4, size=256 - len(t)).astype(np.uint8)])
        probs.append(dict(origin=o, mutant=t, subst=S, go=go, ge=ge, mode=0, type=0, diag_range=None))
    eng = make_engine()
    os.environ['BSQ_NO_STD_BAND'] = '1'
    try:
        eng_k1 = make_engine()
    finally:
        del os.environ['BSQ_NO_STD_BAND']
    try:
        r = engine_batch(eng, probs)
        r1 = engine_batch(eng_k1, probs)
        assert np.array_equal(r.status, r1.status) and np.array_equal(r.score, r1.score)
        assert np.array_equal(r.end_x, r1.end_x) and np.array_equal(r.end_y, r1.end_y)
        assert np.array_equal(r.start_x, r1.start_x) and np.array_equal(r.start_y, r1.start_y)
        assert r.transcripts() == r1.transcripts()
        for i in (0, npairs // 2, npairs - 1):
            same_as(oracle_one(probs[i]), r, i, what='pair %d of a big STD GLOBAL batch' % i)
    finally:
        eng.close()
        eng_k1.close()


def run_lifetime(make_engine):
    """Batches that outlive their engine (an Engine closed, or collected, before its SolvedBatch /
    ResidentBatch): release and every other call on them must not touch the dead context."""
    import pytest
    from biseqt_b200._ffi import BsqError
    rng = np.random.default_rng(8)
    o = rng.integers(0, 4, 40).astype(np.uint8)
    codes = np.concatenate([o, o])
    args = ([0], [40], [40], [40], 0, 0, np.eye(4), -1.0, -1.0)
    eng = make_engine()
    sb = eng.solve(codes, *args)
    rb = eng.upload(codes, *args)
    rb.run()
    eng.close()
    with pytest.raises(BsqError):
        sb.traceback()
    with pytest.raises(BsqError):
        rb.run()
    with pytest.raises(BsqError):
        rb.fetch()
    sb.release()
    rb.release()
    eng2 = make_engine()      # a fresh context is unaffected
    try:
        sb2 = eng2.solve(codes, *args)
        assert sb2.traceback().transcript(0) == 'M' * 40
        sb2.release()
    finally:
        eng2.close()


def run_table_scores(eng, seed):
    rng = np.random.default_rng(seed)
    for typ in range(7):
        pr = random_problem(rng, max_len=30, L=4, min_len=3)
        e = lean.align(pr['origin'], pr['mutant'], pr['subst'], pr['go'], pr['ge'], 0, typ, None, want_table=True)
        H = eng.table_scores(pr['origin'], pr['mutant'], typ, pr['subst'], pr['go'], pr['ge'])
        assert np.array_equal(H, e['table'])


def run_chunked(make_engine, seed):
    """Arena cap so small that the batch is cut into several chunks."""
    os.environ['BSQ_ARENA_CAP_BYTES'] = str(1 << 18)
    try:
        eng = make_engine()
    finally:
        del os.environ['BSQ_ARENA_CAP_BYTES']
    try:
        rng = np.random.default_rng(seed)
        probs = make_batch(rng, 40, 0, 1, 100, 140, family=1, gomode=-1)
        check_batch(eng, probs)
        assert eng.timing()['chunks'] > 1
        probs = make_batch(rng, 30, 1, 2, 300, 400, family=2, gomode=0, band=30)
        check_batch(eng, probs)
        assert eng.timing()['chunks'] > 1
        # equal-sized pairs keep the caller's order: results leave chunk by chunk
        S, go, ge = score_family(rng, 4, 1, -1)
        probs = []
        for i in range(70):
            o = rng.integers(0, 4, size=120).astype(np.uint8)
            t = o.copy()
            t[rng.integers(0, 120, size=12)] = rng.integers(0, 4, size=12)
            if i % 9 == 0:
                t = rng.integers(0, 4, size=120).astype(np.uint8)
            probs.append(dict(origin=o, mutant=t, subst=S, go=go, ge=ge, mode=0, type=1, diag_range=None))
        check_batch(eng, probs)
        assert eng.timing()['chunks'] > 2
    finally:
        eng.close()
