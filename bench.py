#!/usr/bin/env python
"""bench.py -- GCUPS of the pwlib DP path (fill + traceback) on B200.

A "step" is one pass of the hot path over one batch of synthetic pairs of
BASELINE.json's configs[1] (1M read pairs of 150 bp, STD LOCAL, affine gaps;
integer IgBlast-like scores).  Three numbers:

  value   GCUPS with inputs already resident in HBM (bsq_batch_run), timed with
          CUDA events inside the library on its own stream
  e2e     the same metric through the public batch call (bsq_align_batch) with
          pinned HOST buffers, H2D + D2H inside the timed region
  cpu_baseline / --impl reference
          the reference's own C library (oracle/_ref/pwlib.so, compiled from the
          untouched sources) on the host cores, bounded sample of the same pairs

One process per GPU (torchrun); pairs are sharded by rank, no collective on
the data path (weak scaling: every rank aligns its own 1M pairs).  torch is
imported only for the barrier / max-over-ranks of a multi-GPU run.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bench_workloads as bw  # noqa: E402

WORKLOAD_PAIRS = {'c2_int': 1000000, 'c2_logodds': 1000000, 'c3_int': 100000, 'c3_logodds': 100000,
                  'c3_blocal_int': 100000, 'c3_blocal_logodds': 100000,
                  'c4_int': 10000, 'c1': 1, 'c2w_int': 400000, 'c5_int': 1, 'c5_logodds': 1, 'c5s_int': 64, 'c5s_logodds': 64}


# which SM pipes each microbenchmarked rate (bsq_microbench, csrc/bsq_microbench.cu) spans; the
# rates themselves are measured in every run and stored in profiles/ALU_PEAKS.json
PIPE_NOTES = {
    'DADD': 'add.rn.f64 -> DADD, FP64 pipe only (64 lanes/clk/SM)',
    'DMAX': 'setp.gt.f64 + selp.f64 -> DSETP (FP64 pipe) + 2 SEL (ALU pipe); counted as one op',
    'DSETP': 'setp.eq.f64 + predicated LOP3 -> FP64 pipe + ALU pipe; counted as one op',
    'IADD': 'add.s32 -> ptxas alternates IADD3 (ALU pipe) and IMAD (FMA pipe): a TWO-pipe rate; '
            'the ALU pipe alone is the VIMAX3 figure',
    'VIMAX3': 'VIMNMX3.S32 (DPX), ALU pipe only',
    'VIADDMAX': 'VIADDMNMX.S32 (DPX), ALU pipe only',
    'VADD2': 'VIADD.16x2, ALU pipe only; counted as two halfword adds per instruction',
    'SHFL': 'shfl.sync.up.b32, one per warp instruction counted as 32 thread-ops',
    'LDS64': 'ld.shared.b64, conflict free',
}


def k1h_traffic():
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_k1h_traffic.json')) as f:
            return json.load(f)['traffic']
    except Exception:
        return None


def ops_per_cell(go):
    """Algorithmic ALU ops per cell: 5 adds + 3 max + 4 equality
    tests; the two '+go' adds vanish when go == 0."""
    return 12 if go != 0 else 10


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        # the sampler is started before the warm-up (nvidia-smi takes a while to come up on an
        # 8-GPU box); keep the samples that arrived during the timed region, or, if the region was
        # shorter than the sampling period, the ones around it
        rows = [r for t, r in self.rows
                if self.t_begin is None or (self.t_begin - 0.02 <= t <= (self.t_end or t) + 0.05)]
        if not rows:
            rows = [r for t, r in self.rows if self.t_begin is None or t >= self.t_begin - 1.0]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if r[5 + k].lower().startswith('active'):
                    reasons.add(nme)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(mx)) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------
# CPU reference: the unmodified pwlib through oracle/pwref.py, one process per core
# ---------------------------------------------------------------------------
def _cpu_worker(task):
    """Aligns pairs `idx` of workload w with the reference library (or the oracle port) and
    returns (cells, busy seconds, pairs, results) -- results[i] = (status, score, end, start,
    transcript) of every pair, for the per-pair parity verdict against the GPU arm."""
    kind, w, idx, variant, keep = task
    from oracle import pwref, lean
    use_ref = (kind == 'reference')
    subst = w['subst'].tolist()
    cells, busy, results = 0, 0.0, []
    for i in idx:
        o = w['codes'][w['o_off'][i]:w['o_off'][i] + w['o_len'][i]]
        t = w['codes'][w['m_off'][i]:w['m_off'][i] + w['m_len'][i]]
        dr = None if w['mode'] == 0 else (int(w['dmin'][i]), int(w['dmax'][i]))
        if use_ref:
            P = pwref.Problem(o, t, subst, w['go'], w['ge'], w['mode'], w['type'], dr, variant=variant)
            res = [pwref.INIT_FAIL, None, None, None, None]
            t0 = time.perf_counter()     # init + solve + traceback + free
            if P.init() == 0:
                sc = P.solve()
                tb = None
                if sc is not None:
                    # pw.c:132-134 exit(1)s when the path is all indels from (0,0); such a pair
                    # (never produced by these generators) is classified, not run
                    end = P.end_xy()
                    if end == (0, 0) or 0 in (len(o), len(t)):
                        res = [pwref.REF_WOULD_ABORT, sc, end, None, None]
                    else:
                        tb = P.traceback()
                        res = [pwref.OK, sc, end, (tb[1], tb[2]), tb[3]] if tb is not None else [pwref.EMPTY, sc, end, None, None]
                else:
                    res = [pwref.NO_ALIGNMENT, None, None, None, None]
                P.free()
            busy += time.perf_counter() - t0
        else:
            t0 = time.perf_counter()
            a = lean.align(o, t, w['subst'], w['go'], w['ge'], w['mode'], w['type'], dr)
            busy += time.perf_counter() - t0
            res = [a['status'], a['score'], a['end'], a['start'], a['transcript']]
        if keep:
            results.append((int(i), res))
        cells += bw_cells_one(w, i)
    return cells, busy, len(idx), results


def bw_cells_one(w, i):
    n, m = int(w['o_len'][i]), int(w['m_len'][i])
    if w['mode'] == 0:
        return (n + 1) * (m + 1)
    from oracle import lean
    return lean.count_cells(1, n, m, int(w['dmin'][i]), int(w['dmax'][i]))


def parity_verdict(w, results, r):
    """the reference's (status, score, end, start, transcript) for the very
    pairs the CPU leg just timed against the GPU arm's result r (BatchResult) on the same pairs."""
    mism, worst = 0, []
    for i, (st, sc, end, start, tx) in results:
        ok = int(r.status[i]) == st
        if ok and st in (0, 2, 4):
            ok = float(r.score[i]) == sc and (int(r.end_x[i]), int(r.end_y[i])) == tuple(end)
        if ok and st == 0:
            ok = r.transcript(i) == tx and (int(r.start_x[i]), int(r.start_y[i])) == tuple(start)
        if not ok:
            mism += 1
            if len(worst) < 5:
                worst.append(int(i))
    out = {'pairs': len(results), 'mismatches': mism,
           'compared': 'status, score (bit-exact double), end cell, start cell, transcript'}
    if worst:
        out['first_mismatching_pairs'] = worst
    return out


def cpu_reference_run(w, pairs_per_core, cores=None, pool=None, variant='O2', keep=False):
    """Aligns a bounded sample on all host cores; returns dict(value GCUPS, ...)."""
    from oracle import pwref
    import multiprocessing as mp
    kind = 'reference' if pwref.available(variant) else 'port'
    cores = cores or (os.cpu_count() or 1)
    N = len(w['o_len'])
    total = min(N, pairs_per_core * cores)
    idx = np.arange(total)
    slim = {k: w[k] for k in ('codes', 'o_off', 'o_len', 'm_off', 'm_len', 'mode', 'type', 'subst', 'go', 'ge')}
    if w['mode'] == 1:
        slim['dmin'], slim['dmax'] = w['dmin'], w['dmax']
    hi = int(max(w['o_off'][total - 1] + w['o_len'][total - 1], w['m_off'][total - 1] + w['m_len'][total - 1]))
    slim['codes'] = w['codes'][:hi]
    tasks = [(kind, slim, idx[c::cores], variant, keep) for c in range(cores)]
    own = pool is None
    if own:
        pool = mp.get_context('fork').Pool(cores)
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, tasks)
    wall = time.perf_counter() - t0
    if own:
        pool.close()
    cells = sum(r[0] for r in res)
    # every worker ran concurrently; aggregate rate = sum of per-worker rates over
    # the timed regions (struct construction in Python is excluded)
    rate = sum(r[0] / r[1] for r in res if r[1] > 0)
    return {'value': rate / 1e9, 'unit': 'GCUPS', 'cores': cores, 'kind': kind,
            'sample': '%d of the workload\'s pairs (%d per core), init+solve+traceback+free per pair, %s build'
                      % (total, pairs_per_core, '-O2' if variant == 'O2' else 'reference-flags (-g, no -O: Makefile:1)'),
            'pairs_per_s': sum(r[2] / r[1] for r in res if r[1] > 0), 'wall_s': wall, 'cells': cells,
            'results': [x for r in res for x in r[3]]}


def cpu_sample_per_core(kind, cores):
    """Pairs per host core of the bounded CPU sample: C2 1250 per core (= BASELINE.md section 3's
    2e4 pairs on a 16-core box, ~2.5 s per step), the long-pair configs one pair per core."""
    return 1250 if kind.startswith('c2') else 1


def bench_config(desc, pairs_per_gpu):
    """`config` of the JSON line: identical (keys and values) for both arms."""
    return {'workload': desc, 'pairs_per_gpu': int(pairs_per_gpu), 'seed': 2}


def run_reference_arm(args, rank, world, emit):
    if rank != 0:
        return
    kind = args.workload
    cores = os.cpu_count() or 1
    per_core = cpu_sample_per_core(kind, cores)
    npairs = max(per_core * cores, 64)
    w = bw.make(kind, min(npairs, WORKLOAD_PAIRS[kind]), seed=2)
    import multiprocessing as mp
    pool = mp.get_context('fork').Pool(cores)
    times, last = [], None
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        last = cpu_reference_run(w, per_core, cores, pool)
        if s >= args.warmup:
            times.append((time.perf_counter() - t0, last))
    pool.close()
    val = float(np.mean([r['value'] for _, r in times]))
    line = {
        'impl': 'reference', 'metric': 'GCUPS', 'value': val, 'unit': 'GCUPS', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * float(np.mean([t for t, _ in times])), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        # the GPU arm's workload (same keys and values as its `config`); each step of this arm times
        # a bounded sample of it (`sample_pairs_per_step`)
        'config': bench_config(w['desc'].replace('%d pairs' % len(w['o_len']), '%d pairs' % (args.pairs or WORKLOAD_PAIRS[kind]), 1),
                               int(args.pairs or WORKLOAD_PAIRS[kind])),
        'sample_pairs_per_step': int(len(w['o_len'])),
        'cpu_baseline': {'value': val, 'unit': 'GCUPS', 'cores': last['cores'], 'kind': last['kind'],
                         'sample': last['sample']},
        'pairs_per_s': last['pairs_per_s'],
        'e2e': {'value': val, 'unit': 'GCUPS', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    emit(line)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def pinned_copy(eng, a):
    p = eng.pinned_empty(len(a), a.dtype)
    p[:] = a
    return p


def measure(eng, w, steps, warmup, sampler_gpu, do_e2e=True, compact=True):
    """Returns dict with resident + e2e timings for workload w on this rank."""
    from biseqt_b200 import batch
    N = len(w['o_len'])
    kw = dict(dmin=w.get('dmin'), dmax=w.get('dmax'))
    args = (w['mode'], w['type'], w['subst'], w['go'], w['ge'])
    # ---- device-resident: upload once, run W + K times -----------------
    rb = eng.upload(w['codes'], w['o_off'], w['o_len'], w['m_off'], w['m_len'], *args, **kw)
    sampler = ClockSampler(sampler_gpu).start()
    for _ in range(warmup):
        rb.run()
    barrier()
    sampler.mark_begin()
    res_ms, fill_ms, launches = [], [], 0
    t = None
    for _ in range(steps):
        t = rb.run()
        res_ms.append(t['total_ms'])
        fill_ms.append(t['fill_ms'])
        launches += t['fill_launches']
    sampler.mark_end()
    clocks = sampler.stop()
    barrier()
    cells = t['cells']
    out = {'cells': cells, 'pairs': N, 'resident_ms': float(np.sum(res_ms)), 'fill_ms': float(np.sum(fill_ms)),
           'launches': launches, 'clocks': clocks, 'dtype': batch.DTYPE_NAMES[t['dtype']],
           'ptr_bytes': t['ptr_bytes'], 'chunks': t['chunks']}
    rb.release()
    # ---- end to end: pinned host buffers in, pinned host buffers out ------
    if do_e2e:
        out.update(measure_e2e(eng, w, steps, warmup, compact=compact))
        launches += out.pop('e2e_launches')
        out['launches'] = launches
    return out


def measure_e2e(eng, w, steps, warmup, compact):
    """The public batch call (Engine.align -> bsq_align_batch) from pinned HOST buffers to pinned
    host buffers, H2D and D2H inside the timed region.  compact: the DNA forms of the two big
    arrays -- sequences four letters per byte (BSQ_F_CODES_2BIT), transcripts four ops per byte
    (BSQ_F_TX_2BIT), and no offset arrays when the pairs lie back to back in codes[]."""
    from biseqt_b200 import batch
    N = len(w['o_len'])
    args = (w['mode'], w['type'], w['subst'], w['go'], w['ge'])
    kw = {}
    if w['mode'] == 1:
        kw = dict(dmin=pinned_copy(eng, w['dmin']), dmax=pinned_copy(eng, w['dmax']))
    cap = int(w['o_len'].astype(np.int64).sum() + w['m_len'].astype(np.int64).sum())
    pin = {k: pinned_copy(eng, np.ascontiguousarray(w[k])) for k in ('o_len', 'm_len')}
    if compact:
        pin['codes'] = pinned_copy(eng, batch.pack_codes_2bit(w['codes']))
        kw.update(codes_2bit=True, ncodes=len(w['codes']), tx_2bit=True)
        cap = cap // 4 + 64
    else:
        pin['codes'] = pinned_copy(eng, np.ascontiguousarray(w['codes']))
    if compact and w.get('back_to_back'):
        pin['o_off'] = pin['m_off'] = None
    else:
        pin['o_off'], pin['m_off'] = pinned_copy(eng, w['o_off']), pinned_copy(eng, w['m_off'])
    outb = {'score': eng.pinned_empty(N, np.float64), 'tx_off': eng.pinned_empty(N + 1, np.int64),
            'tx_buf': eng.pinned_empty(cap, np.uint8)}
    for k in ('end_x', 'end_y', 'start_x', 'start_y', 'status'):
        outb[k] = eng.pinned_empty(N, np.int32)
    call = lambda: eng.align(pin['codes'], pin['o_off'], pin['o_len'], pin['m_off'], pin['m_len'], *args, out=outb, **kw)   # noqa: E731
    for _ in range(max(1, warmup - 1)):
        call()
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = launches = 0
    plan_ms, call_ms = [], []
    for _ in range(steps):
        r = call()
        tt = eng.timing()
        h2d, d2h = tt['h2d_bytes'], tt['d2h_bytes']
        launches += tt['fill_launches']
        plan_ms.append(tt['plan_ms'])
        call_ms.append(tt['call_ms'])
    out = {'e2e_s': time.perf_counter() - t0}
    out['e2e_breakdown_ms'] = {'python_call': 1e3 * out['e2e_s'] / steps, 'c_call': float(np.mean(call_ms)),
                               'host_plan': float(np.mean(plan_ms)), 'device_span': tt['total_ms'],
                               'fill_kernels': tt['fill_ms'], 'chunks': tt['chunks']}
    out['h2d'], out['d2h'] = h2d, d2h
    out['ok_frac'] = float((r.status == 0).mean())
    out['result'] = r          # the last timed call's output (pinned buffers), for the parity verdict
    out['e2e_launches'] = launches
    out['e2e_form'] = ('2-bit packed sequences in, 2-bit packed transcripts out' +
                       (', pairs back to back (no offset arrays)' if pin['o_off'] is None else '')) if compact \
        else 'one byte per letter in, one byte per op out, int64 offsets'
    barrier()
    return out


def sharded_batch_sweep(ndev, steps, npairs):
    """SURVEY.md section 8(e)'s named sweep: ONE config-3 batch (long-read pairs, B_OVERLAP, radius 64)
    handed to ONE process, which shards it over `ndev` GPUs (bsq_multi: contiguous ranges of equal
    work, one host thread + context per GPU, results straight into slices of the caller's arrays;
    no collective).  End to end from pinned host buffers, compact forms, wall clock around the call."""
    from biseqt_b200 import batch
    eng = batch.Engine(0)
    w = bw.make_device(eng, 'c3_int', npairs, seed=11)
    eng.close()
    me = batch.MultiEngine(ndev)
    try:
        N = len(w['o_len'])
        pin = {'codes': me.pinned_empty(0, np.uint8)}
        packed = batch.pack_codes_2bit(w['codes'])
        pin['codes'] = me.pinned_empty(len(packed), np.uint8)
        pin['codes'][:] = packed
        for k in ('o_len', 'm_len', 'dmin', 'dmax'):
            pin[k] = me.pinned_empty(N, np.int32)
            pin[k][:] = w[k]
        outb = {'score': me.pinned_empty(N, np.float64), 'tx_off': me.pinned_empty(N + 1, np.int64),
                'tx_len': me.pinned_empty(N, np.int32),
                'tx_buf': me.pinned_empty(me.tx_capacity(w['o_len'], w['m_len'], tx_2bit=True), np.uint8)}
        for k in ('end_x', 'end_y', 'start_x', 'start_y', 'status'):
            outb[k] = me.pinned_empty(N, np.int32)
        call = lambda: me.align(pin['codes'], None, pin['o_len'], None, pin['m_len'], w['mode'], w['type'], w['subst'],   # noqa: E731
                                w['go'], w['ge'], dmin=pin['dmin'], dmax=pin['dmax'], out=outb, codes_2bit=True,
                                ncodes=len(w['codes']), tx_2bit=True)
        for _ in range(2):
            r = call()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = call()
        dt = (time.perf_counter() - t0) / steps
        sh = me.shards()
        cells = sum(t['cells'] for _, _, _, t in sh)
        return {'workload': w['desc'], 'n_gpus': ndev, 'gcups_e2e': cells / dt / 1e9, 'ms_per_call': 1e3 * dt,
                'pairs_per_s': N / dt, 'status_ok_fraction': float((r.status == 0).mean()),
                'shards': [{'device': d, 'pairs': int(c), 'device_span_ms': t['total_ms'], 'fill_ms': t['fill_ms'],
                            'h2d_bytes': int(t['h2d_bytes']), 'd2h_bytes': int(t['d2h_bytes'])} for _, c, d, t in sh],
                'note': 'one process, one batch, bsq_multi_align_batch; wall clock around the call, pinned host '
                        'buffers, 2-bit sequences in / 2-bit transcripts out'}
    finally:
        me.close()


_dist = None


def barrier():
    if _dist is not None:
        import torch
        torch.cuda.synchronize()
        _dist.barrier()


def allreduce(val, op):
    if _dist is None:
        return val
    import torch
    t = torch.tensor([float(val)], dtype=torch.float64, device='cuda')
    _dist.all_reduce(t, op=getattr(_dist.ReduceOp, op))
    return float(t.item())


def main():
    global _dist
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2_int', choices=sorted(WORKLOAD_PAIRS))
    ap.add_argument('--pairs', type=int, default=0, help='pairs per GPU (default: the config\'s size)')
    ap.add_argument('--no-extras', action='store_true', help='skip the secondary workloads')
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline')
    host_group = None
    ap.add_argument('--extras', default='c2_logodds,c3_int,c3_logodds,c3_blocal_int,c3_blocal_logodds,c4_int,c5_int,c5_logodds,c5s_int,c1',
                    help='comma-separated secondary workloads (single GPU only)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    # exactly ONE line goes to stdout (the JSON); libraries that print there
    # (NCCL's version banner, pwlib's printf) are routed to stderr meanwhile
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj))
        sys.stdout.flush()
        os.dup2(2, 1)

    if args.impl == 'reference':
        run_reference_arm(args, rank, world, emit)
        return

    from biseqt_b200 import batch
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('nccl')
        _dist = dist
        # a CPU-side group for the last wait: ranks parked in an NCCL barrier spin a kernel on their GPU,
        # which slowed the one-batch-over-N-GPUs sweep on those GPUs 2.2x (measured)
        host_group = dist.new_group(backend='gloo')
    eng = batch.Engine(local)

    kind = args.workload
    npairs = args.pairs or WORKLOAD_PAIRS[kind]
    w = bw.make_device(eng, kind, npairs, seed=2 + 1000 * rank)     # every rank aligns its own shard
    m = measure(eng, w, args.steps, args.warmup, local)

    # whole-job numbers: all ranks' cells over the slowest rank's time
    cells_all = allreduce(m['cells'], 'SUM')
    pairs_all = allreduce(m['pairs'], 'SUM')
    res_ms = allreduce(m['resident_ms'], 'MAX')
    e2e_s = allreduce(m['e2e_s'], 'MAX')
    launches = allreduce(m['launches'], 'SUM')
    value = cells_all * args.steps / (res_ms * 1e-3) / 1e9
    e2e = cells_all * args.steps / e2e_s / 1e9

    line = None
    if rank == 0:
        # roofline of the dominant kernel (DP fill): ALU-pipe bound, SURVEY.md section 8(d)
        opc = ops_per_cell(w['go'])
        which = {'i32': 'IADD', 'i16x2': 'VADD2'}.get(m['dtype'], 'DADD')
        try:
            pipe = {k: eng.microbench(k) for k in ('DADD', 'DMAX', 'DSETP', 'IADD', 'VADD2', 'VIMAX3', 'VIADDMAX', 'SHFL', 'LDS64')}
        except Exception as e:   # never hide the bench line behind a microbenchmark problem
            pipe = {'error': str(e)}
        if 'error' not in pipe and pipe.get('DADD', 0) > 5e12:   # (a run under ncu measures nonsense: do not store it)
            try:   # stored next to the other measured peaks (gpurun_out/ travels back; copied to profiles/)
                os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
                with open(os.path.join(ROOT, 'gpurun_out', 'ALU_PEAKS.json'), 'w') as f:
                    json.dump({'unit': 'thread-ops/s, whole GPU', 'how': 'bsq_microbench: 8 independent chains per thread, '
                               '8 x 256 threads per SM, best of 3 after a warm-up, CUDA events',
                               'sm_clock_mhz_during_bench': m['clocks'].get('sm_mhz'),
                               'rates': pipe, 'pipes': PIPE_NOTES}, f, indent=1)
            except Exception:
                pass
        fill_s = m['fill_ms'] * 1e-3
        achieved = m['cells'] * args.steps * opc / fill_s / 1e9
        peak = pipe.get(which, 0.0) / 1e9 if 'error' not in pipe else None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        roofline = {
            'bound': 'alu', 'achieved': achieved, 'peak': peak, 'unit': 'Gop/s',
            'frac': (achieved / peak) if peak else None,
            # DRAM bytes (read + write) of ONE fill launch: read from the file the round's ncu --set full
            # capture of this kernel on this workload was summarised into (profiles/r02_k1h_traffic.json,
            # made from profiles/r02_k1h_ncu.txt); algorithmic pointer bytes of a launch (125k pairs x
            # 0.5 B/cell) are 1.43e9
            'traffic': k1h_traffic() if (kind == 'c2_int' and npairs == 1000000 and m['dtype'] == 'i16x2') else None,
            'ops_per_cell': opc, 'pipe': which,
            'note': 'achieved = cells x algorithmic ops/cell / fill-kernel time (CUDA events on the library stream); '
                    'peak = %s issue rate measured in this run (bsq_microbench)' % {'IADD': 'add.s32', 'VADD2': 'packed 16-bit add (VIADD.16x2, two adds per instruction)', 'DADD': 'add.f64'}[which],
            'pipe_rates_gops': {k: v / 1e9 for k, v in pipe.items()} if 'error' not in pipe else pipe,
            'pipe_notes': {which: PIPE_NOTES[which]},
            'hbm': {'achieved': m['ptr_bytes'] * args.steps / fill_s / 1e9, 'peak': hbm_peak,
                    'unit': 'GB/s', 'frac': m['ptr_bytes'] * args.steps / fill_s / 1e9 / hbm_peak,
                    'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback',
                    'note': 'traceback-pointer bytes written by the fill kernel / fill time'},
        }
        line = {
            'metric': 'GCUPS', 'value': value, 'unit': 'GCUPS', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': res_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': m['dtype'], 'data': 'synthetic',
            'config': bench_config(w['desc'], npairs),
            'l2': 'inputs larger than L2 (%.0f MB of sequences + %.1f GB of pointer tables per step)'
                  % (len(w['codes']) / 1e6, m['ptr_bytes'] / 1e9),
            'timed': 'fill + traceback + transcript packing, inputs/outputs resident in HBM',
            'pairs_per_s': pairs_all * args.steps / (res_ms * 1e-3),
            'fill_only_gcups': cells_all * args.steps / (allreduce(m['fill_ms'], 'MAX') * 1e-3) / 1e9 if world == 1 else None,
            'clocks': m['clocks'],
            'e2e': {'value': e2e, 'unit': 'GCUPS', 'h2d_bytes_per_step': m['h2d'], 'd2h_bytes_per_step': m['d2h'],
                    'pairs_per_s': pairs_all * args.steps / e2e_s, 'breakdown_ms': m.get('e2e_breakdown_ms'),
                    'form': m.get('e2e_form'),
                    'note': 'bsq_align_batch from pinned host arrays: plan + H2D + fill + traceback + pack + D2H'},
            'gpu_launches': int(launches),
            'roofline': roofline,
            'roofline_frac_solve_plus_traceback': (m['cells'] * args.steps * opc / (m['resident_ms'] * 1e-3) / 1e9 / peak) if peak else None,
            'status_ok_fraction': m['ok_frac'],
        }

    if rank == 0 and world == 1:   # the same call with the one-byte forms of both arrays, for comparison
        try:
            mb = measure_e2e(eng, w, max(2, args.steps // 4), 2, compact=False)
            line['e2e_byte_forms'] = {'value': m['cells'] * max(2, args.steps // 4) / mb['e2e_s'] / 1e9, 'unit': 'GCUPS',
                                      'h2d_bytes_per_step': mb['h2d'], 'd2h_bytes_per_step': mb['d2h'], 'form': mb['e2e_form']}
        except Exception as e:
            line['e2e_byte_forms'] = {'error': str(e)}

    # secondary workloads (rank 0, single GPU): f64 path and the banded kernels
    if rank == 0 and world == 1 and not args.no_extras:
        extras = {}
        made = {}
        for k2 in [k for k in args.extras.split(',') if k]:
            n2 = min(WORKLOAD_PAIRS[k2], npairs) if k2.startswith('c2') else WORKLOAD_PAIRS[k2]
            if args.pairs and not k2.startswith('c2') and WORKLOAD_PAIRS[k2] > 1:
                n2 = max(1, min(n2, args.pairs * WORKLOAD_PAIRS[k2] // WORKLOAD_PAIRS[kind]))   # --pairs scales every batch
            try:
                if k2 in ('c3_logodds', 'c3_blocal_int', 'c3_blocal_logodds') and 'c3_int' in made:
                    # the same pairs and bands as c3_int; other scores and / or banded LOCAL
                    w2 = dict(made['c3_int'])
                    if k2.endswith('logodds'):
                        S3, go3, ge3 = bw.st.log_odds_scores(4, 0.03, 0.06, 0.12)
                        w2.update(subst=S3, go=go3, ge=ge3)
                    if 'blocal' in k2:
                        w2.update(type=bw.B_LOCAL)
                    w2.update(kind=k2, desc=w2['desc'].replace('B_OVERLAP', 'B_LOCAL' if 'blocal' in k2 else 'B_OVERLAP')
                              .replace('int +2/-3 go0 ge-2', 'log-odds' if k2.endswith('logodds') else 'int +2/-3 go0 ge-2'))
                else:
                    w2 = bw.make_device(eng, k2, n2, seed=3)
                    made[k2] = w2
                st2 = max(2, args.steps // 2)
                m2 = measure(eng, w2, st2, 3 if not k2.startswith('c5') else 1, local, do_e2e=True)
                opc2 = ops_per_cell(w2['go'])
                which2 = {'i32': 'IADD', 'i16x2': 'VADD2'}.get(m2['dtype'], 'DADD')
                pk = roofline['pipe_rates_gops'].get(which2) if isinstance(roofline['pipe_rates_gops'], dict) else None
                ach = m2['cells'] * st2 * opc2 / (m2['fill_ms'] * 1e-3) / 1e9
                extras[k2] = {'workload': w2['desc'], 'dtype': m2['dtype'], 'e2e_form': m2.get('e2e_form'),
                              'h2d_bytes_per_step': m2.get('h2d'), 'd2h_bytes_per_step': m2.get('d2h'),
                              'gcups': m2['cells'] * st2 / (m2['resident_ms'] * 1e-3) / 1e9,
                              'fill_only_gcups': m2['cells'] * st2 / (m2['fill_ms'] * 1e-3) / 1e9,
                              'e2e_gcups': m2['cells'] * st2 / m2['e2e_s'] / 1e9,
                              'roofline_frac': ach / pk if pk else None, 'ops_per_cell': opc2,
                              # SURVEY.md 8(d) defines the metric on solve + traceback time: the same
                              # fraction with the whole resident step (fill + traceback + packing) as time
                              'roofline_frac_solve_plus_traceback':
                                  (m2['cells'] * st2 * opc2 / (m2['resident_ms'] * 1e-3) / 1e9) / pk if pk else None,
                              'status_ok_fraction': m2['ok_frac'], 'chunks': m2['chunks'],
                              'e2e_breakdown_ms': m2.get('e2e_breakdown_ms')}
            except Exception as e:
                extras[k2] = {'error': str(e)}
        line['other_workloads'] = extras

    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cores = os.cpu_count() or 1
            per_core = cpu_sample_per_core(kind, cores)
            cpu = cpu_reference_run(w, per_core, cores, keep=True)
            line['cpu_baseline'] = {k: v for k, v in cpu.items()
                                    if k in ('value', 'unit', 'cores', 'kind', 'sample', 'pairs_per_s')}
            # the reference's verdict on the very pairs it just timed vs the GPU arm's e2e output
            line['parity_spotcheck'] = parity_verdict(w, cpu['results'], m['result'])
            from oracle import pwref
            if pwref.available('g'):   # the reference's own build flags (Makefile:1: -g, no -O), quarter sample
                cg = cpu_reference_run(w, max(1, per_core // 4), cores, variant='g')
                line['cpu_baseline']['reference_flags_build'] = {k: cg[k] for k in ('value', 'unit', 'cores', 'sample', 'pairs_per_s')}
        except Exception as e:
            line['cpu_baseline'] = {'error': str(e)}
    eng.close()
    if _dist is not None:
        _dist.barrier()        # every rank has given its GPU memory back
        import torch
        torch.cuda.synchronize()
    if rank == 0 and not args.no_extras:
        try:    # the one-batch-over-N-GPUs engine (section 8(e)); the other ranks wait at the barrier below
            line['sharded_batch'] = sharded_batch_sweep(world, max(2, args.steps // 4),
                                                        min(WORKLOAD_PAIRS['c3_int'], args.pairs or 10 ** 9))
        except Exception as e:
            line['sharded_batch'] = {'error': str(e)}
    if rank == 0:
        emit(line)
    if _dist is not None:
        _dist.barrier(group=host_group)
        _dist.destroy_process_group()


if __name__ == '__main__':
    main()
