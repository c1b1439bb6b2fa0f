"""Multi-GPU path on CPU: world_size-2 gloo processes shard a batch with
bsq_partition_pairs (LPT, no collective on the data path), each rank aligns its
shard on the emulator build, and a host-side gather reassembles the batch."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %(root)r)
    import numpy as np
    import torch.distributed as dist
    from biseqt_b200 import _ffi, batch
    from biseqt_b200._ffi import ffi
    from tests import parity_cases as pc
    from tests.util import oracle_one, same_as
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = _ffi.load(%(emu)r)
    rng = np.random.default_rng(7)            # same batch on every rank
    probs = pc.make_batch(rng, 14, 1, 2, 20, 90, family=1, gomode=-1, band=12)
    o_len = np.array([len(p['origin']) for p in probs], np.int32)
    m_len = np.array([len(p['mutant']) for p in probs], np.int32)
    dmin = np.array([p['diag_range'][0] for p in probs], np.int32)
    dmax = np.array([p['diag_range'][1] for p in probs], np.int32)
    shard = np.zeros(len(probs), np.int32)
    load = np.zeros(world, np.int64)
    c = lambda a, t: ffi.cast(t, a.ctypes.data)
    assert lib.bsq_partition_pairs(1, len(probs), c(o_len, 'int32_t*'), c(m_len, 'int32_t*'),
                                   c(dmin, 'int32_t*'), c(dmax, 'int32_t*'), world,
                                   c(shard, 'int32_t*'), c(load, 'int64_t*')) == 0
    mine = [i for i in range(len(probs)) if shard[i] == rank]
    eng = batch.Engine(0, lib=lib)
    from tests.util import engine_batch
    r = engine_batch(eng, [probs[i] for i in mine])
    for k, i in enumerate(mine):
        same_as(oracle_one(probs[i]), r, k, what='rank %%d pair %%d' %% (rank, i))
    # host-side gather of (index, score); no collective touched the DP path
    out = [None] * world
    dist.all_gather_object(out, [(int(i), float(r.score[k])) for k, i in enumerate(mine)])
    if rank == 0:
        got = dict(x for part in out for x in part)
        assert sorted(got) == list(range(len(probs)))
        for i, p in enumerate(probs):
            assert got[i] == oracle_one(p)['score']
        assert abs(int(load[0]) - int(load[1])) <= max(int(load.max()) // 2, 1)
        print('OK', len(probs), load.tolist())
    dist.barrier()
    dist.destroy_process_group()
''')


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_sharding_gloo(emu_lib, tmp_path):
    emu = os.path.join(ROOT, 'tests', 'simt_emu', '_build', 'libbsqgpu_emu.so')
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT, 'emu': emu})
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert 'OK 14' in res.stdout
