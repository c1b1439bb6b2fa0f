"""Records the lean oracle's answer for the full-size config-5 problems (one 1 Mb x 1 Mb
B_GLOBAL pair, radius 2048; bench_workloads 'c5_int' / 'c5_logodds', seed 5) as a small
fixture: score, end/start, transcript length, op counts and SHA-256 of the transcript.
The oracle needs ~4 GB and a few CPU-minutes per problem, so this is run once, here;
tests/test_gpu_parity.py compares the GPU result with the fixture.
usage: python tests/golden/make_c5_golden.py"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench_workloads as bw  # noqa: E402
from oracle import lean  # noqa: E402


def main():
    out = {}
    for kind in ('c5_int', 'c5_logodds'):
        w = bw.make(kind, 1, 5)
        t0 = time.time()
        e = lean.align_batch(w['mode'], w['type'], w['subst'], w['go'], w['ge'], w['codes'], w['o_off'],
                             w['o_len'], w['m_off'], w['m_len'], w.get('dmin'), w.get('dmax'))
        tx = lean.transcripts(e)[0]
        out[kind] = dict(seed=5, n=int(w['o_len'][0]), m=int(w['m_len'][0]), status=int(e['status'][0]),
                         score=float(e['score'][0]), score_hex=float(e['score'][0]).hex(),
                         end=[int(e['end_x'][0]), int(e['end_y'][0])],
                         start=[int(e['start_x'][0]), int(e['start_y'][0])], tx_len=len(tx),
                         ops={c: tx.count(c) for c in 'MSID'},
                         tx_sha256=hashlib.sha256(tx.encode()).hexdigest(),
                         oracle_seconds=round(time.time() - t0, 1))
        print(kind, out[kind], flush=True)
    with open(os.path.join(ROOT, 'tests', 'golden', 'c5_golden.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
