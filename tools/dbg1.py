import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biseqt_b200 import batch
from oracle import lean
eng = batch.Engine(0)
S = np.zeros(10, np.uint8); J = np.ones(10, np.uint8)
o = np.concatenate([S, J]); t = np.concatenate([J, S])
for n_ in (20, 12, 7):
    oo, tt = o[:n_], t[20 - n_:]
    for dt in (0, -1):
        codes, o_off, o_len, m_off, m_len = batch.pack_pairs([(oo, tt)])
        r = eng.align(codes, o_off, o_len, m_off, m_len, 0, 1, np.eye(4), 0, 0, dtype=dt)
        e = lean.align(oo, tt, np.eye(4), 0, 0, 0, 1)
        print(n_, 'dtype', dt, 'gpu', int(r.status[0]), r.score[0], r.end_x[0], r.end_y[0], r.start_x[0], r.start_y[0], r.transcript(0), '| oracle', e['status'], e['score'], e['end'], e['start'], e['transcript'])
    H = eng.table_scores(oo, tt, 1, np.eye(4), 0, 0)
    print(H[-3:, -6:])
