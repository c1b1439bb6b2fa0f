import sys; sys.path.insert(0,'/root/repo')
from biseqt_b200 import batch
e=batch.Engine(0)
for k in ('IADD','VADD2','VIMAX3','VIADDMAX'):
    print(k, e.microbench(k)/1e9)
