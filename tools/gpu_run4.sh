mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "multi or compact or lifetime or generator or edge" > gpurun_out/r02d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02d_launches_c3.csv python tools/kbench.py c3_int 32768 2 > gpurun_out/r02d_ncu_c3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02d_launches_c2.csv python tools/kbench.py c2_int 1000000 2 > gpurun_out/r02d_ncu_c2.log 2>&1
tail -3 gpurun_out/r02d_pytest.log
python - <<'PY'
import csv,collections
for f in ('gpurun_out/r02d_launches_c3.csv','gpurun_out/r02d_launches_c2.csv'):
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    agg=collections.OrderedDict()
    for r in rows:
        name=r[4][:60]; v=float(r[-1].replace(',',''))
        agg.setdefault(name,[]).append(v)
    print(f)
    for k,v in agg.items(): print('  %-60s n=%3d  sum %.3f ms  each %.3f ms'%(k,len(v),sum(v)/1e6,sum(v)/len(v)/1e6))
PY
