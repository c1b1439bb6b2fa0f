mkdir -p gpurun_out /tmp/ncu
L=gpurun_out/r02u_kbench.log
: > $L
python tools/kbench.py c5_int 1 3 >> $L 2>&1
python tools/kbench.py c5_logodds 1 2 >> $L 2>&1
python tools/api_latency.py >> $L 2>&1
BSQ_K3_CFG=0 python tools/k3prof.py c5_int > gpurun_out/r02u_k3prof.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "traceback_kernels or bench_shapes or wide or full_size or std_as_band or forced or k3" > gpurun_out/r02u_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02u_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bsq_k3l_fill -c 1 -o /tmp/ncu/k3l python tools/kbench.py c5_int 1 1 > gpurun_out/r02u_full_k3l.log 2>&1
python profiles/summarize_ncu.py /tmp/ncu/k3l.ncu-rep > gpurun_out/r02u_k3l_ncu.txt 2>&1
ncu -i /tmp/ncu/k3l.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02u_k3l_source.csv.gz
tail -3 gpurun_out/r02u_pytest.log; cat $L; tail -8 gpurun_out/r02u_k3prof.txt; du -sh gpurun_out
