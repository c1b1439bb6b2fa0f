mkdir -p gpurun_out
L=gpurun_out/r02C_kbench.log
: > $L
python tools/kbench.py c5_int 1 3 >> $L 2>&1
python tools/kbench.py c5_logodds 1 2 >> $L 2>&1
python tools/kbench.py c5s_int 64 2 >> $L 2>&1
python tools/kbench.py c1n 64 3 >> $L 2>&1
python tools/api_latency.py >> $L 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "traceback_kernels or bench_shapes or wide or full_size or std_as_band or forced or k3" > gpurun_out/r02C_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02C_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02C_launches_c5.csv python tools/kbench.py c5_int 1 1 > gpurun_out/r02C_ncu_c5.log 2>&1
tail -3 gpurun_out/r02C_pytest.log; cat $L; grep -o '"bsq_[a-z0-9_]*[^"]*","[^"]*","[^"]*","[0-9.]*"$' gpurun_out/r02C_launches_c5.csv | tail -6
