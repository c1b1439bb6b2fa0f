mkdir -p gpurun_out
L=gpurun_out/r02e_kbench.log
: > $L
python tools/kbench.py c3_int 32768 4 >> $L 2>&1
python tools/kbench.py c3_logodds 32768 3 >> $L 2>&1
python tools/kbench.py c4_int 4096 3 >> $L 2>&1
python tools/kbench.py c3_int 100000 3 >> $L 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02e_launches_c3.csv python tools/kbench.py c3_int 32768 2 > gpurun_out/r02e_ncu_c3.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "traceback_kernels or bench_shapes or banded or wide or full_size or compact or multi" > gpurun_out/r02e_pytest.log 2>&1
cat $L; tail -3 gpurun_out/r02e_pytest.log; grep -E "walk|k2_fill|pack" gpurun_out/r02e_launches_c3.csv | awk -F'","' '{print $5, $NF}' | head
