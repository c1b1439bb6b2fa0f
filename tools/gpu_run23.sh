mkdir -p gpurun_out
L=gpurun_out/r02w_kbench.log
: > $L
python tools/kbench.py c5_int 1 3 >> $L 2>&1
python tools/kbench.py c5_int 10 2 >> $L 2>&1
python tools/kbench.py c5s_int 64 2 >> $L 2>&1
python tools/kbench.py c5s_logodds 64 2 >> $L 2>&1
BSQ_K3_CFG=1 python tools/kbench.py c5s_int 64 2 >> $L 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "traceback_kernels or bench_shapes or wide or full_size or std_as_band or forced or k3" > gpurun_out/r02w_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_pytest.log
tail -3 gpurun_out/r02w_pytest.log; cat $L
