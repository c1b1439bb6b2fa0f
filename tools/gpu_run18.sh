mkdir -p gpurun_out
L=gpurun_out/r02q_kbench.log
: > $L
for k in c3_int c3_blocal_int c3_logodds; do python tools/kbench.py $k 32768 4 >> $L 2>&1; done
python tools/kbench.py c4_int 4096 3 >> $L 2>&1
python tools/kbench.py c5_int 1 3 >> $L 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "traceback_kernels or bench_shapes or banded or wide or full_size or compact or multi or batches or chunked" > gpurun_out/r02q_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02q_pytest.log
tail -3 gpurun_out/r02q_pytest.log; cat $L
