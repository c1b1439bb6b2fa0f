mkdir -p gpurun_out
L=gpurun_out/r02c_kbench.log
: > $L
for k in c3_int c3_logodds c3_blocal_int c3_blocal_logodds; do python tools/kbench.py $k 32768 4 >> $L 2>&1; python tools/kbench.py $k 32768 3 notrace >> $L 2>&1; done
echo "--- narrow walker for comparison" >> $L
BSQ_TRACE_KERNEL=narrow python tools/kbench.py c3_int 32768 4 >> $L 2>&1
python tools/kbench.py c4_int 4096 3 >> $L 2>&1
python tools/kbench.py c4_int 4096 3 notrace >> $L 2>&1
python tools/kbench.py c2_int 1000000 4 >> $L 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "traceback_kernels or bench_shapes or banded or wide or batches or full_size" > gpurun_out/r02c_pytest.log 2>&1
cat $L; tail -3 gpurun_out/r02c_pytest.log
