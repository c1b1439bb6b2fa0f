mkdir -p gpurun_out
L=gpurun_out/r02k_kbench.log
: > $L
python tools/kbench.py c3_blocal_int 32768 4 >> $L 2>&1
python tools/kbench.py c3_blocal_int 32768 3 notrace >> $L 2>&1
python tools/kbench.py c3_int 32768 3 >> $L 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or singles or batches or banded or wide or bench_shapes or full_size or compiled_reference or uniform" > gpurun_out/r02k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02k_pytest.log
tail -4 gpurun_out/r02k_pytest.log; cat $L
