mkdir -p gpurun_out
L=gpurun_out/r02O_kbench.log
: > $L
python tools/kbench.py c3_blocal_int 32768 3 >> $L 2>&1
python tools/kbench.py c3_blocal_int 100000 3 >> $L 2>&1
python tools/kbench.py c3_int 32768 3 >> $L 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "bench_shapes or banded or batches or blocal or local" > gpurun_out/r02O_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02O_pytest.log
tail -3 gpurun_out/r02O_pytest.log; grep -E "^c3|fill_ms|total_ms" $L | cut -c1-120
