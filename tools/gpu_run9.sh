mkdir -p gpurun_out
L=gpurun_out/r02i_kbench.log
: > $L
for k in c3_int c3_logodds c3_blocal_int; do python tools/kbench.py $k 32768 4 >> $L 2>&1; done
python tools/kbench.py c4_int 4096 3 >> $L 2>&1
BSQ_TRACE_KERNEL=w python tools/kbench.py c4_int 4096 3 >> $L 2>&1
BSQ_TRACE_KERNEL=w python tools/kbench.py c3_int 32768 3 >> $L 2>&1
python tools/kbench.py c3_int 2048 3 >> $L 2>&1
BSQ_TRACE_KERNEL=n python tools/kbench.py c3_int 2048 3 >> $L 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not c5 and not big_ragged and not ragged_reads" > gpurun_out/r02i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02i_pytest.log
tail -4 gpurun_out/r02i_pytest.log; cat $L
