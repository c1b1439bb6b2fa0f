mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "very_wide or forced_k3 or c5 or std_as_band or wide_bands or banded_memory or golden or table_scores or edge" > gpurun_out/r02f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02f_pytest.log
L=gpurun_out/r02f_kbench.log
: > $L
timeout 300 python tools/kbench.py c5_int 1 3 >> $L 2>&1
timeout 300 python tools/kbench.py c5_int 1 3 notrace >> $L 2>&1
timeout 300 python tools/kbench.py c5_logodds 1 3 >> $L 2>&1
BSQ_K3_CFG=1 timeout 300 python tools/kbench.py c5_int 1 2 >> $L 2>&1
timeout 300 python tools/kbench.py c1 1 5 >> $L 2>&1
BSQ_NO_STD_BAND=1 timeout 300 python tools/kbench.py c1 1 5 >> $L 2>&1
timeout 300 python tools/api_latency.py >> $L 2>&1
tail -4 gpurun_out/r02f_pytest.log; cat $L
