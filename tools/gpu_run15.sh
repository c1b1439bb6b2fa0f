mkdir -p gpurun_out
L=gpurun_out/r02n_kbench.log
: > $L
BSQ_K3_CFG=0 timeout 300 python tools/kbench.py c5_int 1 3 >> $L 2>&1
BSQ_K3_CFG=0 timeout 300 python tools/kbench.py c5_logodds 1 2 >> $L 2>&1
timeout 300 python tools/kbench.py c1 1 5 >> $L 2>&1
timeout 300 python tools/api_latency.py >> $L 2>&1
BSQ_K3_CFG=0 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "c5 or forced_k3 or std_as_band or very_wide" > gpurun_out/r02n_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02n_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02n_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02n_smoke.log
tail -3 gpurun_out/r02n_pytest.log; tail -8 gpurun_out/r02n_smoke.log; cat $L
