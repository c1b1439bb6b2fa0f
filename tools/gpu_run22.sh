mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_pytest.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "rc=$?" >> gpurun_out/r02v_bench.err
L=gpurun_out/r02v_kbench.log
: > $L
for cfg in 0 1; do echo "K3 cfg $cfg, 10 pairs" >> $L; BSQ_K3_CFG=$cfg python tools/kbench.py c5_int 10 2 >> $L 2>&1; done
for cfg in 0 1; do echo "K3 cfg $cfg, 64 x c5s" >> $L; BSQ_K3_CFG=$cfg python tools/kbench.py c5s_int 64 2 >> $L 2>&1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02v_launches_c5.csv python tools/kbench.py c5_int 1 1 > gpurun_out/r02v_ncu_c5.log 2>&1
tail -3 gpurun_out/r02v_pytest.log; tail -2 gpurun_out/r02v_bench.err; cat $L; python tools/show_bench.py gpurun_out/r02v_bench.json | grep -E "^value|^e2e|c5_int|c1 |c3_int" | cut -c1-400
