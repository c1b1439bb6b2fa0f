mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?" >> gpurun_out/r02b_bench.err
BSQ_PLAN_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/r02b_bench_dbg.json 2> gpurun_out/r02b_bench_dbg.err
tail -3 gpurun_out/r02b_pytest.log; tail -c 400 gpurun_out/r02b_bench.err
