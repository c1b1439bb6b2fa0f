mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
timeout 1200 python bench.py --steps 8 --warmup 3 > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "bench rc=$?" >> gpurun_out/r02r_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02r_bench_ref.json 2> gpurun_out/r02r_bench_ref.err
tail -3 gpurun_out/r02r_pytest.log; tail -c 600 gpurun_out/r02r_bench.err
