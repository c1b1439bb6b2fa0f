mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02a_gpu.txt; nproc >> gpurun_out/r02a_gpu.txt; free -g >> gpurun_out/r02a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?" >> gpurun_out/r02a_bench.err
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_cases.py > gpurun_out/r02a_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_memcheck.log
timeout 500 compute-sanitizer --tool racecheck python tools/sanitize_cases.py > gpurun_out/r02a_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_racecheck.log
timeout 300 compute-sanitizer --tool synccheck python tools/sanitize_cases.py > gpurun_out/r02a_synccheck.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_synccheck.log
tail -3 gpurun_out/r02a_pytest.log; tail -c 600 gpurun_out/r02a_bench.err; tail -5 gpurun_out/r02a_memcheck.log; tail -5 gpurun_out/r02a_racecheck.log; tail -3 gpurun_out/r02a_synccheck.log
