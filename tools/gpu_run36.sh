mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02J_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02J_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02J_smoke.log 2>&1
python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/r02J_bench.json 2> gpurun_out/r02J_bench.err; echo "rc=$?" >> gpurun_out/r02J_bench.err
tail -3 gpurun_out/r02J_pytest.log; tail -1 gpurun_out/r02J_smoke.log; tail -1 gpurun_out/r02J_bench.err
python tools/show_bench.py gpurun_out/r02J_bench.json 2>/dev/null | grep -E "^value|^e2e |c5_int|c1 " | cut -c1-200
