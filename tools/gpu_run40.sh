mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02N_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02N_pytest.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r02N_bench.json 2> gpurun_out/r02N_bench.err; echo "rc=$?" >> gpurun_out/r02N_bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02N_smoke.log 2>&1
tail -3 gpurun_out/r02N_pytest.log; tail -2 gpurun_out/r02N_bench.err; tail -1 gpurun_out/r02N_smoke.log
python tools/show_bench.py gpurun_out/r02N_bench.json | cut -c1-330
