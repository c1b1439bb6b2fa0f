mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02Q_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02Q_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02Q_smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r02Q_bench.json 2> gpurun_out/r02Q_bench.err; echo "rc=$?" >> gpurun_out/r02Q_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02Q_bench_ref.json 2> gpurun_out/r02Q_bench_ref.err; echo "rc=$?" >> gpurun_out/r02Q_bench_ref.err
tail -3 gpurun_out/r02Q_pytest.log; tail -1 gpurun_out/r02Q_smoke.log; tail -1 gpurun_out/r02Q_bench.err; tail -1 gpurun_out/r02Q_bench_ref.err
python tools/show_bench.py gpurun_out/r02Q_bench.json 2>/dev/null | grep -E "^value|^e2e |^parity|c3_blocal_int" | cut -c1-260; cut -c1-300 gpurun_out/r02Q_bench_ref.json
