mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r02H_bench_n2.json 2> gpurun_out/r02H_bench_n2.err; echo "rc=$?" >> gpurun_out/r02H_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02H_ref_n2.json 2> gpurun_out/r02H_ref_n2.err; echo "rc=$?" >> gpurun_out/r02H_ref_n2.err
timeout 600 python -m pytest tests -x -q -m gpu -k "multi" > gpurun_out/r02H_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r02H_pytest_multi.log
tail -2 gpurun_out/r02H_bench_n2.err; tail -2 gpurun_out/r02H_ref_n2.err; tail -3 gpurun_out/r02H_pytest_multi.log
python tools/show_bench.py gpurun_out/r02H_bench_n2.json | head -8 | cut -c1-300; cut -c1-400 gpurun_out/r02H_ref_n2.json
