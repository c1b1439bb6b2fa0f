mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02l_gpus.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "multi" > gpurun_out/r02l_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r02l_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r02l_bench_n2.json 2> gpurun_out/r02l_bench_n2.err; echo "rc=$?" >> gpurun_out/r02l_bench_n2.err
tail -3 gpurun_out/r02l_pytest_multi.log; tail -c 1500 gpurun_out/r02l_bench_n2.err; cat gpurun_out/r02l_bench_n2.json | head -c 3000
