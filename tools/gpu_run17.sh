mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02p_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r02p_bench_n8.json 2> gpurun_out/r02p_bench_n8.err; echo "rc=$?" >> gpurun_out/r02p_bench_n8.err
tail -c 400 gpurun_out/r02p_bench_n8.err; python tools/show_bench.py gpurun_out/r02p_bench_n8.json | head -12; python tools/show_bench.py gpurun_out/r02p_bench_n8.json | tail -1 | cut -c1-1500
