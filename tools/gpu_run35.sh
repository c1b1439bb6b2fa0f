mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r02I_bench_n8.json 2> gpurun_out/r02I_bench_n8.err; echo "rc=$?" >> gpurun_out/r02I_bench_n8.err
tail -2 gpurun_out/r02I_bench_n8.err
python tools/show_bench.py gpurun_out/r02I_bench_n8.json 2>/dev/null | grep -E "^value|^e2e |^sharded" | cut -c1-300
