mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r02o_bench_n2.json 2> gpurun_out/r02o_bench_n2.err; echo "rc=$?" >> gpurun_out/r02o_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02o_bench_ref_n2.json 2> gpurun_out/r02o_bench_ref_n2.err
tail -c 600 gpurun_out/r02o_bench_n2.err; python tools/show_bench.py gpurun_out/r02o_bench_n2.json | tail -3; head -c 600 gpurun_out/r02o_bench_ref_n2.json
