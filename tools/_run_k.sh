mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_k_bench_n8.json 2> gpurun_out/r2_k_bench_n8.err; tail -c 4500 gpurun_out/r2_k_bench_n8.json; tail -3 gpurun_out/r2_k_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 4 --steps 3 --warmup 3 --no-train > gpurun_out/r2_k_bench_n4.json 2> gpurun_out/r2_k_bench_n4.err; cut -c1-300 gpurun_out/r2_k_bench_n4.json
