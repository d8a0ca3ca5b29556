mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_f_pytest.log
tail -4 gpurun_out/r2_f_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_f_bench_n2.json 2> gpurun_out/r2_f_bench_n2.err; tail -c 5000 gpurun_out/r2_f_bench_n2.json; tail -5 gpurun_out/r2_f_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2_f_ref_n2.json 2> gpurun_out/r2_f_ref_n2.err; cat gpurun_out/r2_f_ref_n2.json | cut -c1-600
