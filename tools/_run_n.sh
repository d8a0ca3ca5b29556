mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_n_bench_n2.json 2> gpurun_out/r2_n_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r2_n_bench_n2.json')); print(d['value'], d['e2e']['value'], d['roofline']['per_rank']); t=d['train']; print({k:t[k] for k in ('ms_per_step','ms_per_step_tf32_convs','loss_max_rel_diff','gpu_launches_per_step')})"; tail -3 gpurun_out/r2_n_bench_n2.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_n_bench.json 2> gpurun_out/r2_n_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_n_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['inputs']); t=d['train']; print({k:t[k] for k in ('ms_per_step','ms_per_step_tf32_convs','loss_max_rel_diff','gpu_launches_per_step')})"; tail -3 gpurun_out/r2_n_bench.err
