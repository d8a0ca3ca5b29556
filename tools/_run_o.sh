mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra > gpurun_out/r2_o_bench_n2.json 2> gpurun_out/r2_o_bench_n2.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_o_bench_n2.json') if l.startswith('{')][-1])
t=d['train']; print(d['value'], {k:t[k] for k in ('ms_per_step','ms_per_step_tf32_convs','loss_max_rel_diff','gpu_launches_per_step')})
PY
