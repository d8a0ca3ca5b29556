mkdir -p gpurun_out
for m in plain plain_tf32 plain_torchdec plain_fused ddp ddp_static ddp_nobuf ddp_freeze; do python tools/train_bench.py $m 2>&1 | grep world; done | tee gpurun_out/r2_g_trainbench.log
for m in ddp ddp_freeze ddp_freeze_fused; do python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/train_bench.py $m 2>&1 | grep world; done | tee -a gpurun_out/r2_g_trainbench.log
