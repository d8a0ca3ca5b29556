#!/usr/bin/env python
"""Phase timing of one training step (BASELINE configs[4] shapes) with and without DDP: forward / backward / Adam.
    python tools/train_bench.py [plain|ddp|ddp_static|ddp_nobuf]       (single process)
    torchrun --nproc-per-node N tools/train_bench.py ddp"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slice3d_b200 import Slices3DRegModel, synth  # noqa: E402
from slice3d_b200 import train as T  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if mode != "plain":
    if "MASTER_ADDR" not in os.environ:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29655", RANK="0", WORLD_SIZE="1")
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.allow_tf32 = "tf32" in mode
torch.backends.cuda.matmul.allow_tf32 = False
S, K, B, NQ = 128, 12, 4, 256
m = Slices3DRegModel(S, K, "train")
m.load_state_dict(synth.synthetic_state_dict(m.state_dict(), 8))
m = m.to(dev).train()
if "torchdec" in mode:
    m.native_train = False
net = m
if mode.startswith("ddp"):
    from torch.nn.parallel import DistributedDataParallel as DDP
    kw = dict(device_ids=[local], find_unused_parameters=True)
    if "static" in mode:
        kw["static_graph"] = True
    if "nobuf" in mode:
        kw["broadcast_buffers"] = False
    if "freeze" in mode:
        for n, p in m.named_parameters():
            if n.startswith("att_layer.") or ".down5_." in n:
                p.requires_grad_(False)
        kw["find_unused_parameters"] = False
    net = DDP(m, **kw)
opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=3e-4, fused="fused" in mode)
host = {k: v.pin_memory() for k, v in synth.synthetic_train_batch(S, K, B, NQ, seed=100 + rank).items()}


def step(ev):
    batch = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    ev[0].record()
    opt.zero_grad()
    x = net(batch)
    lp, li, lv = T.cal_loss_pred(x, batch)
    loss = lp + li + lv
    ev[1].record()
    loss.backward()
    ev[2].record()
    opt.step()
    ev[3].record()
    return lp.item()


for _ in range(4):
    step([torch.cuda.Event(enable_timing=True) for _ in range(4)])
torch.cuda.synchronize()
tot = [0.0, 0.0, 0.0]
t0 = time.perf_counter()
N = 6
for _ in range(N):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    step(ev)
    torch.cuda.synchronize()
    for i in range(3):
        tot[i] += ev[i].elapsed_time(ev[i + 1])
wall = (time.perf_counter() - t0) / N * 1e3
if rank == 0:
    print(f"{mode} world {world}: fwd {tot[0] / N:.1f} ms, bwd {tot[1] / N:.1f} ms, adam {tot[2] / N:.1f} ms, wall {wall:.1f} ms/step", flush=True)
if mode != "plain":
    dist.destroy_process_group()
