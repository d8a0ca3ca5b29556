#!/usr/bin/env python
"""All-gather timing probe (torchrun): 64 MiB fp32 volume split in world slabs, NCCL transport from NCCL_DEBUG."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 256 ** 3
vol = torch.zeros(n, device=dev)
slab = n // world
for it in range(6):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dist.all_gather_into_tensor(vol, vol[rank * slab:(rank + 1) * slab])
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"all_gather_into_tensor in place, iter {it}: {e0.elapsed_time(e1):.3f} ms", flush=True)
if rank == 0:
    print("can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1), flush=True)
dist.destroy_process_group()
