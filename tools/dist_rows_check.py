#!/usr/bin/env python
"""torchrun check of the row-granular multi-GPU slabs (Generator3D.slab_unit = "row"): the gathered volume must equal,
bit for bit, the volume one rank computes alone, whatever the slab boundaries are.  Run with
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_rows_check.py [nx]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slice3d_b200 import Generator3D, Slices3DRegModel, synth  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
S, K = 128, 12
m = Slices3DRegModel(S, K, "test", precision="fp16f8")
m.load_state_dict(synth.synthetic_state_dict(m.state_dict(), 0))
m = m.to(dev).eval()
feed = {k: v.to(dev) for k, v in synth.synthetic_inputs(S, K, 0).items()}
gen = Generator3D(m, upsampling_steps=0, resolution0=nx, pred_type="sdf")
gen.slab_unit = "row"
nat = m.native()
ok = True
with torch.no_grad():
    planes = m.encode(feed["img_input"])
    ax = gen.grid_axes(nx, dev)
    alone = torch.empty(nx ** 3, device=dev)
    nat.decode_grid(planes, 0, (ax, ax, ax), 0, nx ** 3, feed["trans_mat_wo_rot_tp"][0], out_scale=-1.0, precision="fp16f8", out=alone)
    for it, skew in enumerate([None, 1.37, 0.61, 1.003, 1.0]):
        if skew is not None:  # pretend rank 0 was `skew` times as fast as it was: the next call's boundaries move
            e0, e1, p, a, b = gen._last_dec
            gen._last_dec = (e0, e1, p * (skew if rank == 0 else 1.0), a, b)
        vol = gen.generate_grid(feed, resolution=nx, precision="fp16f8", as_numpy=False)
        torch.cuda.synchronize()
        same = bool(torch.equal(vol.view(-1), alone))
        e0, e1, p = gen._last_dec[:3]
        print(f"rank {rank} call {it}: share {p:.4f} planes of {nx}, decoder {e0.elapsed_time(e1):.2f} ms, identical to the "
              f"single-rank volume: {same}", flush=True)
        ok = ok and same
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
if rank == 0:
    print("ROWS CHECK", "PASSED" if int(flag.item()) else "FAILED")
sys.exit(0 if int(flag.item()) else 3)
