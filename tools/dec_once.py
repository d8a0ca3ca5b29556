#!/usr/bin/env python
"""One encoder pass + `reps` decoder launches over an nx^3 grid (profiling target for ncu: -k regex:decoder_tc_kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slice3d_b200 import Slices3DRegModel, synth  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = "cuda:0"
m = Slices3DRegModel(256, 12, "test")
m.load_state_dict(synth.synthetic_state_dict(m.state_dict(), 0))
m = m.to(dev).eval()
feed = synth.synthetic_inputs(256, 12, 0)
nat = m.native()
if os.environ.get("DEC_BENCH_DBG"):
    from slice3d_b200 import _native as _n
    _n.lib().s3d_debug_set_decoder_flags(int(os.environ["DEC_BENCH_DBG"]))
planes = nat.encode(feed["img_input"].to(dev))
ax = torch.linspace(-0.5, 0.5, nx).to(dev)
T = feed["trans_mat_wo_rot_tp"][0].to(dev)
out = torch.empty(nx ** 3, device=dev)
for _ in range(reps):
    nat.decode_grid(planes, 0, (ax, ax, ax), 0, nx ** 3, T, precision=prec, out=out)
torch.cuda.synchronize()
print("done", float(out.abs().max()))
