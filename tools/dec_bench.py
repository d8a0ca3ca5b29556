#!/usr/bin/env python
"""Decoder-only timing harness for kernel experiments (not the contract bench): times
decode_grid for each precision on an nx^3 grid with CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slice3d_b200 import Slices3DRegModel, synth  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
precs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["bf16x3", "bf16"]
S = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = "cuda:0"
m = Slices3DRegModel(S, 12, "test")
m.load_state_dict(synth.synthetic_state_dict(m.state_dict(), 0))
m = m.to(dev).eval()
feed = synth.synthetic_inputs(S, 12, 0)
nat = m.native()
if os.environ.get("DEC_BENCH_DBG"):  # timing experiment: e.g. 1 = no weight copies (garbage results)
    from slice3d_b200 import _native as _n
    _n.lib().s3d_debug_set_decoder_flags(int(os.environ["DEC_BENCH_DBG"]))
planes = nat.encode(feed["img_input"].to(dev))
ax = torch.linspace(-0.5, 0.5, nx).to(dev)
T = feed["trans_mat_wo_rot_tp"][0].to(dev)
n = nx ** 3
out = torch.empty(n, device=dev)
for prec in precs:
    for _ in range(2):
        nat.decode_grid(planes, 0, (ax, ax, ax), 0, n, T, precision=prec, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        nat.decode_grid(planes, 0, (ax, ax, ax), 0, n, T, precision=prec, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    from slice3d_b200 import _native
    _native.debug_profile(reset=True)
    nat.decode_grid(planes, 0, (ax, ax, ax), 0, n, T, precision=prec, out=out)
    pf = _native.debug_profile(reset=True)
    t = max(pf["tiles"], 1)
    print("   kcyc/tile:", " ".join(f"{k}={v / t / 1e3:.1f}" for k, v in pf.items() if k != "tiles"))
    tiles = (n + 8) // 9
    cyc_per_tile = ms * 1e-3 * 1.965e9 / (tiles / 148)
    print(f"{prec:7s} grid {nx}^3: {ms:8.2f} ms  {n / ms * 1e3:.3e} q/s  frac {n / ms * 1e3 * 32.82e6 / 1414.1e12:.3f}"
          f"  ~{cyc_per_tile / 1e3:.0f} kcyc/tile", flush=True)
