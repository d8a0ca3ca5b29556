#!/usr/bin/env python
"""MISE branch against the dense grid at the same final resolution (timing + points evaluated)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slice3d_b200 import Generator3D, Slices3DRegModel, synth  # noqa: E402

res0 = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
S = 256
m = Slices3DRegModel(S, 12, "test")
m.load_state_dict(synth.synthetic_state_dict(m.state_dict(), 0))
m = m.to("cuda:0").eval()
feed = synth.synthetic_inputs(S, 12, 0)
R = res0 << steps
gen = Generator3D(m, resolution0=res0, upsampling_steps=steps, pred_type="sdf")
with torch.no_grad():
    for it in range(2):
        stats = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        g = gen.generate_sparse_grid(feed, stats=stats, as_numpy=False)
        torch.cuda.synchronize(); t_sparse = time.perf_counter() - t0
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d = gen.generate_grid(feed, resolution=R + 1, as_numpy=False)
        torch.cuda.synchronize(); t_dense = time.perf_counter() - t0
n = stats["points_evaluated"]
print(f"MISE {res0} x 2^{steps} -> {R + 1}^3: {len(stats['points_per_round'])} rounds, {n} points evaluated "
      f"({100.0 * n / (R + 1) ** 3:.1f} % of the lattice), {t_sparse * 1e3:.1f} ms;  dense {R + 1}^3: {t_dense * 1e3:.1f} ms")
print("points per round:", stats["points_per_round"])
