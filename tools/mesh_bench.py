#!/usr/bin/env python
"""generate_mesh end to end at the reference's default setting (resolution0 32, 3 upsampling steps -> 257^3):
encoder + MISE-refined value grid + marching cubes on the device, with the marching-cubes share timed separately."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slice3d_b200 import Generator3D, Slices3DRegModel, synth  # noqa: E402

S = 256
m = Slices3DRegModel(S, 12, "test")
m.load_state_dict(synth.synthetic_state_dict(m.state_dict(), 0))
m = m.to("cuda:0").eval()
feed = synth.synthetic_inputs(S, 12, 0)
gen = Generator3D(m, resolution0=32, upsampling_steps=3, pred_type="sdf")
with torch.no_grad():
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        grid = gen.generate_sparse_grid(feed, as_numpy=False)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        mesh = gen.extract_mesh(grid)
        torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"257^3: value grid (MISE) {1e3 * (t1 - t0):.1f} ms, marching cubes + transform + copy to host {1e3 * (t2 - t1):.1f} ms: "
      f"{len(mesh.vertices)} vertices, {len(mesh.faces)} faces")
