#!/usr/bin/env python
"""The bench's `sparse.smooth` case in isolation: per-iteration wall time of generate_sparse_grid (kernel-development tool)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slice3d_b200 import Generator3D, Slices3DRegModel, synth  # noqa: E402

S, K, dev = 256, 12, "cuda:0"
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16f8"
for name in ("noisy", "smooth"):
    m = Slices3DRegModel(S, K, "test", precision=prec)
    sd = synth.synthetic_state_dict(m.state_dict(), 0)
    if name == "smooth":
        for k in sd:
            if "att_decoder" in k and ("out_proj" in k or "linear2" in k):
                sd[k] = torch.zeros_like(sd[k])
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    feed = synth.synthetic_inputs(S, K, 0)
    gen = Generator3D(m, resolution0=32, upsampling_steps=3, pred_type="sdf")
    ts = []
    for _ in range(5):
        stats = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        grid = gen.generate_sparse_grid(feed, stats=stats, as_numpy=False)
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    print(name, prec, "ms per call:", " ".join(f"{t:.1f}" for t in ts), "rounds", stats["points_per_round"][:8], flush=True)
