#!/usr/bin/env python
"""Two plane-encoder calls at S=256, K=12 (run under ncu for a per-kernel launch list)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slice3d_b200 import Slices3DRegModel, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = Slices3DRegModel(S, 12, "test")
m.load_state_dict(synth.synthetic_state_dict(m.state_dict(), 0))
m = m.to("cuda:0").eval()
img = synth.synthetic_inputs(S, 12, 0)["img_input"].to("cuda:0")
nat = m.native()
for _ in range(2):
    nat.encode(img, want_slices_rec=True)
torch.cuda.synchronize()
if os.environ.get("ENC_PROF_TIME"):  # wall clock vs CUDA events over 20 calls (is the encoder launch-bound on the host?)
    import time
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(20):
        nat.encode(img, want_slices_rec=False)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"S={S}: host enqueue {(t1 - t0) / 20 * 1e3:.3f} ms / call, events {e0.elapsed_time(e1) / 20:.3f} ms / call, "
          f"wall {(t2 - t0) / 20 * 1e3:.3f} ms / call")
