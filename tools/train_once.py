#!/usr/bin/env python
"""A few training steps of BASELINE configs[4]'s per-GPU shape (profiling target: ncu launch list of one train step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slice3d_b200 import Slices3DRegModel, synth, train_step  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
dev = "cuda:0"
m = Slices3DRegModel(128, 12, "train")
m.load_state_dict(synth.synthetic_state_dict(m.state_dict(), 8))
m = m.to(dev).train()
opt = torch.optim.Adam(m.parameters(), lr=3e-4)
batch = synth.synthetic_train_batch(128, 12, 4, 256, seed=100)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    print(train_step(dict(batch), m, opt))
torch.cuda.synchronize()
