#!/usr/bin/env python
"""VGG19 perceptual loss at S=256, 12 slices: the CUDA library against the torch module (cuDNN, fp32)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slice3d_b200 import Slices3DRegModel, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = Slices3DRegModel(S, 12, "test")
m.load_state_dict(synth.synthetic_state_dict(m.state_dict(), 0))
m = m.to("cuda:0").eval()
a = (torch.rand(12, 3, S, S, device="cuda:0") * 2 - 1)
b = (torch.rand(12, 3, S, S, device="cuda:0") * 2 - 1)
nat = m.native()
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def t(f, n=5):
    for _ in range(2):
        f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        r = f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, float(r)


with torch.no_grad():
    ms_n, v_n = t(lambda: nat.vgg_loss(a, b))
    ms_t, v_t = t(lambda: m.vggptlossfunc(a, b)["pt_c_loss"])
print(f"VGG19 perceptual loss, 2 x 12 images {S}x{S}: library {ms_n:.2f} ms ({v_n:.6f}), torch/cuDNN fp32 {ms_t:.2f} ms ({v_t:.6f})")
