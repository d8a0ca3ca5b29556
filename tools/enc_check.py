#!/usr/bin/env python
"""Encoder check: tensor-core encoder vs the fp32 CUDA-core encoder (s3d_debug_set_encoder) vs the golden planes,
plus timing of both.  Kernel-development tool, not a test."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slice3d_b200 import _native, synth  # noqa: E402
from tests import helpers  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "k12_s128_g128"
case = helpers.load_case(name)
dev = "cuda:0"


def run(simt):
    m, sd = helpers.case_weights(case)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    feed = {k: v.to(dev) for k, v in helpers.case_feed(case, 1).items()}
    nat = m.native()
    _native.lib().s3d_debug_set_encoder(nat._h, 1 if simt else 0)
    planes, feats = nat.encode(feed["img_input"], want_feats=True)
    torch.cuda.synchronize()
    for _ in range(2):
        nat.encode(feed["img_input"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        nat.encode(feed["img_input"])
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    return [f.cpu() for f in feats], planes.slices_rec.cpu(), planes.blob.cpu(), ms


ft, rt, bt, ms_t = run(False)
fs, rs, bs, ms_s = run(True)
print(f"{name}: tensor-core encoder {ms_t:.2f} ms, fp32 CUDA-core encoder {ms_s:.2f} ms")
sub_t, sub_s = helpers.sub_planes(ft), helpers.sub_planes(fs)
for i in range(5):
    g = torch.from_numpy(case[f"plane{i}"])
    print(f"plane{i}: |golden| max {float(g.abs().max()):.3f}  tc-golden {helpers.maxabs(sub_t[i], g):.3e}  "
          f"simt-golden {helpers.maxabs(sub_s[i], g):.3e}  tc-simt(full) {float((ft[i] - fs[i]).abs().max()):.3e}")
print(f"slices_rec tc-simt {float((rt - rs).abs().max()):.3e}; projected planes tc-simt {float((bt - bs).abs().max()):.3e}")
