#!/usr/bin/env python
"""Test infrastructure: golden vectors for the MISE branch, generated with the REFERENCE's own MISE class
(compiled by oracle/build_ref_mise.py from /root/reference, unmodified).

    python oracle/make_golden_mise.py

* tests/golden/mise_<case>.npz: points-per-round and the dense (R+1)^3 volume for the synthetic fields of
  tests/mise_fields.py (incl. the reference's libmise/test.py known answer: 3 rounds, 5^3 dense, sum 105.0).
* tests/golden/sparse_k12_s128_r8_d2.npz: the reference's generate_from_latent MISE loop (reconstruct.py:147-167),
  restated around the reference MISE class with the CPU oracle as the model, on the weights / inputs of the
  k12_s128_g128 case: the 33^3 float64 value grid.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
from oracle import build_ref_mise, oracle  # noqa: E402
from tests import helpers, mise_fields  # noqa: E402

build_ref_mise.build()
from mise import MISE as RefMISE  # noqa: E402  (oracle/_ref: the compiled reference)

GOLD = os.path.join(ROOT, "tests", "golden")


def run(res0, depth, thr, f):
    m = RefMISE(res0, depth, thr)
    rounds = []
    p = m.query()
    while p.shape[0] != 0:
        rounds.append(p.shape[0])
        m.update(p, f(p))
        p = m.query()
    return np.array(rounds), m.to_dense()


for name, (res0, depth, thr, field) in mise_fields.CASES.items():
    res = res0 << depth
    rounds, dense = run(res0, depth, thr, lambda p: field(p, res))
    np.savez_compressed(os.path.join(GOLD, f"mise_{name}.npz"), rounds=rounds, dense=dense)
    print(name, rounds.tolist(), dense.shape, float(dense.sum()))

# ---- the reconstruct.py loop with the oracle as the model
case = helpers.load_case("k12_s128_g128")
_, sd = helpers.case_weights(case)
feed = helpers.case_feed(case)
torch.set_num_threads(os.cpu_count() or 1)
with torch.no_grad():
    feats, _ = oracle.unet_forward(sd, feed["img_input"], 12)

    def model_values(points, resolution, box_size=1.0):
        pointsf = points / resolution                      # reconstruct.py:152
        pointsf = box_size * (pointsf - 0.5)               # :154
        q = torch.FloatTensor(pointsf).unsqueeze(0)        # :157-158
        q = oracle.prepare_queries(q, None, "test")
        sdf = oracle.decode(sd, feats, q, feed["trans_mat_wo_rot_tp"], 12)
        return (-sdf).squeeze(0).numpy().astype(np.float64)  # :96,160-161

    res0, depth = 8, 2
    rounds, dense = run(res0, depth, 0.0, lambda p: model_values(p, res0 << depth))
np.savez_compressed(os.path.join(GOLD, "sparse_k12_s128_r8_d2.npz"), rounds=rounds, dense=dense, resolution0=res0,
                    depth=depth)
print("sparse", rounds.tolist(), dense.shape, float(np.abs(dense).max()))
