#!/usr/bin/env python
"""Test infrastructure: golden vectors for marching cubes, generated with the REFERENCE's own marching cubes core
(compiled by oracle/build_ref_mcubes.py from /root/reference, unmodified) through the call sequence of
Generator3D.extract_mesh (reconstruct.py:175-223).

    python oracle/make_golden_mcubes.py    # -> tests/golden/mcubes_<case>.npz (volume, isovalue, vertices, triangles)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref_mcubes  # noqa: E402
from tests import mc_volumes  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
for name, (vol, iso) in mc_volumes.cases().items():
    v, t = build_ref_mcubes.run(vol, iso)
    np.savez_compressed(os.path.join(GOLD, f"mcubes_{name}.npz"), isovalue=iso, vertices=v, triangles=t.astype(np.int64))
    print(name, vol.shape, v.shape, t.shape)

# extract_mesh: the reference's padding and vertex post-transform (reconstruct.py:186-207), box_size 1
vol, iso = mc_volumes.cases()["blob_17"]
padded = np.pad(vol, 1, "constant", constant_values=-1e6)
v, t = build_ref_mcubes.run(padded, iso)
v = v - 0.5
v = v - 1
v = v / np.array([vol.shape[0] - 1, vol.shape[1] - 1, vol.shape[2] - 1])
v = 1.0 * (v - 0.5)
np.savez_compressed(os.path.join(GOLD, "mcubes_extract_blob_17.npz"), vertices=v, triangles=t.astype(np.int64))
print("extract_mesh", v.shape, t.shape)
