"""Shared helpers for the tests: load a golden case, rebuild its weights/inputs."""
import os

import numpy as np
import torch

from slice3d_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PLANE_STRIDES = [(8, 1), (8, 2), (8, 4), (8, 8), (4, 16)]
REC_STRIDE = 8


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def case_weights(case):
    """The deterministic state_dict the golden was generated with."""
    from slice3d_b200.models import Slices3DRegModel
    S, K, seed = int(case["img_size"]), int(case["n_slices"]), int(case["seed"])
    m = Slices3DRegModel(img_size=S, n_slices=K, mode=str(case["mode"]))
    sd = synth.synthetic_state_dict(m.state_dict(), seed)
    return m, sd


def case_feed(case, batch=1):
    S, K, seed = int(case["img_size"]), int(case["n_slices"]), int(case["seed"])
    return synth.synthetic_inputs(S, K, seed, batch=batch)


def sub_planes(feats):
    return [f[:, ::cs, ::ps, ::ps] for f, (cs, ps) in zip(feats, PLANE_STRIDES)]


def maxabs(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max())


def record(key, value):
    """Append a measured figure (max-abs errors of the GPU parity tests) to gpurun_out/parity_figures.json, so the
    numbers `pytest -q` hides end up in a file that travels back from the GPU box."""
    import json
    path = os.path.join(ROOT, "gpurun_out", "parity_figures.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        d = json.load(open(path)) if os.path.exists(path) else {}
        d[key] = float(value)
        json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass
