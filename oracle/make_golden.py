"""ORACLE support -- generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py
Every case loads the reference ``Slices3DRegModel`` (oracle/ref_shim.py) with the
deterministic weights of ``slice3d_b200.synth.synthetic_state_dict`` and the
inputs of ``synthetic_inputs``; queries come from the reference's own
``make_3d_grid``.  Stored: the query indices / points, ``sdf_pred``, ``vgg_loss``
and strided sub-samples of the five feature planes and of ``slices_rec``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from slice3d_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name, img_size, n_slices, weight seed, [(grid nx, n sampled points)], mode
CASES = [
    ("cfg0_k4_s128_g64", 128, 4, 3, [(64, 1200)], "test"),
    ("k12_s128_g128", 128, 12, 1, [(128, 1500)], "test"),
    ("k12_s256_g128_g256", 256, 12, 2, [(128, 2000), (256, 2000)], "test"),
    ("k12_s128_val_rot", 128, 12, 4, [], "val"),
]

PLANE_STRIDES = [(8, 1), (8, 2), (8, 4), (8, 8), (4, 16)]  # (channel stride, pixel stride) per scale
REC_STRIDE = 8


def subsample_planes(feats):
    return [f[:, ::cs, ::ps, ::ps].contiguous().numpy() for f, (cs, ps) in zip(feats, PLANE_STRIDES)]


def run_case(name, S, K, seed, grids, mode):
    Model, make_3d_grid = ref_shim.import_reference()
    torch.manual_seed(0)
    tmpl = _template(S, K)
    sd = synth.synthetic_state_dict(tmpl, seed)
    model = ref_shim.build_reference_model(S, mode, sd, K)
    feed = synth.synthetic_inputs(S, K, seed)
    out = {"img_size": S, "n_slices": K, "seed": seed, "mode": mode}
    with torch.no_grad():
        feats, rec = model.slices_generator(feed["img_input"])
        for i, p in enumerate(subsample_planes(feats)):
            out[f"plane{i}"] = p
        out["slices_rec_sub"] = rec[:, :, ::REC_STRIDE, ::REC_STRIDE].contiguous().numpy()
        if mode == "test":
            for nx, n in grids:
                idx = synth.sample_grid_indices(nx, n, seed)
                pts = make_3d_grid((-0.5,) * 3, (0.5,) * 3, (nx,) * 3)[idx]
                f = dict(feed)
                f["qry_norot"] = pts.clone().unsqueeze(0)
                ret = model(f)
                out[f"idx_g{nx}"] = idx.numpy()
                out[f"pts_g{nx}"] = pts.numpy()
                out[f"sdf_g{nx}"] = ret["sdf_pred"][0].numpy()
                out["vgg_loss"] = ret["vgg_loss"].numpy()
                # the in-place flip of the caller's tensor (models.py:55)
                out[f"pts_after_g{nx}"] = f["qry_norot"][0].numpy()
        else:
            g = torch.Generator().manual_seed(77)
            B = 2
            feed = synth.synthetic_inputs(S, K, seed, batch=B)
            q = torch.rand(B, 256, 3, generator=g) - 0.5
            rot = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
            feed["obj_rot_mat"] = rot
            feed["qry_norot"] = q.clone()
            ret = model(feed)
            out["qry"] = q.numpy()
            out["obj_rot_mat"] = rot.numpy()
            out["sdf"] = ret["sdf_pred"].numpy()
            out["vgg_loss"] = ret["vgg_loss"].numpy()
            out["slices_rec_sub_b"] = ret["slices_rec"][:, :, ::REC_STRIDE, ::REC_STRIDE].contiguous().numpy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, f"{os.path.getsize(path) / 1024:.0f} KiB",
          "sdf range", [(float(out[k].min()), float(out[k].max())) for k in out if k.startswith("sdf")])


_TEMPLATES = {}


def _template(S, K):
    """name -> empty tensor of the right shape, taken from the reference module itself."""
    if (S, K) not in _TEMPLATES:
        Model, _ = ref_shim.import_reference()
        m = Model(img_size=S, n_slices=K, mode="test")
        if K != 12:
            m.slices_generator.n_slices = K
            m.slices_generator.emds = torch.nn.Embedding(K, 128)
        _TEMPLATES[(S, K)] = {k: torch.empty_like(v) for k, v in m.state_dict().items()}
    return _TEMPLATES[(S, K)]


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    for case in CASES:
        if only and case[0] not in only:
            continue
        run_case(*case)
