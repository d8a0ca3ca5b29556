"""ORACLE support -- goldens of ``Slices3DGTModel`` from the UNMODIFIED reference (build container only).

    python oracle/make_golden_gt.py        # -> tests/golden/gt_k12_s128.npz

The reference module (reg_slices/src/model_gt.py) is imported through oracle/ref_shim.py, loaded with the seeded weights of
``slice3d_b200.synth.synthetic_state_dict`` and run in eval mode: 'test' mode on a sample of the 64^3 grid (in-place y,z
flip included) and 'val' mode with random rotations, batch 2.  S = 128 (the reference's unused classifier fixes it).
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
TAP_STRIDES = [(4, 16), (8, 8), (8, 4), (8, 2), (8, 1)]  # (channel stride, pixel stride) per tap conv1_2 .. conv5_3


def main(S=128, K=12, seed=9):
    ref_shim.import_reference()
    from src.model_gt import Slices3DGTModel
    from src_convonet.common import make_3d_grid
    torch.manual_seed(0)
    model = Slices3DGTModel(img_size=S, n_slices=K, mode="test")
    sd = synth.synthetic_state_dict({k: torch.empty_like(v) for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    feed = synth.synthetic_inputs(S, K, seed)
    out = {"img_size": S, "n_slices": K, "seed": seed}
    with torch.no_grad():
        feats, _ = model.img_encoder(feed["img_slices"].view(K, 3, S, S))
        for i, (f, (cs, ps)) in enumerate(zip(feats, TAP_STRIDES)):
            out[f"tap{i}"] = f[:, ::cs, ::ps, ::ps].contiguous().numpy()
        idx = synth.sample_grid_indices(64, 1500, seed)
        pts = make_3d_grid((-0.5,) * 3, (0.5,) * 3, (64,) * 3)[idx]
        f = dict(feed)
        f["qry_norot"] = pts.clone().unsqueeze(0)
        out["idx_g64"], out["pts_g64"] = idx.numpy(), pts.numpy()
        out["sdf_g64"] = model(f)["sdf_pred"][0].numpy()
        out["pts_after_g64"] = f["qry_norot"][0].numpy()
        model.mode = "val"
        g = torch.Generator().manual_seed(78)
        feed2 = synth.synthetic_inputs(S, K, seed, batch=2)
        feed2["qry_norot"] = torch.rand(2, 200, 3, generator=g) - 0.5
        feed2["obj_rot_mat"] = torch.linalg.qr(torch.randn(2, 3, 3, generator=g))[0]
        out["val_qry"], out["val_rot"] = feed2["qry_norot"].numpy(), feed2["obj_rot_mat"].numpy()
        out["val_sdf"] = model(feed2)["sdf_pred"].numpy()
    path = os.path.join(OUT, "gt_k12_s128.npz")
    np.savez_compressed(path, **out)
    print("->", path, f"{os.path.getsize(path) / 1024:.0f} KiB; sdf range", float(out["sdf_g64"].min()), float(out["sdf_g64"].max()))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    main()
