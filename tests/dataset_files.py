"""A tiny synthetic dataset in the reference's on-disk layout (reg_slices/src/datasets.py:14-53, 89-145), written from
seeded arrays: PNG is lossless, so every reader decodes the same bytes.  Shared by tests/test_datasets.py and
oracle/make_golden_dataset.py (which runs the reference's own ``Slice3DDataset`` over these files)."""
import os
import pickle
from types import SimpleNamespace

import numpy as np
from PIL import Image

SHAPES = ["shape_a", "shape_b", "shape_c"]
N_VIEWS = 6
SLICE_STEMS = [f"{ax}_{i}" for ax in "XYZ" for i in (1, 2, 3, 4)]


def _rgba(rng, h, w):
    a = rng.randint(0, 256, size=(h, w, 4)).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    kind = rng.randint(3)
    if kind == 0:    # hard mask (what a renderer writes)
        a[..., 3] = np.where((yy - h / 2) ** 2 + (xx - w / 2) ** 2 < (min(h, w) * 0.4) ** 2, 255, 0)
    elif kind == 1:  # ramp with runs of 0 and 255
        a[..., 3] = np.clip((xx * 300 // w) - 20, 0, 255)
    return a         # kind 2: random alpha


def write(root, name_dataset="custom", img_hw=(45, 45), slice_hw=(40, 40), img_size=32, seed=3):
    """Write the files; returns the directory to pass as ``args.dir_data``."""
    rng = np.random.RandomState(seed)
    d = os.path.join(root, name_dataset)
    os.makedirs(os.path.join(d, "03_splits"), exist_ok=True)
    for split, ids in (("train", SHAPES), ("val", SHAPES[:2]), ("test", SHAPES[1:])):
        with open(os.path.join(d, "03_splits", f"{split}.lst"), "w") as f:
            f.write("\n".join(ids) + "\n")
    os.makedirs(os.path.join(d, "02_sdfs"), exist_ok=True)
    for s in SHAPES:
        os.makedirs(os.path.join(d, "00_img_input", s), exist_ok=True)
        for v in range(N_VIEWS):
            Image.fromarray(_rgba(rng, *img_hw)).save(os.path.join(d, "00_img_input", s, "%03d.png" % v))
            for sub, hw, mode in (("01_img_slices", slice_hw, "RGBA"), ("04_img_slices_gen", (img_size, img_size), "RGB")):
                os.makedirs(os.path.join(d, sub, s, "%03d" % v), exist_ok=True)
                for stem in SLICE_STEMS:
                    a = _rgba(rng, *hw)
                    Image.fromarray(a if mode == "RGBA" else a[..., :3].copy()).save(os.path.join(d, sub, s, "%03d" % v, stem + ".png"))
        az = (rng.rand(N_VIEWS) * 2 * np.pi).tolist()
        el = ((rng.rand(N_VIEWS) - 0.5) * 1.0).tolist()
        dist = (1.0 + rng.rand(N_VIEWS) * 0.5).tolist()
        meta = [None, az, el, dist, None, float(0.8 + 0.3 * rng.rand()), (rng.rand(3) * 0.1 - 0.05).tolist()]
        with open(os.path.join(d, "00_img_input", s, "meta.pkl"), "wb") as f:
            pickle.dump(meta, f)
        sdf = np.concatenate([rng.rand(700, 3) - 0.5, rng.randn(700, 1) * 0.05], 1).astype(np.float32)
        np.save(os.path.join(d, "02_sdfs", s + ".npy"), sdf)
    return root


def args(root, name_dataset="custom", img_size=32, n_qry=48, use_white_bg=False, from_which_slices="gt"):
    """The fields of options.py's parser that the dataset reads."""
    return SimpleNamespace(dir_data=root, name_dataset=name_dataset, img_size=img_size, n_qry=n_qry, n_views=N_VIEWS,
                           use_white_bg=use_white_bg, from_which_slices=from_which_slices,
                           categories_train="", categories_test="")


CONFIGS = {"gt_black": dict(use_white_bg=False, from_which_slices="gt"),
           "gt_white": dict(use_white_bg=True, from_which_slices="gt"),
           "gen_black": dict(use_white_bg=False, from_which_slices="gen")}
