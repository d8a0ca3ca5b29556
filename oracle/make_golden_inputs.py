"""ORACLE support -- goldens of the input pipeline from the reference's OWN functions (build container only).

The reference modules cannot be imported here (src/datasets.py and src/utils.py import trimesh / h5py / open3d, SURVEY.md
section 8c), so the functions on this path are taken from the reference files by name with ``ast`` and executed unmodified:
``png_2_whitebg`` / ``png_2_rgb`` (reg_slices/src/datasets.py:75-88) with the ``preprocess`` transform of :37, and
``getBlenderProj`` / ``get_rotate_matrix`` / ``get_W2O_mat`` (reg_slices/src/utils.py) chained as datasets.py:123-140 does.

    python oracle/make_golden_inputs.py      # -> tests/golden/inputs_*.npz
"""
import ast
import os
import sys

import numpy as np
import torch
import torchvision.transforms as T
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = "/root/reference/reg_slices/src"
OUT = os.path.join(ROOT, "tests", "golden")


def extract(path, names, ns):
    """exec the top-level functions / class methods called ``names`` of the file at ``path`` into ``ns``."""
    src = open(path).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = ast.get_source_segment(src, node)
            lines = code.split("\n")
            indent = len(lines[-1]) - len(lines[-1].lstrip()) if node.col_offset else 0
            code = "\n".join(l[node.col_offset:] if l[:node.col_offset].strip() == "" else l for l in lines)
            exec(code, ns)
    return ns


def reference_functions():
    ns = {"np": np, "Image": Image}
    extract(os.path.join(REF, "datasets.py"), {"png_2_whitebg", "png_2_rgb"}, ns)
    extract(os.path.join(REF, "utils.py"), {"getBlenderProj", "get_rotate_matrix", "get_W2O_mat"}, ns)
    return ns


def rgba_case(rng, h, w):
    rgba = rng.randint(0, 256, size=(3, h, w, 4)).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    rgba[0, ..., 3] = np.where((yy - h / 2) ** 2 + (xx - w / 2) ** 2 < (min(h, w) * 0.4) ** 2, 255, 0)  # hard mask
    rgba[1, ..., 3] = np.clip((xx * 300 // w) - 20, 0, 255)                                             # ramp with 0 and 255 runs
    return rgba                                                                                         # [2]: random alpha


def main():
    ns = reference_functions()
    rng = np.random.RandomState(7)
    out = {}
    for tag, (h, w, S) in {"77to64": (77, 77, 64), "64to64": (64, 64, 64), "137to128": (137, 137, 128),
                           "90x70to64": (90, 70, 64), "40to64": (40, 40, 64)}.items():
        rgba = rgba_case(rng, h, w)
        pre = T.Compose([T.Resize((S, S)), T.ToTensor(), T.Normalize(mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5])])  # datasets.py:37
        out[f"rgba_{tag}"] = rgba
        for name, fn in (("white", ns["png_2_whitebg"]), ("black", ns["png_2_rgb"])):
            imgs = [pre(fn(None, Image.fromarray(a, "RGBA"))) for a in rgba]
            out[f"{name}_{tag}"] = torch.stack(imgs).numpy()
        out[f"size_{tag}"] = np.array([h, w, S])
    cams = []
    for az, el, dist in [(0.0, 0.0, 1.2), (0.7, 0.3, 1.5), (-2.1, -0.4, 0.9), (3.0, 1.0, 2.0)]:
        K, RT = ns["getBlenderProj"](az, el, dist, img_w=1, img_h=1)
        W2O = ns["get_W2O_mat"]((0, 0, 0))
        rot_full = np.linalg.multi_dot([RT, ns["get_rotate_matrix"](-np.pi / 2)])
        obj_rot = np.transpose(rot_full)[:3, :]
        tmp = np.concatenate((np.eye(3), rot_full[:, 3:4]), axis=1)
        trans_tp = np.transpose(np.linalg.multi_dot([K, tmp, W2O]))
        cams.append(np.concatenate([[az, el, dist], torch.tensor(obj_rot).float().numpy().reshape(-1),
                                    torch.tensor(trans_tp).float().numpy().reshape(-1)]))
    out["cameras"] = np.stack(cams)
    path = os.path.join(OUT, "inputs_pipeline.npz")
    np.savez_compressed(path, **out)
    print("->", path, f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
