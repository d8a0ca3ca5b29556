"""ORACLE support -- goldens of ``Slice3DDataset.__getitem__`` from the reference's OWN class (build container only).

    python oracle/make_golden_dataset.py      # -> tests/golden/dataset_items.npz

reg_slices/src/datasets.py cannot be imported here (it imports trimesh / h5py), so the class
``Slice3DDataset`` and the three camera functions it calls are taken from the reference files with ``ast`` and executed
UNMODIFIED in a namespace that supplies the modules they use.  The class then reads the synthetic on-disk dataset of
tests/dataset_files.py (test split: view 4, numpy legacy seed 1234 -- deterministic) and its feed_dicts are stored:
images as the uint8 code c with value ((c / 255) - 0.5) / 0.5 (exact: that is how ToTensor + Normalize made them).
"""
import ast
import os
import pickle
import random
import sys
import tempfile

import numpy as np
import torch
import torchvision.transforms as T
from PIL import Image
from torch.utils.data import Dataset

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = "/root/reference/reg_slices/src"
OUT = os.path.join(ROOT, "tests", "golden")

from tests import dataset_files  # noqa: E402


def reference_dataset_class():
    ns = {"np": np, "torch": torch, "Dataset": Dataset, "os": os, "Image": Image, "T": T, "random": random, "pickle": pickle}
    src = open(os.path.join(REF, "utils.py")).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in {"getBlenderProj", "get_rotate_matrix", "get_W2O_mat"}:
            exec(ast.get_source_segment(src, node), ns)
    src = open(os.path.join(REF, "datasets.py")).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == "Slice3DDataset":
            exec(ast.get_source_segment(src, node), ns)
    return ns["Slice3DDataset"]


def encode_images(t):
    """float tensor made by ToTensor + Normalize(0.5, 0.5) -> its uint8 code (checked to be exact)."""
    code = torch.round((t * 0.5 + 0.5) * 255).to(torch.uint8)
    back = code.to(torch.float32).div(255).sub(0.5).div(0.5)
    assert torch.equal(back, t), "image tensor is not an exact function of an 8-bit code"
    return code.numpy()


def main():
    Ref = reference_dataset_class()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        root = dataset_files.write(tmp)
        for tag, cfg in dataset_files.CONFIGS.items():
            ds = Ref("test", dataset_files.args(root, **cfg))
            out[f"{tag}:len"] = len(ds)
            for i in range(len(ds)):
                item = ds[i]
                for k, v in item.items():
                    out[f"{tag}:{i}:{k}"] = encode_images(v) if k in ("img_input", "img_slices") else v.numpy()
                    out[f"{tag}:{i}:{k}:dtype"] = str(v.dtype)
    path = os.path.join(OUT, "dataset_items.npz")
    np.savez_compressed(path, **out)
    print("->", path, f"{os.path.getsize(path) / 1024:.0f} KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
