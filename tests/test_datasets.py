"""``slice3d_b200.datasets.Slice3DDataset`` (reference reg_slices/src/datasets.py:14-177) against
feed_dicts produced by the reference's OWN class, executed unmodified over the synthetic on-disk dataset of
tests/dataset_files.py (oracle/make_golden_dataset.py -> tests/golden/dataset_items.npz).  Everything compares EQUAL:
image tensors (Pillow's resize + torchvision's ToTensor / Normalize), camera matrices, query subsets, dtypes."""
import os

import numpy as np
import pytest
import torch

pytest.importorskip("PIL.Image")  # the files are written and decoded with Pillow

from slice3d_b200.datasets import Slice3DDataset  # noqa: E402
from tests import dataset_files, helpers  # noqa: E402

KEYS = {"img_input", "qry_norot", "obj_rot_mat", "trans_mat_wo_rot_tp", "occ", "sdf", "img_slices"}


@pytest.fixture(scope="module")
def root(tmp_path_factory):
    return dataset_files.write(str(tmp_path_factory.mktemp("s3d_dataset")))


def _decode(a):
    return torch.from_numpy(a).to(torch.float32).div(255).sub(0.5).div(0.5)


def _golden_item(g, tag, i):
    item = {k: g[f"{tag}:{i}:{k}"] for k in KEYS}
    for k in ("img_input", "img_slices"):
        item[k] = _decode(item[k]).numpy()
    return item


@pytest.mark.parametrize("tag", list(dataset_files.CONFIGS))
def test_getitem_equals_reference_class(root, tag):
    g = helpers.load_case("dataset_items")
    ds = Slice3DDataset("test", dataset_files.args(root, **dataset_files.CONFIGS[tag]))
    assert len(ds) == int(g[f"{tag}:len"]) == 2
    for i in range(len(ds)):
        item, want = ds[i], _golden_item(g, tag, i)
        assert set(item) == KEYS
        for k in KEYS:
            assert str(item[k].dtype) == str(g[f"{tag}:{i}:{k}:dtype"]), k
            assert item[k].shape == want[k].shape, (k, item[k].shape, want[k].shape)
            assert np.array_equal(item[k].numpy(), want[k]), (tag, i, k)
    assert ds[0]["img_input"].shape == (3, 32, 32) and ds[0]["img_slices"].shape == (36, 32, 32)


def test_batch_equals_collated_items_and_dataloader_works(root):
    from torch.utils.data import DataLoader
    for tag, cfg in dataset_files.CONFIGS.items():
        ds = Slice3DDataset("test", dataset_files.args(root, **cfg))
        want = next(iter(DataLoader(ds, batch_size=2, shuffle=False)))  # what train.py / reconstruct.py iterate over
        got = ds.batch([0, 1], "cpu")
        assert set(got) == KEYS
        for k in KEYS:
            assert got[k].dtype == want[k].dtype and torch.equal(got[k], want[k]), (tag, k)


def test_train_split_draws_views_and_query_subsets(root):
    ds = Slice3DDataset("train", dataset_files.args(root, n_qry=48))
    assert len(ds) == 3
    a = ds.batch([0, 1, 2], "cpu", rng=np.random.RandomState(1))
    assert a["qry_norot"].shape == (3, 48, 3) and a["sdf"].shape == (3, 48) and a["img_slices"].shape == (3, 36, 32, 32)
    assert torch.equal(a["occ"], (a["sdf"] <= 0).float())
    # the permutation of a train sample is not the fixed one of the test split, and the view is drawn from all n_views
    fixed = Slice3DDataset("val", dataset_files.args(root, n_qry=48))[0]["qry_norot"]
    assert not torch.equal(a["qry_norot"][0], fixed)
    views = {tuple(ds[0]["trans_mat_wo_rot_tp"].flatten().tolist()) + tuple(ds[0]["obj_rot_mat"].flatten().tolist()) for _ in range(24)}
    assert len(views) > 1


def test_non_rgba_input_is_rejected_like_the_reference(root):
    from PIL import Image
    ds = Slice3DDataset("test", dataset_files.args(root))
    p = os.path.join(root, "custom", "00_img_input", ds.files[0][1], "004.png")
    keep = open(p, "rb").read()
    try:
        Image.fromarray(np.zeros((45, 45, 3), dtype=np.uint8)).save(p)
        with pytest.raises(ValueError):
            ds[0]
    finally:
        open(p, "wb").write(keep)


def test_live_reference_class_on_the_val_split(root):
    """In the build container the reference's class itself is run beside ours (the val split is not in the golden)."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    from oracle.make_golden_dataset import reference_dataset_class
    Ref = reference_dataset_class()
    for cfg in dataset_files.CONFIGS.values():
        a = dataset_files.args(root, **cfg)
        ref, ours = Ref("val", a), Slice3DDataset("val", a)
        assert len(ref) == len(ours)
        for i in range(len(ours)):
            r, o = ref[i], ours[i]
            for k in KEYS:
                assert r[k].dtype == o[k].dtype and torch.equal(r[k], o[k]), k
