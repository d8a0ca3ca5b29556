"""Input pipeline (reference reg_slices/src/datasets.py:37,75-177, src/utils.py:29-73,132-170)
against goldens produced by the reference's own functions (oracle/make_golden_inputs.py) and, where the libraries are
present, against PIL + torchvision directly.  Bar: bit-exact (8-bit image arithmetic; the float32 tensors compare equal)."""
import numpy as np
import pytest
import torch

from slice3d_b200 import inputs, synth
from tests import helpers

TAGS = ["77to64", "64to64", "137to128", "90x70to64", "40to64"]


@pytest.mark.parametrize("tag", TAGS)
def test_host_tables_reproduce_reference_preprocessing(tag):
    g = helpers.load_case("inputs_pipeline")
    S = int(g[f"size_{tag}"][2])
    for name, white in (("white", True), ("black", False)):
        got = inputs.preprocess_rgba_host(g[f"rgba_{tag}"], S, white).numpy()
        assert got.dtype == np.float32 and np.array_equal(got, g[f"{name}_{tag}"]), (tag, name)


def test_resample_against_pillow_directly():
    Image = pytest.importorskip("PIL.Image")
    T = pytest.importorskip("torchvision.transforms")
    rng = np.random.RandomState(3)
    for h, w, S in [(50, 50, 32), (33, 91, 48), (200, 150, 128), (16, 16, 64)]:
        rgb = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        rgba = np.concatenate([rgb, np.full((h, w, 1), 255, np.uint8)], -1)[None]
        want = T.Compose([T.Resize((S, S)), T.ToTensor(), T.Normalize([0.5] * 3, [0.5] * 3)])(Image.fromarray(rgb)).numpy()
        got = inputs.preprocess_rgba_host(rgba, S, False).numpy()[0]
        assert np.array_equal(got, want), (h, w, S)


def test_camera_matrices_match_reference_and_known_answer():
    g = helpers.load_case("inputs_pipeline")
    for row in g["cameras"]:
        rot, T = inputs.camera_matrices(float(row[0]), float(row[1]), float(row[2]))
        assert np.array_equal(rot.numpy().reshape(-1), row[3:12].astype(np.float32))
        assert np.array_equal(T.numpy().reshape(-1), row[12:24].astype(np.float32))
    # the camera of create_dataset_sin_img.py (az = el = 0, distance 1.2): SURVEY.md section 8(d)
    rot, T = inputs.camera_matrices(0.0, 0.0, 1.2)
    assert torch.equal(T, torch.tensor(synth.CAMERA_T))
    assert helpers.maxabs(rot, torch.tensor(synth.OBJ_ROT)) < 1e-7


def test_query_preparation_and_sample_layout():
    rng = np.random.RandomState(0)
    sdf_npy = np.concatenate([rng.rand(500, 3) - 0.5, rng.randn(500, 1) * 0.05], 1)
    scale, offset = 0.9, (0.01, -0.02, 0.03)
    q, occ, sdf = inputs.prepare_queries(sdf_npy, scale, offset, 64, split="val")
    perm = np.random.RandomState(1234).permutation(500)[:64]  # datasets.py:161-165
    want_pt = sdf_npy[:, :3] * scale + np.array([offset[0], offset[2], -offset[1]])
    want_sdf = (sdf_npy[:, 3] - 0.003) * scale
    assert np.array_equal(q.numpy(), want_pt[perm].astype(np.float32))
    assert np.array_equal(sdf.numpy(), want_sdf[perm].astype(np.float32))
    assert np.array_equal(occ.numpy(), (want_sdf[perm] <= 0).astype(np.float32))
    q2, _, _ = inputs.prepare_queries(sdf_npy, scale, offset, 64, split="train", rng=np.random.RandomState(5))
    assert q2.shape == (64, 3) and not np.array_equal(q2.numpy(), q.numpy())
    assert inputs.SLICE_ORDER == ["X_1", "X_2", "X_3", "X_4", "Z_4", "Z_3", "Z_2", "Z_1", "Y_1", "Y_2", "Y_3", "Y_4"]
    imgs = torch.arange(13 * 3 * 4 * 4, dtype=torch.float32).view(13, 3, 4, 4)
    feed = inputs.assemble_sample(imgs, 0.0, 0.0, 1.2, q, occ, sdf)
    assert feed["img_input"].shape == (3, 4, 4) and feed["img_slices"].shape == (36, 4, 4)
    assert torch.equal(feed["img_slices"][3:6], imgs[2])  # slice k occupies channels 3k .. 3k+2 (models.py:86)
    assert set(feed) == {"img_input", "qry_norot", "obj_rot_mat", "trans_mat_wo_rot_tp", "occ", "sdf", "img_slices"}


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_device_preprocessing_is_bit_exact(tag):
    g = helpers.load_case("inputs_pipeline")
    S = int(g[f"size_{tag}"][2])
    rgba = torch.from_numpy(g[f"rgba_{tag}"]).to("cuda:0")
    for name, white in (("white", True), ("black", False)):
        got = inputs.preprocess_rgba(rgba, S, white)
        assert got.shape == (3, 3, S, S) and got.dtype == torch.float32
        assert np.array_equal(got.cpu().numpy(), g[f"{name}_{tag}"]), (tag, name)


@pytest.mark.gpu
def test_device_preprocessing_batch_of_samples():
    """A training batch worth of images in one call (4 samples x 13 PNGs, 137 -> 128), against the host tables."""
    rng = np.random.RandomState(11)
    rgba = rng.randint(0, 256, size=(52, 137, 137, 4)).astype(np.uint8)
    rgba[::2, :, :40, 3] = 0
    got = inputs.preprocess_rgba(torch.from_numpy(rgba).to("cuda:0"), 128, True).cpu()
    want = inputs.preprocess_rgba_host(rgba, 128, True)
    assert torch.equal(got, want)
    with pytest.raises(Exception):
        inputs.preprocess_rgba(torch.from_numpy(rgba), 128, True)  # host tensor: no silent CPU path


def test_factored_projection_equals_full_camera_projection():
    """The reference's own check (reg_slices/test_projection.py:8-20, 99-112, printed side by side there): projecting a point
    with the full matrix K.RT.rot.W2O equals rotating it by ``obj_rot_mat`` and applying ``trans_mat_wo_rot_tp`` -- the
    factorisation Slices3DRegModel.forward relies on (models.py:57-60, 28-36)."""
    rng = np.random.RandomState(11)
    for _ in range(20):
        az, el, dist = rng.rand() * 2 * np.pi, (rng.rand() - 0.5) * 1.2, 1.0 + rng.rand()
        rot, T = inputs.camera_matrices(az, el, dist)
        K, RT = inputs.blender_proj(az, el, dist, img_w=1, img_h=1)
        full = np.transpose(np.linalg.multi_dot([K, RT, inputs.rotate_matrix(-np.pi / 2), np.eye(4)]))  # (4,3)
        pts = rng.rand(50, 3) - 0.5
        a = np.concatenate([pts, np.ones((50, 1))], 1) @ full
        b = np.concatenate([pts @ rot.double().numpy(), np.ones((50, 1))], 1) @ T.double().numpy()
        assert np.allclose(a, b, rtol=0, atol=2e-6), np.abs(a - b).max()
        assert np.all(a[:, 2] > 0.3)  # in front of the camera: the perspective divide of project_coord is safe
