"""The oracle (oracle/oracle.py, a CPU restatement) against the golden vectors that
oracle/make_golden.py produced by running the UNMODIFIED reference module.  This is what
pins the oracle; the GPU parity tests then compare the CUDA path with the oracle/goldens."""
import pytest
import torch

from oracle import oracle
from tests import helpers

CASES = ["cfg0_k4_s128_g64", "k12_s128_g128", "k12_s256_g128_g256"]
TOL = 2e-5


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    case = helpers.load_case(name)
    _, sd = helpers.case_weights(case)
    feed = helpers.case_feed(case)
    K = int(case["n_slices"])
    torch.manual_seed(0)
    with torch.no_grad():
        feats, rec = oracle.unet_forward(sd, feed["img_input"], K)
        for i, f in enumerate(helpers.sub_planes(feats)):
            assert helpers.maxabs(f, case[f"plane{i}"]) < TOL, f"plane {i}"
        assert helpers.maxabs(rec[:, :, ::helpers.REC_STRIDE, ::helpers.REC_STRIDE], case["slices_rec_sub"]) < TOL
        for key in [k for k in case if k.startswith("pts_g")]:
            nx = key[len("pts_g"):]
            pts = torch.from_numpy(case[key]).unsqueeze(0)
            q = oracle.prepare_queries(pts, None, "test")
            assert helpers.maxabs(q[0], case[f"pts_after_g{nx}"]) == 0.0  # the in-place flip
            sdf = oracle.decode(sd, feats, q, feed["trans_mat_wo_rot_tp"], K)
            assert helpers.maxabs(sdf[0], case[f"sdf_g{nx}"]) < TOL, f"sdf g{nx}"


def test_oracle_grid_is_reference_grid():
    case = helpers.load_case("k12_s128_g128")
    grid = oracle.make_3d_grid((-0.5,) * 3, (0.5,) * 3, (128,) * 3)
    idx = torch.from_numpy(case["idx_g128"])
    assert torch.equal(grid[idx], torch.from_numpy(case["pts_g128"]))  # bit-exact coordinates


def test_oracle_val_mode_rotation_and_vgg():
    case = helpers.load_case("k12_s128_val_rot")
    _, sd = helpers.case_weights(case)
    feed = helpers.case_feed(case, batch=2)
    feed["qry_norot"] = torch.from_numpy(case["qry"])
    feed["obj_rot_mat"] = torch.from_numpy(case["obj_rot_mat"])
    with torch.no_grad():
        ret = oracle.model_forward(sd, feed, mode="val", n_slices=12, with_vgg=True)
    assert helpers.maxabs(ret["sdf_pred"], case["sdf"]) < TOL
    assert helpers.maxabs(ret["vgg_loss"], case["vgg_loss"]) < 1e-6
    assert helpers.maxabs(ret["slices_rec"][:, :, ::8, ::8], case["slices_rec_sub_b"]) < TOL


def test_timing_port_matches_oracle():
    """bench.py's CPU baseline (oracle/timing_port.py: the reference's library operators) computes
    the same values as the restatement."""
    from oracle.timing_port import TimingPort
    case = helpers.load_case("k12_s128_g128")
    _, sd = helpers.case_weights(case)
    feed = helpers.case_feed(case)
    with torch.no_grad():
        feats, _ = oracle.unet_forward(sd, feed["img_input"], 12)
        q = oracle.prepare_queries(torch.from_numpy(case["pts_g128"][:600]).unsqueeze(0), None, "test")
        got = TimingPort(sd).decode(feats, q, feed["trans_mat_wo_rot_tp"])
    assert helpers.maxabs(got[0], case["sdf_g128"][:600]) < TOL
