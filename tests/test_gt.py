"""Slices3DGTModel (reference reg_slices/src/model_gt.py:12-111, src/vgg16bn_feats.py) against
goldens produced by the unmodified reference module (oracle/make_golden_gt.py)."""
import numpy as np
import pytest
import torch

from oracle import oracle
from slice3d_b200 import Generator3D, Slices3DGTModel, synth
from tests import helpers

TOL = 1e-4


def _case():
    g = helpers.load_case("gt_k12_s128")
    m = Slices3DGTModel(int(g["img_size"]), int(g["n_slices"]), "test")
    sd = synth.synthetic_state_dict(m.state_dict(), int(g["seed"]))
    m.load_state_dict(sd, strict=True)
    return g, m, sd


def test_gt_oracle_and_torch_path_match_reference_golden():
    g, m, sd = _case()
    feed = synth.synthetic_inputs(128, 12, int(g["seed"]))
    feed["qry_norot"] = torch.from_numpy(g["pts_g64"][:600]).unsqueeze(0)
    with torch.no_grad():
        want = oracle.gt_model_forward(sd, feed, "test", 12)
    assert helpers.maxabs(want["sdf_pred"][0], g["sdf_g64"][:600]) < 2e-5
    for i, (cs, ps) in enumerate([(4, 16), (8, 8), (8, 4), (8, 2), (8, 1)]):
        assert helpers.maxabs(want["feats"][i][:, ::cs, ::ps, ::ps], g[f"tap{i}"]) < 2e-5
    # the module's own torch path (used for training) in val mode, batch 2, rotations
    m.mode = "val"
    m.eval()
    feed2 = synth.synthetic_inputs(128, 12, int(g["seed"]), batch=2)
    feed2["qry_norot"] = torch.from_numpy(g["val_qry"])
    feed2["obj_rot_mat"] = torch.from_numpy(g["val_rot"])
    got = m._forward_autograd(feed2)["sdf_pred"].detach()
    assert helpers.maxabs(got, g["val_sdf"]) < 2e-5
    assert len(m.state_dict()) == 157  # the reference's checkpoint layout (checked key by key when the golden was made)


@pytest.mark.gpu
def test_gt_native_matches_reference_golden():
    g, m, sd = _case()
    dev = "cuda:0"
    m = m.to(dev).eval()
    feed = {k: v.to(dev) for k, v in synth.synthetic_inputs(128, 12, int(g["seed"])).items()}
    nat = m.native()
    planes, taps = nat.encode_gt(feed["img_slices"].view(12, 3, 128, 128), 1, want_taps=True)
    for i, (cs, ps) in enumerate([(4, 16), (8, 8), (8, 4), (8, 2), (8, 1)]):
        assert helpers.maxabs(taps[i][:, ::cs, ::ps, ::ps].cpu(), g[f"tap{i}"]) < TOL, f"tap {i}"
    for prec in ("fp32", "fp16f8", "fp16x3", "bf16x3"):
        m.precision = prec
        feed["qry_norot"] = torch.from_numpy(g["pts_g64"]).unsqueeze(0).to(dev)
        with torch.no_grad():
            ret = m(feed)
        err = helpers.maxabs(ret["sdf_pred"][0].cpu(), g["sdf_g64"])
        print(f"GT model, {prec}: max-abs vs reference {err:.3e}")
        helpers.record(f"gt_model_{prec}_max_abs_vs_reference", err)
        assert err < TOL
        assert torch.equal(feed["qry_norot"][0].cpu(), torch.from_numpy(g["pts_after_g64"]))  # in-place flip (model_gt.py:75)
    # val mode: batch 2, rotations, one launch
    m.mode, m.precision = "val", "fp16f8"
    feed2 = {k: v.to(dev) for k, v in synth.synthetic_inputs(128, 12, int(g["seed"]), batch=2).items()}
    feed2["qry_norot"] = torch.from_numpy(g["val_qry"]).to(dev)
    feed2["obj_rot_mat"] = torch.from_numpy(g["val_rot"]).to(dev)
    with torch.no_grad():
        got = m(feed2)["sdf_pred"]
    assert helpers.maxabs(got.cpu(), g["val_sdf"]) < TOL
    # the reference's extraction driver on top: dense 24^3 grid through Generator3D == explicit points
    m.mode = "test"
    gen = Generator3D(m, upsampling_steps=0, resolution0=24, pred_type="sdf")
    with torch.no_grad():
        vol = gen.generate_grid({k: v.cpu() for k, v in feed.items() if k != "qry_norot"})
        pts = synth.make_3d_grid((-0.5,) * 3, (0.5,) * 3, (24,) * 3)
        want = oracle.gt_model_forward(sd, {**synth.synthetic_inputs(128, 12, int(g["seed"])), "qry_norot": pts.unsqueeze(0)},
                                       "test", 12)["sdf_pred"][0]
    assert helpers.maxabs(vol.reshape(-1), -want) < TOL
    # more queries than one token pass (32768) and a ragged tail
    big = (torch.rand(1, 40001, 3, generator=torch.Generator().manual_seed(1)) - 0.5).to(dev)
    with torch.no_grad():
        a = m({**feed, "qry_norot": big.clone()})["sdf_pred"][0]
        b = m({**feed, "qry_norot": big[:, 32700:32900].clone().contiguous()})["sdf_pred"][0]
    assert helpers.maxabs(a[32700:32900].cpu(), b.cpu()) < 1e-6


def test_gt_train_step_cpu_matches_reference_gradients():
    """Training of Slices3DGTModel (reference reg_slices/train_gt.py:21-50) against a golden produced by the UNMODIFIED
    reference module in train mode (oracle/make_golden_train_gt.py; dropout p = 0 on both sides): the loss, the sign
    accuracy, sdf_pred, the autograd gradients of tensors spanning the graph from fc_out back to the first trunk
    convolution, the set of parameters that never get a gradient, and -- through ``train_step_gt`` itself -- the loss
    after one Adam step and the BatchNorm running statistics the two train-mode forwards leave behind."""
    from slice3d_b200 import train_step_gt, val_step_gt
    case = helpers.load_case("gt_train_grads_b2_s128")
    S, K, seed, B = int(case["img_size"]), int(case["n_slices"]), int(case["seed"]), int(case["batch"])
    torch.manual_seed(0)
    m = Slices3DGTModel(S, K, "train")
    m.load_state_dict(synth.synthetic_state_dict(m.state_dict(), seed), strict=True)
    m = synth.set_dropout(m.train(), 0.0)
    feed = synth.synthetic_train_batch(S, K, batch=B, n_qry=256, seed=seed)

    opt = torch.optim.Adam(m.parameters(), lr=float(case["lr"]))  # train_gt.py:115
    # (train_step zeroes the gradients BEFORE the forward, so they can be read after the step)
    loss, acc = train_step_gt({k: v.clone() for k, v in feed.items()}, m, opt)
    assert abs(loss - float(case["loss"])) <= 2e-5 * abs(float(case["loss"])) + 1e-6, (loss, float(case["loss"]))
    assert abs(acc - float(case["acc"])) < 1e-6, (acc, float(case["acc"]))
    named = dict(m.named_parameters())
    rels = {}
    for key in [k[5:] for k in case if k.startswith("grad:")]:
        g = named[key].grad.detach().reshape(-1)[::int(case["stride:" + key])].double().numpy()
        ref = case["grad:" + key].astype(np.float64)
        rels[key] = float(np.linalg.norm(g - ref) / max(np.linalg.norm(ref), 1e-30))
    print("GT train gradients, |g - ref| / |ref|: " + ", ".join(f"{k} {v:.1e}" for k, v in rels.items()))
    assert max(rels.values()) < 2e-4, rels
    unused = sorted(k for k, p in named.items() if p.requires_grad and p.grad is None)
    assert unused == sorted(str(x) for x in case["unused"])
    # second call: the loss after the first Adam step (what the reference's loop prints next)
    loss2, _ = train_step_gt({k: v.clone() for k, v in feed.items()}, m, opt)
    ref2 = float(case["loss_after_step"])
    assert abs(loss2 - ref2) <= 1e-3 * abs(ref2), (loss2, ref2)
    rm = m.state_dict()["img_encoder.conv1_2.1.running_mean"].numpy()
    assert np.allclose(rm, case["running_mean_conv1_2_1"], rtol=1e-4, atol=1e-6)
    # val_step_gt over a one-batch "loader": the same loss function, eval-mode arithmetic on grad-free tensors is the
    # native path (CUDA only), so the CPU check runs the loader through the train-mode module
    lv, av = val_step_gt(m, [{k: v.clone() for k, v in feed.items()}])
    assert np.isfinite(lv) and 0.0 <= av <= 1.0
