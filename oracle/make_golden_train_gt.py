"""ORACLE support -- training goldens of ``Slices3DGTModel`` from the UNMODIFIED reference (build container only).

    python oracle/make_golden_train_gt.py      # -> tests/golden/gt_train_grads_b2_s128.npz

The reference module (reg_slices/src/model_gt.py:12-111) in train mode (``model.train()``: batch-statistics BatchNorm in the
VGG16-BN trunk; dropout set to p = 0 on both sides so that the arithmetic is deterministic) on one seeded batch, with the
loss of reg_slices/train_gt.py:29-36 (L1 on ``sdf_pred``), the sign accuracy of train_gt.py:21-27, and the autograd
gradients of a fixed set of tensors spanning the graph from ``fc_out`` back to the first trunk convolution; then one Adam
step at train_gt.py's learning rate and the loss after it (what ``train_step`` returns on the second call).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from slice3d_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
GRAD_KEYS = ["fc_out.0.weight", "fc_out.0.bias", "fc_local.0.weight", "fc_local.2.bias", "pts_feat_extractor.0.weight",
             "pts_feat_extractor.4.weight", "att_decoder.layers.2.linear2.weight",
             "att_decoder.layers.1.self_attn.in_proj_weight", "att_decoder.layers.0.norm1.weight",
             "img_encoder.conv5_3.37.weight", "img_encoder.conv3_3.14.weight", "img_encoder.conv2_2.8.weight",
             "img_encoder.conv1_2.1.weight", "img_encoder.conv1_2.0.weight"]
MAX_STORE = 40000  # larger gradients are stored at a fixed strided subset


def main(name="gt_train_grads_b2_s128", S=128, K=12, B=2, seed=12, lr=3e-4):
    ref_shim.import_reference()
    from src.model_gt import Slices3DGTModel
    torch.manual_seed(0)
    model = Slices3DGTModel(img_size=S, n_slices=K, mode="train")
    sd = synth.synthetic_state_dict({k: torch.empty_like(v) for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd, strict=True)
    model.train()
    synth.set_dropout(model, 0.0)
    feed = synth.synthetic_train_batch(S, K, batch=B, n_qry=256, seed=seed)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    opt.zero_grad()
    ret = model({k: v.clone() for k, v in feed.items()})
    loss = F.l1_loss(ret["sdf_pred"], feed["sdf"])
    loss.backward()
    acc = ((ret["sdf_pred"] >= 0) == (feed["sdf"] >= 0)).float().sum(dim=-1) / ret["sdf_pred"].shape[1]
    named = dict(model.named_parameters())
    out = {"img_size": S, "n_slices": K, "seed": seed, "batch": B, "lr": lr, "loss": float(loss.item()),
           "acc": float(acc.mean(-1).item()), "sdf_pred": ret["sdf_pred"].detach().numpy()}
    for k in GRAD_KEYS:
        g = named[k].grad.reshape(-1)
        stride = max(1, g.numel() // MAX_STORE)
        out["grad:" + k] = g[::stride].numpy().copy()
        out["stride:" + k] = stride
        out["norm:" + k] = float(named[k].grad.double().norm())
    out["unused"] = np.array(sorted(k for k, p in named.items() if p.requires_grad and p.grad is None))
    opt.step()
    ret2 = model({k: v.clone() for k, v in feed.items()})
    out["loss_after_step"] = float(F.l1_loss(ret2["sdf_pred"], feed["sdf"]).item())
    # BatchNorm running statistics after the two train-mode forwards (momentum updates are part of the step)
    rm = model.state_dict()["img_encoder.conv1_2.1.running_mean"]
    out["running_mean_conv1_2_1"] = rm.numpy().copy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, f"{os.path.getsize(path) / 1024:.0f} KiB", "loss", out["loss"], "->", out["loss_after_step"],
          "acc", out["acc"], "unused", len(out["unused"]))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    main()
