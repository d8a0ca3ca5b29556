"""ORACLE support -- training goldens from the UNMODIFIED reference (build container only).

    python oracle/make_golden_train.py grads      # tests/golden/train_grads_b2_s128.npz
    python oracle/make_golden_train.py traj       # tests/golden/train_traj_b4_s128.npz  (slow: ~30 min of CPU)

``grads``: the reference ``Slices3DRegModel`` in train mode (``model.train()``: batch-statistics BatchNorm; dropout
set to p = 0 on both sides so that the arithmetic is deterministic) on one seeded batch: the three
loss terms of ``cal_loss_pred`` (reg_slices/train.py:29-39), ``sdf_pred`` and the autograd gradients of a fixed set of
tensors spanning the whole graph (decoder head to the first trunk convolution).

``traj``: BASELINE configs[4] -- per-GPU batch 4, S = 128, n_qry = 256, Adam lr 3e-4 -- for world sizes 1, 2, 4, 8 as
DDP runs it (per-replica batch statistics, gradients averaged over ranks): three optimizer steps from the same seeded
weights; stored: the per-step loss terms averaged over ranks.  bench.py's train leg prints them next to its own.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle.make_golden import _template  # noqa: E402
from slice3d_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
GRAD_KEYS = ["fc_out.0.weight", "fc_out.0.bias", "fc_s.weight", "fc_s.bias", "fc_p.weight",
             "att_decoder.layers.2.linear2.weight", "att_decoder.layers.1.self_attn.in_proj_weight",
             "att_decoder.layers.0.norm1.weight", "att_decoder.layers.0.linear1.bias",
             "slices_generator.up4.conv.double_conv.0.weight", "slices_generator.up1.up.weight",
             "slices_generator.down1.0.weight", "slices_generator.emds.weight", "slices_generator.outc.conv.weight"]
MAX_STORE = 40000  # larger gradients are stored at a fixed strided subset


def loss_terms(ret, feed):
    """cal_loss_pred (reg_slices/train.py:29-39), pred_type 'sdf'."""
    return (F.l1_loss(ret["sdf_pred"], feed["sdf"]), F.l1_loss(ret["slices_rec"], feed["img_slices"]), ret["vgg_loss"])


def build(S, K, seed):
    torch.manual_seed(0)
    sd = synth.synthetic_state_dict(_template(S, K), seed)
    model = ref_shim.build_reference_model(S, "train", sd, K)
    model.train()
    synth.set_dropout(model, 0.0)
    return model


def grads_case(name="train_grads_b2_s128", S=128, K=12, B=2, seed=6):
    model = build(S, K, seed)
    feed = synth.synthetic_train_batch(S, K, batch=B, n_qry=256, seed=seed)
    ret = model({k: v.clone() for k, v in feed.items()})
    lp, li, lv = loss_terms(ret, feed)
    (lp + li + lv).backward()
    named = dict(model.named_parameters())
    out = {"img_size": S, "n_slices": K, "seed": seed, "batch": B, "loss": np.array([lp.item(), li.item(), lv.item()]),
           "sdf_pred": ret["sdf_pred"].detach().numpy()}
    for k in GRAD_KEYS:
        g = named[k].grad.reshape(-1)
        stride = max(1, g.numel() // MAX_STORE)
        out["grad:" + k] = g[::stride].numpy()
        out["stride:" + k] = stride
        out["norm:" + k] = float(named[k].grad.double().norm())
    no_grad = sorted(k for k, p in named.items() if p.requires_grad and p.grad is None)
    out["unused"] = np.array(no_grad)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, f"{os.path.getsize(path) / 1024:.0f} KiB", "loss", out["loss"], "unused", len(no_grad))


def traj_case(name="train_traj_b4_s128", S=128, K=12, B=4, seed=8, steps=3, worlds=(1, 2, 4, 8)):
    out = {"img_size": S, "n_slices": K, "seed": seed, "batch_per_rank": B, "steps": steps, "lr": 3e-4}
    for W in worlds:
        model = build(S, K, seed)
        opt = torch.optim.Adam(model.parameters(), lr=3e-4)  # reg_slices/train.py:136
        batches = [synth.synthetic_train_batch(S, K, batch=B, n_qry=256, seed=100 + r) for r in range(W)]
        traj = np.zeros((steps, 3))
        for s in range(steps):
            opt.zero_grad()
            for r in range(W):
                ret = model({k: v.clone() for k, v in batches[r].items()})
                lp, li, lv = loss_terms(ret, batches[r])
                ((lp + li + lv) / W).backward()  # DDP averages the ranks' gradients
                traj[s] += np.array([lp.item(), li.item(), lv.item()]) / W
            opt.step()
            print(f"world {W} step {s}: {traj[s]}", flush=True)
        out[f"loss_w{W}"] = traj
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "done")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    what = sys.argv[1] if len(sys.argv) > 1 else "grads"
    if what == "grads":
        grads_case()
    else:
        traj_case()
