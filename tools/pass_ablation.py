"""Per-GEMM pass ablation of the tensor-core decoder's split-operand scheme (CPU emulation).

For each of the four contractions of a transformer layer {qkv, out, ffn1, ffn2} the operands are
split x = hi + lo in bf16 or fp16 and the products that a reduced-pass tcgen05 schedule would issue
are summed in fp32 (exact products, fp32 accumulate = what the tensor core does up to summation
order).  Schemes:  "3" = hi.hi + lo.hi + hi.lo;  "2x" = hi.hi + lo.hi (weight lo dropped);
"2w" = hi.hi + hi.lo (activation lo dropped);  "1" = hi.hi.   Everything not ablated runs as "3" in
the same split type, so each row of the table isolates ONE contraction.  Reports max-abs error of
sdf_pred against the fp32 oracle on a sample of the 256^3 grid (K=12, S=256: BASELINE configs[2]).

    python tools/pass_ablation.py [n_queries]
"""
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402  (tool, not product)
from slice3d_b200 import Slices3DRegModel, synth  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def split(x, dt):
    hi = x.to(dt).float()
    lo = (x - hi).to(dt).float()
    return hi, lo


def mm(x, w, scheme, dt):
    """x (.., K) @ w (N, K)^T under a pass scheme."""
    if scheme == "fp32":
        return x @ w.t()
    xh, xl = split(x, dt)
    wh, wl = split(w, dt)
    y = xh @ wh.t()
    if scheme in ("3", "2x"):
        y = y + xl @ wh.t()
    if scheme in ("3", "2w"):
        y = y + xh @ wl.t()
    return y


def layer(sd, p, x, sch, dt):
    N, L, D = x.shape
    hd = 32
    qkv = mm(x, sd[p + ".self_attn.in_proj_weight"], sch["qkv"], dt) + sd[p + ".self_attn.in_proj_bias"]
    q, k, v = qkv.split(D, dim=-1)
    sh = lambda t: t.view(N, L, 4, hd).transpose(1, 2)
    q, k, v = sh(q), sh(k), sh(v)
    att = torch.softmax((q @ k.transpose(-1, -2)) / hd ** 0.5, dim=-1)
    o = (att @ v).transpose(1, 2).reshape(N, L, D)
    o = mm(o, sd[p + ".self_attn.out_proj.weight"], sch["out"], dt) + sd[p + ".self_attn.out_proj.bias"]
    x = F.layer_norm(x + o, (D,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5)
    h = F.relu(mm(x, sd[p + ".linear1.weight"], sch["ffn1"], dt) + sd[p + ".linear1.bias"])
    h = mm(h, sd[p + ".linear2.weight"], sch["ffn2"], dt) + sd[p + ".linear2.bias"]
    return F.layer_norm(x + h, (D,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-5)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    S, K, nx, seed = 256, 12, 256, 2
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    sd = synth.synthetic_state_dict(Slices3DRegModel(S, K, "test").state_dict(), seed)
    feed = synth.synthetic_inputs(S, K, seed)
    idx = synth.sample_grid_indices(nx, n, seed)
    pts = synth.make_3d_grid((-0.5,) * 3, (0.5,) * 3, (nx,) * 3)[idx]
    with torch.no_grad():
        feats, _ = oracle.unet_forward(sd, feed["img_input"], K)
        q = oracle.prepare_queries(pts.unsqueeze(0), None, "test")
        ref, toks = oracle.decode(sd, feats, q, feed["trans_mat_wo_rot_tp"], K, return_tokens=True)
        x0 = toks[0]
        rows = []

        def run(sch, dt):
            x = x0
            for l in range(3):
                x = layer(sd, f"att_decoder.layers.{l}", x, sch, dt)
            out = (x[:, 0, :] @ sd["fc_out.0.weight"].t() + sd["fc_out.0.bias"]).squeeze(-1)
            return float((out - ref[0]).abs().max())

        base = {g: "3" for g in ("qkv", "out", "ffn1", "ffn2")}
        for dt, name in ((torch.bfloat16, "bf16"), (torch.float16, "fp16")):
            rows.append({"split": name, "ablated": "none", "scheme": "3", "max_abs": run(base, dt)})
            for g, s in itertools.product(("qkv", "out", "ffn1", "ffn2"), ("2x", "2w", "1")):
                sch = dict(base)
                sch[g] = s
                rows.append({"split": name, "ablated": g, "scheme": s, "max_abs": run(sch, dt)})
            for s in ("2x", "2w"):
                rows.append({"split": name, "ablated": "all", "scheme": s,
                             "max_abs": run({g: s for g in base}, dt)})
        print(json.dumps({"n_queries": int(idx.numel()), "sdf_abs_max": float(ref.abs().max()), "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
