"""ORACLE -- test infrastructure, not product code.

CPU restatement (plain torch fp32 tensor ops on the host) of the reference's
slice-to-3D path, written from a flat ``state_dict`` with no nn.Module from
either the reference or ``slice3d_b200``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this file; the product (``slice3d_b200``) never does.

Parity status: PINNED against the reference itself.  ``oracle/make_golden.py``
imports the unmodified reference module from /root/reference (three import
shims), runs it on the seeded weights/inputs of
``slice3d_b200.synth`` and stores its outputs under ``tests/golden``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors
(<= 2e-5 max-abs).  The reference ships no golden vectors of its own for this
path.

Every function cites the reference lines it restates (paths relative to
/root/reference/reg_slices).
"""
import math

import torch
import torch.nn.functional as F

EPS_BN = 1e-5
EPS_LN = 1e-5

# torchvision vgg16_bn "features" indices of the conv layers inside each block and
# whether BN+ReLU follows inside the same block (src/unet_custom.py:15-20).
_DOWN_BLOCKS = {
    "down1": [("conv", 0), ("bn", 1), ("relu",), ("conv", 3)],
    "down2": [("bn", 4), ("relu",), ("pool",), ("conv", 7), ("bn", 8), ("relu",), ("conv", 10)],
    "down3": [("bn", 11), ("relu",), ("pool",), ("conv", 14), ("bn", 15), ("relu",), ("conv", 17), ("bn", 18),
              ("relu",), ("conv", 20)],
    "down4": [("bn", 21), ("relu",), ("pool",), ("conv", 24), ("bn", 25), ("relu",), ("conv", 27), ("bn", 28),
              ("relu",), ("conv", 30)],
    "down5": [("bn", 31), ("relu",), ("pool",), ("conv", 34), ("bn", 35), ("relu",), ("conv", 37), ("bn", 38),
              ("relu",), ("conv", 40)],
}


def _bn(sd, prefix, x):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=EPS_BN)


def _down(sd, name, x, prefix=None):
    p = prefix if prefix is not None else "slices_generator." + name + "."
    for op in _DOWN_BLOCKS[name]:
        if op[0] == "conv":
            x = F.conv2d(x, sd[f"{p}{op[1]}.weight"], sd[f"{p}{op[1]}.bias"], padding=1)
        elif op[0] == "bn":
            x = _bn(sd, f"{p}{op[1]}", x)
        elif op[0] == "relu":
            x = F.relu(x)
        else:
            x = F.max_pool2d(x, 2, 2)
    return x


def _double_conv(sd, p, x):
    """src/unet_parts.py:15-22 (eval-mode BN)."""
    x = F.relu(_bn(sd, p + ".1", F.conv2d(x, sd[p + ".0.weight"], None, padding=1)))
    x = F.relu(_bn(sd, p + ".4", F.conv2d(x, sd[p + ".3.weight"], None, padding=1)))
    return x


def _up(sd, n, x, skip):
    """src/unet_parts.py:57-75: ConvTranspose2d 2x2 s2, cat([skip, up]), DoubleConv."""
    p = f"slices_generator.up{n}"
    x = F.conv_transpose2d(x, sd[p + ".up.weight"], sd[p + ".up.bias"], stride=2)
    dy, dx = skip.size(2) - x.size(2), skip.size(3) - x.size(3)
    x = F.pad(x, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return _double_conv(sd, p + ".conv.double_conv", torch.cat([skip, x], 1))


def unet_forward(sd, img, n_slices=12):
    """src/unet_custom.py:40-69.  Returns ([5 planes (B*K,C,H,W)], slices_rec (B*K,3,S,S))."""
    K = n_slices
    x1 = _down(sd, "down1", img)
    x2 = _down(sd, "down2", x1)
    x3 = _down(sd, "down3", x2)
    x4 = _down(sd, "down4", x3)
    x5 = _down(sd, "down5", x4)
    B, _, h, w = x5.shape

    def tile(x):  # expand_bs, :35-38
        b, c, hh, ww = x.shape
        return x.view(b, 1, c, hh, ww).expand(-1, K, -1, -1, -1).reshape(b * K, c, hh, ww)

    emb = sd["slices_generator.emds.weight"].view(1, K, 128, 1, 1).expand(B, K, 128, h, w).reshape(B * K, 128, h, w)
    g = "slices_generator."
    latent = F.conv2d(torch.cat([tile(x5), emb], 1), sd[g + "trans_c.weight"], sd[g + "trans_c.bias"])
    feats = [latent]
    x = latent
    for n, skip in zip((1, 2, 3, 4), (x4, x3, x2, x1)):
        s = F.conv2d(tile(skip), sd[g + f"trans_up{n}.weight"], sd[g + f"trans_up{n}.bias"])
        x = _up(sd, n, x, s)
        feats.append(x)
    out = torch.tanh(F.conv2d(x, sd[g + "outc.conv.weight"], sd[g + "outc.conv.bias"]))
    return feats, out


def project_coord(qry, T):
    """src/models.py:28-36.  qry (B,M,3), T (B,4,3) -> (B,M,2) in [-1,1]."""
    ones = torch.ones(qry.shape[0], qry.shape[1], 1, dtype=qry.dtype)
    p = torch.bmm(torch.cat([qry, ones], -1), T)
    uv = p[:, :, :2] / p[:, :, 2:]
    return torch.clamp(2 * (uv - 0.5), min=-1, max=1)


def sample_plane(plane, grid):
    """F.grid_sample(bilinear, zeros, align_corners=True) restated with explicit
    index arithmetic (src/models.py:38-46).  plane (N,C,H,W), grid (N,M,2) with
    grid[...,0]=x<->W, grid[...,1]=y<->H.  Returns (N,M,C)."""
    N, C, H, W = plane.shape
    ix = (grid[..., 0] + 1) * 0.5 * (W - 1)
    iy = (grid[..., 1] + 1) * 0.5 * (H - 1)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    wx1, wy1 = ix - x0, iy - y0
    wx0, wy0 = 1 - wx1, 1 - wy1
    flat = plane.reshape(N, C, H * W)
    out = torch.zeros(N, C, grid.shape[1], dtype=plane.dtype)
    for dx, dy, wgt in ((0, 0, wx0 * wy0), (1, 0, wx1 * wy0), (0, 1, wx0 * wy1), (1, 1, wx1 * wy1)):
        xx, yy = (x0 + dx).long(), (y0 + dy).long()
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        lin = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).unsqueeze(1).expand(-1, C, -1)
        out = out + torch.gather(flat, 2, lin) * (wgt * ok.to(plane.dtype)).unsqueeze(1)
    return out.permute(0, 2, 1)


def transformer_layer(sd, p, x, nhead=4):
    """nn.TransformerEncoderLayer(d_model=128, nhead=4, dim_ff=2048, relu, post-norm,
    batch_first) in eval mode (src/models.py:18-19).  x (N, L, 128)."""
    N, L, D = x.shape
    hd = D // nhead
    qkv = x @ sd[p + ".self_attn.in_proj_weight"].t() + sd[p + ".self_attn.in_proj_bias"]
    q, k, v = qkv.split(D, dim=-1)
    sh = lambda t: t.view(N, L, nhead, hd).transpose(1, 2)  # (N, h, L, hd)
    q, k, v = sh(q), sh(k), sh(v)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(N, L, D)
    o = o @ sd[p + ".self_attn.out_proj.weight"].t() + sd[p + ".self_attn.out_proj.bias"]
    x = F.layer_norm(x + o, (D,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], EPS_LN)
    h = F.relu(x @ sd[p + ".linear1.weight"].t() + sd[p + ".linear1.bias"])
    h = h @ sd[p + ".linear2.weight"].t() + sd[p + ".linear2.bias"]
    return F.layer_norm(x + h, (D,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], EPS_LN)


def decode(sd, feats, qry, T, n_slices=12, chunk=8192, return_tokens=False):
    """src/models.py:69-84 given the planes and the (already flipped/rotated) queries.
    feats: 5 planes (B*K,C,H,W); qry (B,M,3); T (B,4,3).  Returns sdf_pred (B,M)
    (and, with return_tokens, the token matrices (4, B*M, K+1, 128): after the token
    build and after each attention layer -- used to localise a failing stage)."""
    B, M, _ = qry.shape
    K = n_slices
    outs, toks = [], []
    for s in range(0, M, chunk):
        q = qry[:, s:s + chunk]
        m = q.shape[1]
        uv = project_coord(q, T)
        uvk = uv.view(B, 1, m, 2).expand(-1, K, -1, -1).reshape(B * K, m, 2)
        sampled = torch.cat([sample_plane(f, uvk) for f in feats], dim=2)  # (B*K, m, 992)
        C = sampled.shape[-1]
        sampled = sampled.view(B, K, m, C).permute(0, 2, 1, 3).reshape(B * m, K, C)
        tok_s = sampled @ sd["fc_s.weight"].t() + sd["fc_s.bias"]
        tok_q = (q @ sd["fc_p.weight"].t() + sd["fc_p.bias"]).view(B * m, 1, 128)
        x = torch.cat([tok_q, tok_s], 1)
        stages = [x]
        for l in range(3):
            x = transformer_layer(sd, f"att_decoder.layers.{l}", x)
            stages.append(x)
        if return_tokens:
            toks.append(torch.stack(stages, 0))
        t0 = x.view(B, m, K + 1, 128)[:, :, 0, :]
        outs.append((t0 @ sd["fc_out.0.weight"].t() + sd["fc_out.0.bias"]).squeeze(-1))
    if return_tokens:
        return torch.cat(outs, 1), torch.cat(toks, 1)
    return torch.cat(outs, 1)


def prepare_queries(qry_norot, obj_rot_mat, mode):
    """src/models.py:53-60.  Returns a new tensor (the reference's test mode flips
    y,z of the caller's tensor in place; the in-place effect is reproduced by the
    caller of this oracle when needed)."""
    if mode == "test":
        q = qry_norot.clone()
        q[:, :, 1:] *= -1
        return q
    return torch.bmm(qry_norot, obj_rot_mat)


def vgg_perceptual(sd, a, b):
    """src/vgg_perceptual_loss.py:51-71 restated.  Taps 1-4 are POST-ReLU and tap 5 is
    pre-ReLU: each slice ends on a conv, but the next slice starts with torchvision's
    ``ReLU(inplace=True)``, which overwrites the tensor already stored in the output list
    (vgg_perceptual_loss.py:33-38); only conv5_2 has no successor."""
    p = "vggptlossfunc."
    mean, std = sd[p + "mean"], sd[p + "std"]
    cfg = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]
    ranges = [(0, 3), (3, 8), (8, 13), (13, 22), (22, 31)]
    kinds = []
    for v in cfg:
        kinds += ["pool"] if v == "M" else ["conv", "relu"]

    def taps(x):
        x = ((x + 1) / 2.0 - mean) / std
        out = []
        for n, (lo, hi) in enumerate(ranges, start=1):
            for i in range(lo, hi):
                if kinds[i] == "conv":
                    x = F.conv2d(x, sd[f"{p}vgg.slice{n}.{i}.weight"], sd[f"{p}vgg.slice{n}.{i}.bias"], padding=1)
                elif kinds[i] == "relu":
                    x = F.relu(x)
                else:
                    x = F.max_pool2d(x, 2, 2)
            out.append(F.relu(x) if n < 5 else x)
        return out

    w = [1.0 / 2.6, 1.0 / 4.8, 1.0 / 3.7, 1.0 / 5.6, 10.0 / 1.5]
    return sum(wi * F.l1_loss(x, y) for wi, x, y in zip(w, taps(a), taps(b)))


def model_forward(sd, feed, mode="test", n_slices=12, with_vgg=True):
    """Slices3DRegModel.forward (src/models.py:48-94), eval-mode arithmetic."""
    img = feed["img_input"]
    B, _, S, _ = img.shape
    q = prepare_queries(feed["qry_norot"], feed.get("obj_rot_mat"), mode)
    feats, rec = unet_forward(sd, img, n_slices)
    sdf = decode(sd, feats, q, feed["trans_mat_wo_rot_tp"], n_slices)
    ret = {"sdf_pred": sdf, "slices_rec": rec.view(B, n_slices * 3, S, S), "feats": feats}
    if with_vgg:
        tgt = feed["img_slices"].view(B * n_slices, 3, S, S)
        ret["vgg_loss"] = vgg_perceptual(sd, rec, tgt) * 0.001
    return ret


def eval_points(sd, feed, chunk_size=3000, n_slices=12, hoist_encoder=True):
    """Generator3D.eval_points (reconstruct.py:74-102): chunk the queries, negate
    sdf_pred, concatenate.  With ``hoist_encoder`` the planes are computed once
    (identical values; the reference recomputes U-Net + VGG19 loss per chunk)."""
    qry = feed["qry_norot"]
    M = qry.shape[1]
    feats = unet_forward(sd, feed["img_input"], n_slices)[0] if hoist_encoder else None
    out = []
    for s in range(0, M, chunk_size):
        q = prepare_queries(qry[:, s:s + chunk_size], None, "test")
        if hoist_encoder:
            sdf = decode(sd, feats, q, feed["trans_mat_wo_rot_tp"], n_slices)
        else:
            f = dict(feed)
            f["qry_norot"] = qry[:, s:s + chunk_size].clone()
            sdf = model_forward(sd, f, "test", n_slices, with_vgg=True)["sdf_pred"]
        out.append(-sdf)
    return torch.cat(out, -1).squeeze(0)


def make_3d_grid(bb_min, bb_max, shape):
    """src_convonet/common.py:145-164: per-axis torch.linspace, x slowest / z fastest."""
    axes = [torch.linspace(bb_min[i], bb_max[i], shape[i]) for i in range(3)]
    gx = axes[0].view(-1, 1, 1).expand(*shape)
    gy = axes[1].view(1, -1, 1).expand(*shape)
    gz = axes[2].view(1, 1, -1).expand(*shape)
    return torch.stack([gx, gy, gz], dim=-1).reshape(-1, 3)


def dense_grid_values(sd, feed, nx, n_slices=12, box_size=1.0):
    """Generator3D.generate_from_latent dense branch (reconstruct.py:135-146)."""
    pts = box_size * make_3d_grid((-0.5,) * 3, (0.5,) * 3, (nx,) * 3)
    f = dict(feed)
    f["qry_norot"] = pts.unsqueeze(0)
    return eval_points(sd, f, 32768, n_slices).reshape(nx, nx, nx)


# ---------------------------------------------------------------------------------------------------------------
# Slices3DGTModel (src/model_gt.py:12-111, src/vgg16bn_feats.py:5-58): the 3-D stage of the generation-based pipeline
_GT_BLOCKS = [("conv1_2", "down1"), ("conv2_2", "down2"), ("conv3_3", "down3"), ("conv4_3", "down4"), ("conv5_3", "down5")]


def gt_encoder(sd, img_slices):
    """VGG16BNFeats.forward (vgg16bn_feats.py:44-58), eval mode: the five pre-BatchNorm taps of the trunk run on the
    (B*K,3,S,S) slice images (same cut points as the U-Net trunk; conv_last / classifier are computed by the reference
    but never used)."""
    x, feats = img_slices, []
    for name, block in _GT_BLOCKS:
        x = _down(sd, block, x, prefix=f"img_encoder.{name}.")
        feats.append(x)
    return feats


def gt_decode(sd, feats, qry, T, n_slices=12, chunk=8192):
    """model_gt.py:84-104 given the taps and the (already flipped / rotated) queries."""
    B, M, _ = qry.shape
    K = n_slices
    lin = lambda x, p: x @ sd[p + ".weight"].t() + sd[p + ".bias"]
    outs = []
    for s in range(0, M, chunk):
        q = qry[:, s:s + chunk]
        m = q.shape[1]
        uv = project_coord(q, T)
        uvk = uv.view(B, 1, m, 2).expand(-1, K, -1, -1).reshape(B * K, m, 2)
        sampled = torch.cat([sample_plane(f, uvk) for f in feats], dim=2)  # (B*K, m, 1472)
        sampled = sampled.view(B, K, m, 1472).permute(0, 2, 1, 3).reshape(B * m, K, 1472)
        tok_s = F.relu(lin(F.relu(lin(sampled, "fc_local.0")), "fc_local.2"))
        tq = q
        for i in (0, 2, 4):
            tq = F.relu(lin(tq, f"pts_feat_extractor.{i}"))
        x = torch.cat([tq.reshape(B * m, 1, 128), tok_s], 1)
        for l in range(3):
            x = transformer_layer(sd, f"att_decoder.layers.{l}", x)
        t0 = x.view(B, m, K + 1, 128)[:, :, 0, :]
        outs.append((t0 @ sd["fc_out.0.weight"].t() + sd["fc_out.0.bias"]).squeeze(-1))
    return torch.cat(outs, 1)


def gt_model_forward(sd, feed, mode="test", n_slices=12):
    """Slices3DGTModel.forward (model_gt.py:69-111), eval-mode arithmetic."""
    B, _, S, _ = feed["img_input"].shape
    q = prepare_queries(feed["qry_norot"], feed.get("obj_rot_mat"), mode)
    feats = gt_encoder(sd, feed["img_slices"].view(B * n_slices, 3, S, S))
    return {"sdf_pred": gt_decode(sd, feats, q, feed["trans_mat_wo_rot_tp"], n_slices), "feats": feats}
