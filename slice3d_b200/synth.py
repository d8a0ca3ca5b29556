"""Deterministic synthetic checkpoints and inputs (no network: no pretrained weights).

``synthetic_state_dict`` fills every tensor of a ``Slices3DRegModel``-shaped
state_dict from a per-key seeded CPU generator, so the reference module (in
``oracle/make_golden.py``), the oracle and this package can be loaded with
bit-identical weights irrespective of module-construction order.

Scales are chosen so activations stay O(1) through the 13-conv trunk and the
planes genuinely influence ``sdf_pred`` (He-style conv init, non-trivial BN
running statistics and LayerNorm affines), which makes the parity tests
sensitive to errors in every stage.

``synthetic_inputs`` builds the feed_dict of SURVEY.md section 8(d): uniform
+-1 images, the camera of the reference's single-image dataset creator
(reference: create_dataset_sin_img.py:56-62 -> datasets.py:123-140) and a
``make_3d_grid`` query set.
"""
import math
import zlib

import torch

# trans_mat_wo_rot_tp for az=el=0, dist=1.2 (K: f=35/32, c=0.5 on the unit image).
CAMERA_T = [[1.09375, 0.0, 0.0], [0.0, 1.09375, 0.0], [0.5, 0.5, 1.0], [0.6, 0.60000006, 1.2]]
OBJ_ROT = [[0.0, 0.0, -1.0], [0.0, -1.0, 0.0], [-1.0, 0.0, 0.0]]


def _gen(key, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
    return g


def synthetic_state_dict(template, seed=0):
    """template: mapping name -> tensor (only shape/dtype are used)."""
    out = {}
    for key, ref in template.items():
        shape, g = tuple(ref.shape), _gen(key, seed)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros(shape, dtype=torch.int64)
        elif leaf == "running_mean":
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif key.endswith("vggptlossfunc.mean"):
            t = torch.tensor([0.485, 0.456, 0.406]).view(shape)
        elif key.endswith("vggptlossfunc.std"):
            t = torch.tensor([0.229, 0.224, 0.225]).view(shape)
        elif key.endswith("emds.weight"):
            t = torch.randn(shape, generator=g)
        elif leaf in ("bias", "in_proj_bias"):
            t = torch.randn(shape, generator=g) * 0.1
        elif len(shape) == 1:  # BN / LayerNorm scale
            t = torch.rand(shape, generator=g) * 0.4 + 0.8
        elif len(shape) == 4:
            if ".up.weight" in key:  # ConvTranspose2d (Cin, Cout, 2, 2): one tap per output pixel
                fan_in = shape[0]
                std = math.sqrt(1.0 / fan_in)
            else:
                fan_in = shape[1] * shape[2] * shape[3]
                std = math.sqrt(2.0 / fan_in)
                if "trans_" in key or "outc" in key:  # linear 1x1 adapters (no ReLU after)
                    std = math.sqrt(1.0 / fan_in)
            t = torch.randn(shape, generator=g) * std
        elif len(shape) == 2:  # Linear / in_proj
            fan_in = shape[1]
            std = math.sqrt(1.0 / fan_in)
            if key.endswith("linear1.weight"):
                std = math.sqrt(2.0 / fan_in)
            if key.endswith("fc_p.weight"):
                std = 1.0
            t = torch.randn(shape, generator=g) * std
        else:
            raise ValueError(f"no synthetic rule for {key} {shape}")
        out[key] = t.to(ref.dtype)
    return out


def make_3d_grid(bb_min, bb_max, shape):
    """Dense query grid, x slowest / z fastest (reference: src_convonet/common.py:145-164).

    Uses torch.linspace per axis exactly like the reference so coordinates are
    bit-identical (arange*step differs by <= 8.9e-8).
    """
    nx, ny, nz = shape
    px = torch.linspace(bb_min[0], bb_max[0], nx)
    py = torch.linspace(bb_min[1], bb_max[1], ny)
    pz = torch.linspace(bb_min[2], bb_max[2], nz)
    gx = px.view(-1, 1, 1).expand(nx, ny, nz)
    gy = py.view(1, -1, 1).expand(nx, ny, nz)
    gz = pz.view(1, 1, -1).expand(nx, ny, nz)
    return torch.stack([gx, gy, gz], dim=-1).reshape(nx * ny * nz, 3)


def synthetic_inputs(img_size, n_slices=12, seed=0, batch=1):
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    img_input = torch.rand(batch, 3, img_size, img_size, generator=g) * 2 - 1
    img_slices = torch.rand(batch, 3 * n_slices, img_size, img_size, generator=g) * 2 - 1
    T = torch.tensor(CAMERA_T, dtype=torch.float32).unsqueeze(0).repeat(batch, 1, 1)
    R = torch.tensor(OBJ_ROT, dtype=torch.float32).unsqueeze(0).repeat(batch, 1, 1)
    return {"img_input": img_input, "img_slices": img_slices, "trans_mat_wo_rot_tp": T, "obj_rot_mat": R}


def sample_grid_indices(nx, n, seed=0):
    """A reproducible subset of a nx^3 grid: the 8 corners, face points on the
    clamp boundary and uniformly random interior indices (int64, sorted, unique)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(2000 + seed + nx)
    m = nx - 1
    corners = [(a * nx + b) * nx + c for a in (0, m) for b in (0, m) for c in (0, m)]
    rnd = torch.randint(0, nx ** 3, (n,), generator=g).tolist()
    # near-face points (z index = last ones map to the near plane after the y,z flip)
    face = [((i * 37) % nx * nx + (i * 53) % nx) * nx + (0 if i % 2 else m) for i in range(64)]
    idx = torch.tensor(sorted(set(corners + rnd + face)), dtype=torch.int64)
    return idx


def synthetic_train_batch(img_size, n_slices=12, batch=4, n_qry=256, seed=0):
    """One training batch of BASELINE configs[4]: the dataset's keys (reference datasets.py:169-177)
    with random rotations, queries in the unit box and sdf = randn * 0.1; everything from one seeded CPU generator."""
    feed = synthetic_inputs(img_size, n_slices, seed=seed, batch=batch)
    g = torch.Generator(device="cpu")
    g.manual_seed(5000 + seed)
    feed["qry_norot"] = torch.rand(batch, n_qry, 3, generator=g) - 0.5
    feed["obj_rot_mat"] = torch.linalg.qr(torch.randn(batch, 3, 3, generator=g))[0].contiguous()
    feed["sdf"] = torch.randn(batch, n_qry, generator=g) * 0.1
    feed["occ"] = (feed["sdf"] < 0).float()
    return feed


def set_dropout(model, p):
    """Set every dropout probability of a Slices3DRegModel-shaped module (the transformer's nn.Dropout modules and the
    attention-weight dropout of nn.MultiheadAttention).  Gradient parity tests use p = 0 on both sides."""
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = p
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = p
    return model
