"""Input pipeline of ``Slice3DDataset.__getitem__`` (reference: reg_slices/src/datasets.py:89-177,
reg_slices/src/utils.py:29-73,132-170) with the per-pixel work on the GPU.

What the reference does per sample, in DataLoader worker processes on the host: open 1 + 12 PNGs, composite their alpha
channel (``png_2_whitebg`` / ``png_2_rgb``), ``T.Resize`` each PIL image to (S, S), ``T.ToTensor``, ``T.Normalize``; order the
slices X1..X4, Z4..Z1, Y1..Y4; build the camera matrices from the stored (azimuth, elevation, distance); scale / offset /
subsample the SDF samples.  Here:

* ``preprocess_rgba``  -- the image half for a whole batch of DECODED RGBA arrays in two CUDA kernels
  (``csrc/inputs.cu`` behind ``s3d_preprocess_rgba``): compositing, Pillow's antialiased bilinear resample of the 8-bit
  image bit for bit, to-tensor, normalise.  PNG inflate stays on the host (a byte-serial entropy decoder, not GPU work).
* ``resample_tables``  -- Pillow's ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` (src/libImaging/Resample.c) in float64 on
  the host: the integer tables the kernels consume.  ``preprocess_rgba_host`` applies the same tables with numpy (what the
  CPU tests compare with PIL / torchvision themselves).
* ``camera_matrices``  -- ``getBlenderProj`` / ``get_rotate_matrix`` / ``get_W2O_mat`` and datasets.py:123-140 restated in
  float64 (twelve numbers per sample: host arithmetic).
* ``prepare_queries``  -- datasets.py:142-167: scale / offset / threshold shift of the SDF samples, occupancy, subsampling.
* ``SLICE_ORDER`` / ``assemble_sample`` -- the feed_dict layout of datasets.py:169-177.
"""
import math

import numpy as np
import torch

# file stems of the 12 slice images in the order they are concatenated (datasets.py:107-118)
SLICE_ORDER = [f"X_{i}" for i in (1, 2, 3, 4)] + [f"Z_{i}" for i in (4, 3, 2, 1)] + [f"Y_{i}" for i in (1, 2, 3, 4)]
PRECISION_BITS = 32 - 8 - 2


def _bilinear(x):
    x = -x if x < 0.0 else x
    return 1.0 - x if x < 1.0 else 0.0


def resample_tables(in_size, out_size):
    """Pillow's coefficient tables for an antialiased bilinear resample of ``in_size`` -> ``out_size`` samples over the
    whole axis: (bounds (out,2) int32 = (first index, tap count), coeffs (out, ksize) int32 fixed point, ksize)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bilinear((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def composite_host(rgba, white_bg):
    """png_2_whitebg / png_2_rgb (datasets.py:75-88) on a (N,H,W,4) uint8 array -> (N,H,W,3) uint8."""
    rgb, alpha = rgba[..., :3], rgba[..., 3:4]
    if white_bg:
        a0 = (alpha == 0).astype(np.float32)
        return (np.ones(rgb.shape) * 255 * a0 + rgb * (1 - a0)).astype(np.uint8)
    return (rgb * (alpha / 255.0)).astype(np.uint8)


def _resample_axis_host(img, bounds, kk, axis):
    out_shape = list(img.shape)
    out_shape[axis] = bounds.shape[0]
    out = np.empty(out_shape, dtype=np.uint8)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    dst = np.moveaxis(out, axis, 0)
    for xx in range(bounds.shape[0]):
        lo, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt):
            acc += src[lo + x] * int(kk[xx, x])
        dst[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def preprocess_rgba_host(rgba, img_size, white_bg):
    """The whole image half on the host with the same integer tables (numpy): (N,H,W,4) uint8 -> (N,3,S,S) float32."""
    rgba = np.asarray(rgba)
    N, H, W, _ = rgba.shape
    img = composite_host(rgba, white_bg)
    bh, kh, _ = resample_tables(W, img_size)
    bv, kv, _ = resample_tables(H, img_size)
    img = _resample_axis_host(img, bh, kh, 2)   # horizontal pass first, 8-bit intermediate (ImagingResample)
    img = _resample_axis_host(img, bv, kv, 1)
    t = torch.from_numpy(img).permute(0, 3, 1, 2).to(torch.float32).div(255)
    return t.sub(0.5).div(0.5)


_TABLE_CACHE = {}


def preprocess_rgba(rgba, img_size, white_bg=False):
    """(N,H,W,4) uint8 CUDA tensor of decoded PNGs -> (N,3,S,S) float32 in [-1,1] on the same device (two kernels)."""
    from . import _native
    import ctypes as C
    if not rgba.is_cuda or rgba.dtype != torch.uint8 or rgba.dim() != 4 or rgba.shape[3] != 4:
        raise _native.NativeError("preprocess_rgba needs a (N,H,W,4) uint8 CUDA tensor (no CPU path: see preprocess_rgba_host)")
    rgba = rgba.contiguous()
    N, H, W, _ = rgba.shape
    dev = rgba.device
    key = (str(dev), H, W, img_size)
    if key not in _TABLE_CACHE:
        bh, kh, ksh = resample_tables(W, img_size)
        bv, kv, ksv = resample_tables(H, img_size)
        _TABLE_CACHE[key] = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (bh, kh, bv, kv)) + (ksh, ksv)
    bh, kh, bv, kv, ksh, ksv = _TABLE_CACHE[key]
    L = _native.lib()
    with torch.cuda.device(dev):
        out = torch.empty(N, 3, img_size, img_size, dtype=torch.float32, device=dev)
        ws = torch.empty(L.s3d_preprocess_workspace_bytes(N, H, img_size), dtype=torch.uint8, device=dev)
        _native._check(L.s3d_preprocess_rgba(rgba.data_ptr(), N, H, W, img_size, 1 if white_bg else 0, bh.data_ptr(),
                                             kh.data_ptr(), ksh, bv.data_ptr(), kv.data_ptr(), ksv, out.data_ptr(),
                                             ws.data_ptr(), ws.numel(), _native._stream(dev)))
    return out


# ------------------------------------------------------------------------------------------------ camera
_CAM_ROT = np.asarray([[1.910685676922942e-15, 4.371138828673793e-08, 1.0],
                       [1.0, -4.371138828673793e-08, -0.0],
                       [4.371138828673793e-08, 1.0, -4.371138828673793e-08]])


def blender_proj(az, el, distance, img_w=1, img_h=1):
    """utils.py:29-73 (getBlenderProj): intrinsic K (3,3) and extrinsic RT (3,4) of the Blender camera."""
    f_u = 35.0 * img_w * 1.0 / 32.0
    f_v = 35.0 * img_h * 1.0 * 1.0 / 32.0
    K = np.array([[f_u, 0.0, img_w * 1.0 / 2], [0.0, f_v, img_h * 1.0 / 2], [0.0, 0.0, 1.0]])
    sa, ca, se, ce = np.sin(-az), np.cos(-az), np.sin(-el), np.cos(-el)
    R_world2obj = np.array([[ca * ce, -sa, ca * se], [sa * ce, ca, sa * se], [-se, 0.0, ce]]).T
    R_obj2cam = _CAM_ROT.T
    R_world2cam = R_obj2cam @ R_world2obj
    T_world2cam = -1 * R_obj2cam @ np.array([[distance], [0.0], [0.0]])
    R_camfix = np.array([[1.0, 0, 0], [0, -1.0, 0], [0, 0, -1.0]])
    return K, np.hstack((R_camfix @ R_world2cam, R_camfix @ T_world2cam))


def rotate_matrix(angle):
    """utils.py:132-170 (get_rotate_matrix): neg . Rz . Rz . scale_y_neg . Rx."""
    c, s = np.cos(angle), np.sin(angle)
    rx = np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1.0]])
    rz = np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
    sy = np.diag([1.0, -1.0, 1.0, 1.0])
    neg = np.diag([-1.0, -1.0, -1.0, 1.0])
    return np.linalg.multi_dot([neg, rz, rz, sy, rx])


def camera_matrices(az, el, distance):
    """datasets.py:123-140 with ``az`` already negated by the caller as the dataset does (az = -data[1][idx]):
    (obj_rot_mat (3,3) float32, trans_mat_wo_rot_tp (4,3) float32)."""
    K, RT = blender_proj(az, el, distance, img_w=1, img_h=1)
    W2O = np.eye(4)  # get_W2O_mat((0, 0, 0)): translation by zero
    rot_full = np.linalg.multi_dot([RT, rotate_matrix(-np.pi / 2)])
    obj_rot_mat = np.transpose(rot_full)[:3, :]
    tmp = np.concatenate((np.eye(3), rot_full[:, 3:4]), axis=1)
    trans = np.linalg.multi_dot([K, tmp, W2O])
    return (torch.tensor(np.asarray(obj_rot_mat)).float(), torch.tensor(np.transpose(np.asarray(trans))).float())


# ------------------------------------------------------------------------------------------------ queries
def prepare_queries(sdf_npy, scale, offset, n_qry, split="train", rng=None):
    """datasets.py:142-167: ``sdf_npy`` (n,4) = xyz + sdf sampled at the 0.003 level set.  Returns qry (n_qry,3),
    occ, sdf as float32 tensors.  train: a fresh permutation (``rng``: a numpy Generator/RandomState or None);
    val/test: numpy's legacy seed 1234, exactly like the reference."""
    pt = sdf_npy[:, :3] * scale + np.array([offset[0], offset[2], -offset[1]])
    val = (sdf_npy[:, 3] - 0.003) * scale
    occ = (val <= 0).astype(np.float32)
    if split == "train":
        # the reference re-seeds numpy's global generator from the OS before every draw (``np.random.seed()``), so that
        # forked DataLoader workers do not repeat each other's permutations: a fresh generator does the same
        perm = (rng if rng is not None else np.random.default_rng()).permutation(len(pt))[:n_qry]
    else:
        perm = np.random.RandomState(1234).permutation(len(pt))[:n_qry]
    return torch.tensor(pt[perm]).float(), torch.tensor(occ[perm]).float(), torch.tensor(val[perm]).float()


def assemble_sample(images, az, el, distance, qry, occ, sdf):
    """``images``: (13,3,S,S) = the input view followed by the 12 slices in SLICE_ORDER (as ``preprocess_rgba`` returns
    them for one sample) -> the feed_dict of datasets.py:169-177."""
    rot, T = camera_matrices(az, el, distance)
    S = images.shape[-1]
    return {"img_input": images[0], "qry_norot": qry, "obj_rot_mat": rot.to(images.device),
            "trans_mat_wo_rot_tp": T.to(images.device), "occ": occ, "sdf": sdf, "img_slices": images[1:].reshape(36, S, S)}
