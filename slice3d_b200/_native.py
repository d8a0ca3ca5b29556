"""ctypes binding of the C ABI in include/slice3d_b200.h.

PyTorch supplies device memory and streams only; every pointer handed to the
library is a raw ``data_ptr()``.  There is no CPU path: loading fails loudly if the
shared library has not been built, and ``NativeModel`` refuses non-CUDA tensors.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("S3D_LIB") or os.path.join(HERE, "_lib", "libslice3d_b200.so")  # S3D_LIB: kernel experiments

PREC_FP32, PREC_BF16X3, PREC_BF16, PREC_FP16X3, PREC_FP16F8 = 0, 1, 2, 3, 4
PRECISIONS = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16, "fp16x3": PREC_FP16X3, "fp16f8": PREC_FP16F8}
ABI_VERSION = 1

# every symbol include/slice3d_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "s3d_abi_version", "s3d_last_error", "s3d_model_create", "s3d_model_destroy", "s3d_model_n_slices",
    "s3d_planes_bytes", "s3d_encoder_workspace_bytes", "s3d_encoder_fwd", "s3d_decoder_workspace_bytes",
    "s3d_decoder_fwd", "s3d_decoder_batch_fwd", "s3d_decoder_grid_fwd", "s3d_decoder_debug_tokens", "s3d_vgg_loss_workspace_bytes",
    "s3d_vgg_loss_fwd", "s3d_vgg_loss_train_bytes", "s3d_vgg_loss_train_fwd", "s3d_vgg_loss_train_bwd", "s3d_mc_count", "s3d_mc_emit", "s3d_scan_scratch_bytes", "s3d_exclusive_scan", "s3d_debug_set_encoder", "s3d_debug_set_decoder_flags", "s3d_mise_scratch_ints",
    "s3d_mise_subdivide", "s3d_sparse_scratch_bytes", "s3d_sparse_rounds", "s3d_preprocess_workspace_bytes", "s3d_preprocess_rgba",
    "s3d_gt_encoder_workspace_bytes", "s3d_gt_encoder_fwd", "s3d_gt_decoder_workspace_bytes", "s3d_gt_decoder_fwd", "s3d_train_decoder_saved_bytes", "s3d_train_decoder_bwd_workspace_bytes",
    "s3d_train_decoder_fwd", "s3d_train_decoder_bwd", "s3d_selftest_umma", "s3d_debug_profile", "s3d_launch_count",
]


class S3DTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data_dev", C.c_void_p), ("numel", C.c_int64)]


class S3DGrid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("px_dev", C.c_void_p),
                ("py_dev", C.c_void_p), ("pz_dev", C.c_void_p)]


class NativeError(RuntimeError):
    pass


_lib = None


def lib():
    """Load (once) and type the shared library.  No fallback: a missing build is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(f"{LIB_PATH} not built; run `python -m slice3d_b200.build` "
                          "(slice3d_b200 has no PyTorch/CPU fallback for the inference path)")
    L = C.CDLL(LIB_PATH)
    L.s3d_abi_version.restype = C.c_int
    L.s3d_last_error.restype = C.c_char_p
    L.s3d_launch_count.restype = C.c_int64
    L.s3d_model_create.restype = C.c_int
    L.s3d_model_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(S3DTensor), C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p]
    L.s3d_model_destroy.restype = None
    L.s3d_model_destroy.argtypes = [C.c_void_p]
    L.s3d_model_n_slices.restype = C.c_int
    L.s3d_model_n_slices.argtypes = [C.c_void_p]
    L.s3d_planes_bytes.restype = C.c_size_t
    L.s3d_planes_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.s3d_encoder_workspace_bytes.restype = C.c_size_t
    L.s3d_encoder_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.s3d_encoder_fwd.restype = C.c_int
    L.s3d_encoder_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p),
                                  C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.s3d_decoder_workspace_bytes.restype = C.c_size_t
    L.s3d_decoder_workspace_bytes.argtypes = [C.c_int64, C.c_int32]
    L.s3d_decoder_fwd.restype = C.c_int
    L.s3d_decoder_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_int32, C.c_float, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
    L.s3d_decoder_batch_fwd.restype = C.c_int
    L.s3d_decoder_batch_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                                        C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t,
                                        C.c_void_p]
    L.s3d_decoder_grid_fwd.restype = C.c_int
    L.s3d_decoder_grid_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(S3DGrid), C.c_int64, C.c_int64,
                                       C.c_void_p, C.c_float, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t,
                                       C.c_void_p]
    L.s3d_decoder_debug_tokens.restype = C.c_int
    L.s3d_decoder_debug_tokens.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.s3d_vgg_loss_workspace_bytes.restype = C.c_size_t
    L.s3d_vgg_loss_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    L.s3d_vgg_loss_fwd.restype = C.c_int
    L.s3d_vgg_loss_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.c_size_t, C.c_void_p]
    L.s3d_vgg_loss_train_bytes.restype = C.c_size_t
    L.s3d_vgg_loss_train_bytes.argtypes = [C.c_int32, C.c_int32]
    L.s3d_vgg_loss_train_fwd.restype = C.c_int
    L.s3d_vgg_loss_train_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_size_t, C.c_void_p]
    L.s3d_vgg_loss_train_bwd.restype = C.c_int
    L.s3d_vgg_loss_train_bwd.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                         C.c_void_p]
    L.s3d_mc_count.restype = C.c_int
    L.s3d_mc_count.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p]
    L.s3d_mc_emit.restype = C.c_int
    L.s3d_mc_emit.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.s3d_scan_scratch_bytes.restype = C.c_size_t
    L.s3d_scan_scratch_bytes.argtypes = [C.c_int64]
    L.s3d_exclusive_scan.restype = C.c_int
    L.s3d_exclusive_scan.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.s3d_debug_set_encoder.restype = C.c_int
    L.s3d_debug_set_encoder.argtypes = [C.c_void_p, C.c_int32]
    L.s3d_mise_scratch_ints.restype = C.c_size_t
    L.s3d_mise_scratch_ints.argtypes = [C.c_int32, C.c_int32]
    L.s3d_mise_subdivide.restype = C.c_int
    L.s3d_mise_subdivide.argtypes = [C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    L.s3d_gt_encoder_workspace_bytes.restype = C.c_size_t
    L.s3d_gt_encoder_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.s3d_gt_encoder_fwd.restype = C.c_int
    L.s3d_gt_encoder_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p),
                                     C.c_void_p, C.c_size_t, C.c_void_p]
    L.s3d_gt_decoder_workspace_bytes.restype = C.c_size_t
    L.s3d_gt_decoder_workspace_bytes.argtypes = [C.c_int64, C.c_int32]
    L.s3d_gt_decoder_fwd.restype = C.c_int
    L.s3d_gt_decoder_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t,
                                     C.c_void_p]
    L.s3d_preprocess_workspace_bytes.restype = C.c_size_t
    L.s3d_preprocess_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.s3d_preprocess_rgba.restype = C.c_int
    L.s3d_preprocess_rgba.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p]
    L.s3d_sparse_scratch_bytes.restype = C.c_size_t
    L.s3d_sparse_scratch_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int64]
    L.s3d_sparse_rounds.restype = C.c_int
    L.s3d_sparse_rounds.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_double, C.c_float, C.c_int32, C.c_int32,
                                    C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
    L.s3d_train_decoder_saved_bytes.restype = C.c_size_t
    L.s3d_train_decoder_saved_bytes.argtypes = [C.c_void_p]
    L.s3d_train_decoder_bwd_workspace_bytes.restype = C.c_size_t
    L.s3d_train_decoder_bwd_workspace_bytes.argtypes = [C.c_void_p]
    L.s3d_train_decoder_fwd.restype = C.c_int
    L.s3d_train_decoder_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_size_t, C.c_void_p]
    L.s3d_train_decoder_bwd.restype = C.c_int
    L.s3d_train_decoder_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.s3d_selftest_umma.restype = C.c_int
    L.s3d_selftest_umma.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.s3d_debug_profile.restype = C.c_int
    L.s3d_debug_profile.argtypes = [C.POINTER(C.c_int64), C.c_int32]
    if L.s3d_abi_version() != ABI_VERSION:
        raise NativeError(f"ABI mismatch: library {L.s3d_abi_version()}, binding {ABI_VERSION}")
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise NativeError(f"slice3d_b200 error {rc}: {lib().s3d_last_error().decode()}")


def available_precisions():
    """Decoder arithmetic modes built into this revision of the library."""
    return ("fp32", "fp16x3", "fp16f8", "bf16x3", "bf16")


# precision="auto": the fastest <= 1e-4 mode (fp16f8) is used only if, on this checkpoint, a probe of AUTO_PROBE^3 lattice
# points evaluated on the caller's own planes stays within AUTO_TOL of the fp32 CUDA path; otherwise fp16x3.  The error of
# fp16f8 grows with the transformer's weight scales (tests: LayerNorm gains x 2 -> 1.6e-4, fp16x3 3.6e-5), so the
# default must not assume the synthetic checkpoint's 5e-5.  One probe per packed-weight handle (~10 ms).
AUTO_TOL = 7.5e-5
AUTO_PROBE = 16


def selftest_umma(mode, passes, a, w):
    """d = a . w^T on one tcgen05 tile (see include/slice3d_b200.h); a, w fp32 CUDA tensors."""
    a, w = _f32c(a, "a"), _f32c(w, "w")
    d = torch.empty(128, 128, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _check(lib().s3d_selftest_umma(mode, passes, a.data_ptr(), w.data_ptr(), d.data_ptr(), _stream(a.device)))
    return d


def mise_subdivide(res0, depth, threshold, value, known, cell_level, exists, flags):
    """One refinement step on the dense MISE state tensors (all on one CUDA device; flags int32 zeros, left zeroed)."""
    dev = value.device
    with torch.cuda.device(dev):
        _check(lib().s3d_mise_subdivide(res0, depth, float(threshold), value.data_ptr(), known.data_ptr(),
                                        cell_level.data_ptr(), exists.data_ptr(), flags.data_ptr(), _stream(dev)))


def mise_scratch_ints(res0, depth):
    return int(lib().s3d_mise_scratch_ints(res0, depth))


def marching_cubes(vol, isovalue, tri_table, tri_count):
    """vol (nx,ny,nz) float64 CUDA tensor; tri_table (256,15) int8 and tri_count (256,) int32 on the same device.
    -> (vertices (n,3) float64, triangles (m,3) int64): count, prefix sums and emit, all in the library."""
    if not vol.is_cuda or vol.dtype != torch.float64 or not vol.is_contiguous():
        raise NativeError("marching_cubes needs a contiguous float64 CUDA volume")
    nx, ny, nz = vol.shape
    dev, L = vol.device, lib()
    cells = (nx - 1) * (ny - 1) * (nz - 1)
    with torch.cuda.device(dev):
        vcount = torch.empty(cells, dtype=torch.int32, device=dev)
        tcount = torch.empty(cells, dtype=torch.int32, device=dev)
        owned = torch.empty(cells, dtype=torch.uint8, device=dev)
        _check(L.s3d_mc_count(vol.data_ptr(), nx, ny, nz, float(isovalue), tri_count.data_ptr(), vcount.data_ptr(),
                              tcount.data_ptr(), owned.data_ptr(), _stream(dev)))
        # the running vertex / triangle counters of the sequential scan: exclusive prefix sums (s3d_exclusive_scan)
        vbase = torch.empty(cells, dtype=torch.int64, device=dev)
        tbase = torch.empty(cells, dtype=torch.int64, device=dev)
        totals = torch.empty(2, dtype=torch.int64, device=dev)
        scratch = torch.empty(L.s3d_scan_scratch_bytes(cells), dtype=torch.uint8, device=dev)
        st = _stream(dev)
        _check(L.s3d_exclusive_scan(vcount.data_ptr(), cells, vbase.data_ptr(), totals.data_ptr(), scratch.data_ptr(), st))
        _check(L.s3d_exclusive_scan(tcount.data_ptr(), cells, tbase.data_ptr(), totals.data_ptr() + 8, scratch.data_ptr(), st))
        n_v, n_t = (int(x) for x in totals.tolist())  # the one host read: the output sizes
        verts = torch.empty((n_v, 3), dtype=torch.float64, device=dev)
        tris = torch.empty((n_t, 3), dtype=torch.int64, device=dev)
        if n_v:
            _check(L.s3d_mc_emit(vol.data_ptr(), nx, ny, nz, float(isovalue), tri_table.data_ptr(), vbase.data_ptr(),
                                 tbase.data_ptr(), tcount.data_ptr(), owned.data_ptr(), verts.data_ptr(), tris.data_ptr(),
                                 _stream(dev)))
    return verts, tris


PROFILE_FIELDS = ["token", "vec", "wait_qkv", "attn", "wait_out", "ln1", "ffn_wait_d1", "ffn_math", "ffn_wait_hfree",
                  "ffn_store", "wait_ffn", "ln2", "mma_wait_a", "mma_wait_full", "mma_wait_h", "mma_wait_d1free",
                  "mma_total", "prod_wait_empty", "prod_total", "tiles", "att_kstage", "att_scores", "att_vstage", "att_pv"]


def debug_profile(reset=True):
    """Per-phase cycle counters of the tensor-core decoder (summed over CTAs) as a dict."""
    torch.cuda.synchronize()
    buf = (C.c_int64 * 32)()
    _check(lib().s3d_debug_profile(buf, 1 if reset else 0))
    return {k: int(buf[i]) for i, k in enumerate(PROFILE_FIELDS)}


def launch_count():
    return int(lib().s3d_launch_count())


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32c(t, what):
    if not t.is_cuda:
        raise NativeError(f"{what} must be a CUDA tensor (slice3d_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise NativeError(f"{what} must be float32")
    return t if t.is_contiguous() else t.contiguous()


class Planes:
    """Device blob produced by the encoder for a batch of input views: per image and per
    scale s a (K, R_s, R_s, 128) fp32 channels-last plane = fc_s_s applied to feature plane s."""

    def __init__(self, blob, B, K, S, slices_rec):
        self.blob, self.B, self.K, self.S, self.slices_rec = blob, B, K, S, slices_rec
        self.bytes_per_image = blob.numel() * 4 // B

    def image_ptr(self, b):
        return self.blob.data_ptr() + b * self.bytes_per_image


class NativeModel:
    """Owns one s3d_model handle (packed/folded weights on one device)."""

    def __init__(self, state_dict, n_slices, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise NativeError("NativeModel needs a CUDA device (slice3d_b200 has no CPU path)")
        self.device = device
        self.K = n_slices
        L = lib()
        keep, arr = [], []
        for name, t in state_dict.items():
            if t.dtype != torch.float32 or name.startswith("att_layer."):
                continue
            t = t.detach().to(device=device).contiguous()
            keep.append(t)
            arr.append(S3DTensor(name.encode(), t.data_ptr(), t.numel()))
        tensors = (S3DTensor * len(arr))(*arr)
        h = C.c_void_p()
        with torch.cuda.device(device):
            torch.cuda.current_stream(device).synchronize()
            _check(L.s3d_model_create(C.byref(h), tensors, len(arr), n_slices, device.index or 0, _stream(device)))
        self._h = h
        self._ws = {}
        self._auto = None       # precision "auto" resolved for this handle
        self.auto_info = None   # {"selected", "probe_points", "fp16f8_max_abs_vs_fp32"}

    def resolve_precision(self, precision, probe=None):
        """'auto' -> the mode selected for this handle (probing on first use: ``probe(precision) -> values`` evaluates the
        probe points with the caller's planes and camera); anything else is returned unchanged."""
        if precision != "auto":
            return precision
        if self._auto is None:
            if probe is None:
                raise NativeError("precision='auto' has not been resolved for this model yet")
            ref, fast = probe("fp32"), probe("fp16f8")
            err = float((ref - fast).abs().max())
            ok = err <= AUTO_TOL  # (NaN compares false: falls back)
            self._auto = "fp16f8" if ok else "fp16x3"
            self.auto_info = {"selected": self._auto, "probe_points": int(ref.numel()), "fp16f8_max_abs_vs_fp32": err,
                              "tolerance": AUTO_TOL}
            if not ok:
                import warnings
                warnings.warn(f"slice3d_b200: fp16f8 decoder differs from the fp32 path by {err:.2e} on this checkpoint "
                              f"(> {AUTO_TOL:.1e}); precision='auto' selects fp16x3")
        return self._auto

    def _probe_points(self):
        ax = torch.linspace(-0.45, 0.45, AUTO_PROBE, device=self.device)
        g = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
        return g.contiguous()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.s3d_model_destroy(h)

    def _workspace(self, key, nbytes):
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def release_workspaces(self):
        self._ws.clear()

    # ---- encoder ----------------------------------------------------------------
    def encode(self, img, want_feats=False, want_slices_rec=True):
        """img (B,3,S,S) -> Planes (+ the five raw NCHW feature planes when asked)."""
        img = _f32c(img, "img_input")
        B, c, S, S2 = img.shape
        if c != 3 or S != S2:
            raise NativeError("img_input must be (B,3,S,S)")
        L, K = lib(), self.K
        with torch.cuda.device(self.device):
            blob = torch.empty(L.s3d_planes_bytes(B, K, S) // 4, dtype=torch.float32, device=self.device)
            rec = torch.empty(B * K, 3, S, S, dtype=torch.float32, device=self.device) if want_slices_rec else None
            feats, fptr = None, None
            if want_feats:
                chans = [512, 256, 128, 64, 32]
                feats = [torch.empty(B * K, chans[s], (S // 16) << s, (S // 16) << s, dtype=torch.float32,
                                     device=self.device) for s in range(5)]
                fptr = (C.c_void_p * 5)(*[f.data_ptr() for f in feats])
            nws = L.s3d_encoder_workspace_bytes(B, K, S)
            ws = self._workspace("enc", nws)
            _check(L.s3d_encoder_fwd(self._h, img.data_ptr(), B, S, blob.data_ptr(), fptr,
                                     rec.data_ptr() if rec is not None else None, ws.data_ptr(), ws.numel(),
                                     _stream(self.device)))
        planes = Planes(blob, B, K, S, rec)
        return (planes, feats) if want_feats else planes

    # ---- perceptual loss --------------------------------------------------------
    def vgg_loss(self, a, b):
        """a, b (N,3,S,S) fp32 CUDA in [-1,1] -> 0-dim tensor: VGGPerceptualLoss.forward(a, b)['pt_c_loss']."""
        a, b = _f32c(a, "input_img"), _f32c(b, "target_img")
        if a.shape != b.shape or a.dim() != 4 or a.shape[1] != 3 or a.shape[2] != a.shape[3]:
            raise NativeError("vgg_loss: a and b must both be (N,3,S,S)")
        N, S, L = a.shape[0], a.shape[2], lib()
        with torch.cuda.device(self.device):
            out = torch.empty((), dtype=torch.float32, device=self.device)
            ws = self._workspace("vgg", L.s3d_vgg_loss_workspace_bytes(N, S))
            _check(L.s3d_vgg_loss_fwd(self._h, a.data_ptr(), b.data_ptr(), N, S, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                      _stream(self.device)))
        return out

    # ---- decoder ----------------------------------------------------------------
    def decode(self, planes, b, qry, T, rot=None, flip_in_place=False, out_scale=1.0, precision="auto", out=None):
        """qry (n,3) of image b -> (n,) = out_scale * sdf_pred.  rot None = test mode (y,z negated)."""
        if not qry.is_cuda or qry.dtype != torch.float32 or not qry.is_contiguous():
            raise NativeError("qry must be a contiguous float32 CUDA tensor")
        T = _f32c(T, "trans_mat_wo_rot_tp")
        rot = _f32c(rot, "obj_rot_mat") if rot is not None else None
        n = qry.shape[0]
        precision = self.resolve_precision(precision, lambda p: self.decode(planes, b, self._probe_points(), T, precision=p))
        L, prec = lib(), PRECISIONS[precision]
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty(n, dtype=torch.float32, device=self.device)
            ws = self._workspace("dec", L.s3d_decoder_workspace_bytes(n, prec))
            _check(L.s3d_decoder_fwd(self._h, planes.image_ptr(b), planes.S, qry.data_ptr(), n, T.data_ptr(),
                                     rot.data_ptr() if rot is not None else None, 1 if flip_in_place else 0,
                                     out_scale, out.data_ptr(), prec, ws.data_ptr(), ws.numel(),
                                     _stream(self.device)))
        return out

    def decode_batch(self, planes, qry, T, rot=None, flip_in_place=False, out_scale=1.0, precision="auto", out=None):
        """All images of an encoder batch in ONE launch: qry (B,n,3), T (B,4,3), rot (B,3,3) or None -> (B,n)."""
        if not qry.is_cuda or qry.dtype != torch.float32 or not qry.is_contiguous() or qry.dim() != 3:
            raise NativeError("qry must be a contiguous float32 CUDA tensor (B,n,3)")
        B, n = qry.shape[0], qry.shape[1]
        if B != planes.B:
            raise NativeError("qry batch does not match the encoder batch")
        T = _f32c(T, "trans_mat_wo_rot_tp")
        rot = _f32c(rot, "obj_rot_mat") if rot is not None else None
        precision = self.resolve_precision(precision, lambda p: self.decode(planes, 0, self._probe_points(), T[0], precision=p))
        L, prec = lib(), PRECISIONS[precision]
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty(B, n, dtype=torch.float32, device=self.device)
            ws = self._workspace("dec", L.s3d_decoder_workspace_bytes(B * n, prec))
            _check(L.s3d_decoder_batch_fwd(self._h, planes.image_ptr(0), planes.S, qry.data_ptr(), B, n, T.data_ptr(),
                                           rot.data_ptr() if rot is not None else None, 1 if flip_in_place else 0,
                                           out_scale, out.data_ptr(), prec, ws.data_ptr(), ws.numel(),
                                           _stream(self.device)))
        return out

    # ---- Slices3DGTModel --------------------------------------------------------
    def encode_gt(self, img_slices, B, want_taps=False):
        """img_slices (B*K,3,S,S) -> Planes (fc_local's first Linear applied to the five trunk taps)."""
        img = _f32c(img_slices, "img_slices")
        N, c, S, S2 = img.shape
        K = self.K
        if c != 3 or S != S2 or N != B * K:
            raise NativeError("img_slices must be (B*K,3,S,S)")
        L = lib()
        with torch.cuda.device(self.device):
            blob = torch.empty(L.s3d_planes_bytes(B, K, S) // 4, dtype=torch.float32, device=self.device)
            taps, tptr = None, None
            if want_taps:
                chans = [64, 128, 256, 512, 512]
                taps = [torch.empty(N, chans[i], S >> i, S >> i, dtype=torch.float32, device=self.device) for i in range(5)]
                tptr = (C.c_void_p * 5)(*[t.data_ptr() for t in taps])
            ws = self._workspace("enc", L.s3d_gt_encoder_workspace_bytes(B, K, S))
            _check(L.s3d_gt_encoder_fwd(self._h, img.data_ptr(), B, S, blob.data_ptr(), tptr, ws.data_ptr(), ws.numel(),
                                        _stream(self.device)))
        planes = Planes(blob, B, K, S, None)
        return (planes, taps) if want_taps else planes

    def decode_gt(self, planes, qry, T, rot=None, flip_in_place=False, out_scale=1.0, precision="auto", out=None):
        """qry (B,n,3), T (B,4,3), rot (B,3,3) or None -> (B,n): the GT model's per-query path in the library."""
        if not qry.is_cuda or qry.dtype != torch.float32 or not qry.is_contiguous() or qry.dim() != 3:
            raise NativeError("qry must be a contiguous float32 CUDA tensor (B,n,3)")
        B, n = qry.shape[0], qry.shape[1]
        if B != planes.B:
            raise NativeError("qry batch does not match the encoder batch")
        T = _f32c(T, "trans_mat_wo_rot_tp")
        rot = _f32c(rot, "obj_rot_mat") if rot is not None else None
        precision = self.resolve_precision(
            precision, lambda p: self.decode_gt(planes, self._probe_points().unsqueeze(0).repeat(B, 1, 1), T, precision=p))
        L, prec = lib(), PRECISIONS[precision]
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty(B, n, dtype=torch.float32, device=self.device)
            ws = self._workspace("dec", L.s3d_gt_decoder_workspace_bytes(B * n, prec))
            _check(L.s3d_gt_decoder_fwd(self._h, planes.image_ptr(0), planes.S, qry.data_ptr(), B, n, T.data_ptr(),
                                        rot.data_ptr() if rot is not None else None, 1 if flip_in_place else 0, out_scale,
                                        out.data_ptr(), prec, ws.data_ptr(), ws.numel(), _stream(self.device)))
        return out

    def decode_grid(self, planes, b, axes, first, count, T, out_scale=1.0, precision="auto", out=None):
        """Grid points [first, first+count) of the (nx,ny,nz) grid given by the three per-axis
        coordinate tensors ``axes`` (x slowest, z fastest), test-mode flip applied on the fly."""
        px, py, pz = (_f32c(a, "grid axis") for a in axes)
        T = _f32c(T, "trans_mat_wo_rot_tp")
        precision = self.resolve_precision(precision, lambda p: self.decode(planes, b, self._probe_points(), T, precision=p))
        L, prec = lib(), PRECISIONS[precision]
        g = S3DGrid(px.numel(), py.numel(), pz.numel(), px.data_ptr(), py.data_ptr(), pz.data_ptr())
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty(count, dtype=torch.float32, device=self.device)
            ws = self._workspace("dec", L.s3d_decoder_workspace_bytes(count, prec))
            _check(L.s3d_decoder_grid_fwd(self._h, planes.image_ptr(b), planes.S, C.byref(g), first, count,
                                          T.data_ptr(), out_scale, out.data_ptr(), prec, ws.data_ptr(), ws.numel(),
                                          _stream(self.device)))
        return out

    def sparse_rounds(self, planes, b, T, box_size, out_scale, mise, scratch, capacity, counts, n_rounds, precision):
        """Enqueue ``n_rounds`` device-resident MISE rounds (s3d_sparse_rounds) on the state tensors of ``mise``."""
        T = _f32c(T, "trans_mat_wo_rot_tp")
        precision = self.resolve_precision(precision, lambda p: self.decode(planes, b, self._probe_points(), T, precision=p))
        L, prec = lib(), PRECISIONS[precision]
        with torch.cuda.device(self.device):
            ws = self._workspace("dec", L.s3d_decoder_workspace_bytes(capacity, prec))
            _check(L.s3d_sparse_rounds(self._h, planes.image_ptr(b), planes.S, T.data_ptr(), float(box_size), out_scale,
                                       mise.resolution_0, mise.depth, mise.threshold, mise.value.data_ptr(),
                                       mise.known.data_ptr(), mise.cell_level.data_ptr(), mise.exists.data_ptr(),
                                       mise.flags().data_ptr(), scratch.data_ptr(), capacity, counts.data_ptr(), n_rounds,
                                       prec, ws.data_ptr(), ws.numel(), _stream(self.device)))

    def debug_tokens(self, planes, b, qry, T, rot=None):
        """fp32 validation path: returns (sdf (n,), tokens (4,n,K+1,128))."""
        qry, T = _f32c(qry, "qry"), _f32c(T, "T")
        rot = _f32c(rot, "rot") if rot is not None else None
        n, L = qry.shape[0], lib()
        with torch.cuda.device(self.device):
            out = torch.empty(n, dtype=torch.float32, device=self.device)
            tok = torch.empty(4, n, self.K + 1, 128, dtype=torch.float32, device=self.device)
            ws = self._workspace("dec", L.s3d_decoder_workspace_bytes(n, PREC_FP32))
            _check(L.s3d_decoder_debug_tokens(self._h, planes.image_ptr(b), planes.S, qry.data_ptr(), n,
                                              T.data_ptr(), rot.data_ptr() if rot is not None else None,
                                              out.data_ptr(), tok.data_ptr(), ws.data_ptr(), ws.numel(),
                                              _stream(self.device)))
        return out, tok
