"""Train-mode decoder as an autograd op over the hand-written CUDA kernels of ``csrc/train_decoder.cu``.

``decoder_train(feats, qry, T, params, ...)`` = the per-query half of ``Slices3DRegModel.forward`` in training
(reference: reg_slices/src/models.py:57-84 under ``model.train()``; gradients consumed by ``loss.backward()`` in
reg_slices/train.py:41-53): projection, 5 x grid_sample, fc_s / fc_p, the 3-layer transformer with dropout, fc_out.
Forward and backward both run in the library (``s3d_train_decoder_fwd`` / ``s3d_train_decoder_bwd``); torch only
carries the tensors and the autograd graph edge to the U-Net's feature planes.
"""
import ctypes as C

import torch

from . import _native


class TrainCfg(C.Structure):
    _fields_ = [("B", C.c_int32), ("n_qry", C.c_int32), ("K", C.c_int32), ("S", C.c_int32), ("dropout_p", C.c_float),
                ("seed", C.c_uint64)]


def param_list(model):
    """The 42 decoder parameters in the order include/slice3d_b200.h documents."""
    ps = [model.fc_p.weight, model.fc_p.bias, model.fc_s.weight, model.fc_s.bias]
    for layer in model.att_decoder.layers:
        ps += [layer.self_attn.in_proj_weight, layer.self_attn.in_proj_bias, layer.self_attn.out_proj.weight,
               layer.self_attn.out_proj.bias, layer.linear1.weight, layer.linear1.bias, layer.linear2.weight,
               layer.linear2.bias, layer.norm1.weight, layer.norm1.bias, layer.norm2.weight, layer.norm2.bias]
    ps += [model.fc_out[0].weight, model.fc_out[0].bias]
    return ps


def _ptrs(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _DecoderTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, qry, T, *tensors):
        B, n_qry, K, S, p, seed = cfg
        dev = qry.device
        feats = [t.detach().contiguous() for t in tensors[:5]]
        params = [t.detach().contiguous() for t in tensors[5:]]
        qry, T = qry.detach().float().contiguous(), T.detach().float().contiguous()
        L = _native.lib()
        c = TrainCfg(B, n_qry, K, S, p, seed)
        nbytes = L.s3d_train_decoder_saved_bytes(C.byref(c))
        if nbytes == 0:
            raise _native.NativeError("train decoder: " + L.s3d_last_error().decode())
        with torch.cuda.device(dev):
            saved = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
            sdf = torch.empty(B, n_qry, dtype=torch.float32, device=dev)
            _native._check(L.s3d_train_decoder_fwd(C.byref(c), _ptrs(feats), qry.data_ptr(), T.data_ptr(), _ptrs(params),
                                                   sdf.data_ptr(), saved.data_ptr(), saved.numel() * 4, _native._stream(dev)))
        ctx.cfg, ctx.saved, ctx.qry, ctx.T, ctx.params = cfg, saved, qry, T, params
        ctx.feat_shapes = [f.shape for f in feats]
        return sdf

    @staticmethod
    def backward(ctx, dsdf):
        B, n_qry, K, S, p, seed = ctx.cfg
        dev = dsdf.device
        L = _native.lib()
        c = TrainCfg(B, n_qry, K, S, p, seed)
        dsdf = dsdf.float().contiguous()
        with torch.cuda.device(dev):
            dfeats = [torch.zeros(s, dtype=torch.float32, device=dev) for s in ctx.feat_shapes]
            dparams = [torch.empty_like(t) for t in ctx.params]
            ws = torch.empty(L.s3d_train_decoder_bwd_workspace_bytes(C.byref(c)) // 4 + 1, dtype=torch.float32, device=dev)
            _native._check(L.s3d_train_decoder_bwd(C.byref(c), ctx.qry.data_ptr(), ctx.T.data_ptr(), _ptrs(ctx.params),
                                                   dsdf.data_ptr(), ctx.saved.data_ptr(), ctx.saved.numel() * 4,
                                                   _ptrs(dfeats), _ptrs(dparams), ws.data_ptr(), ws.numel() * 4,
                                                   _native._stream(dev)))
        ctx.saved = None
        return (None, None, None, *dfeats, *dparams)


def decoder_train(feats, qry, T, params, n_slices, img_size, dropout_p=0.0, seed=0):
    """feats: the U-Net's five NCHW planes (B*K, C_s, R_s, R_s); qry (B, n_qry, 3) in model space (already rotated);
    T (B,4,3); params = ``param_list(model)``.  Returns sdf_pred (B, n_qry), differentiable w.r.t. feats and params."""
    if not qry.is_cuda:
        raise _native.NativeError("decoder_train needs CUDA tensors (the CPU path is the torch restatement in models.py)")
    B, n_qry = qry.shape[0], qry.shape[1]
    cfg = (int(B), int(n_qry), int(n_slices), int(img_size), float(dropout_p), int(seed))
    return _DecoderTrainFn.apply(cfg, qry, T, *feats, *params)


# ------------------------------------------------------------------------------------------------ perceptual loss
class _VGGLossTrainFn(torch.autograd.Function):
    """VGGPerceptualLoss.forward(a, b)['pt_c_loss'] with its gradient w.r.t. ``a`` in the CUDA library
    (``s3d_vgg_loss_train_fwd / _bwd``: the forward convolutions and the data-gradient convolutions of the frozen VGG19
    on the tcgen05 kernel)."""

    @staticmethod
    def forward(ctx, nat, a, b):
        a, b = a.detach().float().contiguous(), b.detach().float().contiguous()
        N, S, dev = a.shape[0], a.shape[2], a.device
        L = _native.lib()
        with torch.cuda.device(dev):
            saved = torch.empty(L.s3d_vgg_loss_train_bytes(N, S) // 4 + 1, dtype=torch.float32, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            _native._check(L.s3d_vgg_loss_train_fwd(nat._h, a.data_ptr(), b.data_ptr(), N, S, loss.data_ptr(), saved.data_ptr(),
                                                    saved.numel() * 4, _native._stream(dev)))
        ctx.nat, ctx.saved, ctx.shape = nat, saved, a.shape
        return loss

    @staticmethod
    def backward(ctx, gout):
        N, _, S, _ = ctx.shape
        dev = gout.device
        L = _native.lib()
        gout = gout.detach().float().contiguous()
        with torch.cuda.device(dev):
            grad = torch.empty(ctx.shape, dtype=torch.float32, device=dev)
            _native._check(L.s3d_vgg_loss_train_bwd(ctx.nat._h, N, S, gout.data_ptr(), ctx.saved.data_ptr(),
                                                    ctx.saved.numel() * 4, grad.data_ptr(), _native._stream(dev)))
        ctx.saved = None
        return None, grad, None


_VGG_HANDLES = {}


def vgg_handle(vgg_module):
    """Library handle holding ONLY the frozen perceptual network of ``vgg_module`` (a VGGPerceptualLoss), cached per
    (device, tensor identities / versions): during training the model's other weights change every step, these never do."""
    sd = {"vggptlossfunc." + k: v for k, v in vgg_module.state_dict().items()}
    dev = next(iter(sd.values())).device
    # parameters: identity + version (load_state_dict / .to() change them); buffers (mean, std): identity only -- DDP
    # re-broadcasts buffers before every forward, which bumps their version without changing a bit
    key = (str(dev),) + tuple(t.data_ptr() for t in sd.values()) + tuple(p._version for p in vgg_module.parameters())
    ent = _VGG_HANDLES.get(id(vgg_module))
    if ent is None or ent[0] != key:
        ent = (key, _native.NativeModel(sd, 12, dev))
        _VGG_HANDLES[id(vgg_module)] = ent
    return ent[1]


def vgg_loss_train(vgg_module, a, b):
    """pt_c_loss of VGGPerceptualLoss (vgg_perceptual_loss.py:51-71), differentiable w.r.t. ``a`` (N,3,S,S)."""
    if not a.is_cuda:
        raise _native.NativeError("vgg_loss_train needs CUDA tensors")
    return _VGGLossTrainFn.apply(vgg_handle(vgg_module), a, b)
