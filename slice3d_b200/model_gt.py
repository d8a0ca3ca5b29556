"""``Slices3DGTModel`` -- drop-in for the reference module of the same name, the 3-D stage of the generation-based
pipeline (reference: reg_slices/src/model_gt.py:12-111, reg_slices/src/vgg16bn_feats.py:5-58).

Same constructor, ``forward(feed_dict) -> {'sdf_pred'}`` contract and ``state_dict`` layout.  The 12 GIVEN slice images
(``img_slices``) go through a VGG16-BN trunk; its five pre-BatchNorm taps (64 @ S ... 512 @ S/16, 1472 channels) are
sampled at every query's projection, ``fc_local`` turns each slice's sample into a token, ``pts_feat_extractor`` makes
the query token, and the same 3-layer transformer + ``fc_out`` as in ``Slices3DRegModel`` gives the SDF.

Inference (``eval()`` under ``no_grad``) runs in the CUDA library: the trunk on the tcgen05 convolution kernel with the
first ``fc_local`` Linear hoisted onto the taps (sampling is linear), a token kernel + one tcgen05 GEMM for the rest of
``fc_local``, and the fused tensor-core decoder reading ready tokens.  Training uses torch autograd ops.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native
from .models import NativeHandleMixin, Slices3DRegModel, default_precision
from .unet import _IndexedSequential, _vgg16_bn_feature_list


class VGG16BNFeats(nn.Module):
    """vgg16bn_feats.py:5-58: torchvision's vgg16_bn features cut at the convolution BEFORE each block's last
    BatchNorm (original layer indices kept as names), plus the never-used ``conv_last`` / ``classifier``."""

    def __init__(self):
        super().__init__()
        feats = _vgg16_bn_feature_list()
        cut = lambda a, b: _IndexedSequential([(i, feats[i]) for i in range(a, b)])
        self.conv1_2 = cut(0, 4)
        self.conv2_2 = cut(4, 11)
        self.conv3_3 = cut(11, 21)
        self.conv4_3 = cut(21, 31)
        self.conv5_3 = cut(31, 41)
        self.conv_last = cut(41, 44)
        self.classifier = nn.Linear(512 * 4 * 4, 128)

    def forward(self, img):
        c1 = self.conv1_2(img)
        c2 = self.conv2_2(c1)
        c3 = self.conv3_3(c2)
        c4 = self.conv4_3(c3)
        c5 = self.conv5_3(c4)
        last = self.conv_last(c5)
        g = self.classifier(torch.flatten(last, 1)) if last.shape[-1] == 4 else None  # needs S = 128 (reference: always)
        return [c1, c2, c3, c4, c5], g


class Slices3DGTModel(NativeHandleMixin, nn.Module):
    def __init__(self, img_size=128, n_slices=12, mode="train", precision=None):
        super().__init__()
        self.mode = mode
        self.img_encoder = VGG16BNFeats()
        self.img_size = img_size
        self.n_slices = n_slices
        self.att_layer = nn.TransformerEncoderLayer(d_model=128, nhead=4, batch_first=True)
        self.att_decoder = nn.TransformerEncoder(self.att_layer, num_layers=3)
        self.fc_out = nn.Sequential(nn.Linear(128, 1))
        self.pts_feat_extractor = nn.Sequential(nn.Linear(3, 32), nn.ReLU(), nn.Linear(32, 64), nn.ReLU(),
                                                nn.Linear(64, 128), nn.ReLU())
        self.fc_local = nn.Sequential(nn.Linear(1472, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU())
        self.fc_global = nn.Sequential(nn.Linear(128 + 128, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU())  # unused
        # --- not part of the reference API ---
        self.precision = precision or default_precision(n_slices)
        self.fused_eval_points = True
        self._nat = {"epoch": 0, "dev": {}}
        self._enc_cache = None

    project_coord = staticmethod(Slices3DRegModel.project_coord)

    def encode(self, img_slices):
        """Planes of the given slice images (B, 3K, S, S); cached per tensor object / version."""
        nat = self.native()
        c = self._enc_cache
        if c is not None and c["img"] is img_slices and c["ver"] == img_slices._version:
            return c["planes"]
        B, _, S, _ = img_slices.shape
        planes = nat.encode_gt(img_slices.view(B * self.n_slices, 3, S, S), B)
        self._enc_cache = {"img": img_slices, "ver": img_slices._version, "planes": planes}
        return planes

    def forward(self, feed_dict):
        if self.training or torch.is_grad_enabled():
            return self._forward_autograd(feed_dict)
        img_slices = feed_dict["img_slices"]
        if not img_slices.is_cuda:
            raise _native.NativeError("inference needs CUDA tensors: slice3d_b200 has no CPU/PyTorch fallback")
        nat = self.native()
        planes = self.encode(img_slices)
        qry, T = feed_dict["qry_norot"], feed_dict["trans_mat_wo_rot_tp"]
        if self.mode == "test":
            if qry.is_contiguous() and qry.dtype == torch.float32:
                sdf = nat.decode_gt(planes, qry, T, None, True, 1.0, self.precision)
            else:
                tmp = qry.float().contiguous()
                sdf = nat.decode_gt(planes, tmp, T, None, True, 1.0, self.precision)
                qry.copy_(tmp)
        else:
            sdf = nat.decode_gt(planes, qry.float().contiguous(), T, feed_dict["obj_rot_mat"], False, 1.0, self.precision)
        return {"sdf_pred": sdf}

    def _forward_autograd(self, feed_dict):
        """model_gt.py:69-111 with torch ops (train / val with gradients)."""
        img_input = feed_dict["img_input"]
        n_bs, _, S, _ = img_input.shape
        K = self.n_slices
        qry = feed_dict["qry_norot"]
        if self.mode == "test":
            qry[:, :, 1:] *= -1
        else:
            qry = torch.bmm(qry, feed_dict["obj_rot_mat"])
        n_qry = qry.shape[1]
        img_slices = feed_dict["img_slices"].view(n_bs, K, 3, S, S).view(n_bs * K, 3, S, S)
        feats, _ = self.img_encoder(img_slices)
        uv = self.project_coord(qry, feed_dict["trans_mat_wo_rot_tp"])
        grid = uv.view(n_bs, 1, 1, n_qry, 2).expand(-1, K, -1, -1, -1).reshape(n_bs * K, 1, n_qry, 2)
        sampled = [F.grid_sample(f, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
                   .permute(0, 3, 2, 1).reshape(n_bs * K, n_qry, f.shape[1]) for f in feats]
        agg = torch.cat(sampled, dim=2).view(n_bs, K, n_qry, 1472).permute(0, 2, 1, 3).reshape(n_bs, n_qry, K, 1472)
        tok = torch.cat([self.pts_feat_extractor(qry).view(n_bs * n_qry, 1, 128),
                         self.fc_local(agg).view(n_bs * n_qry, K, 128)], 1)
        att = self.att_decoder(tok).view(n_bs, n_qry, K + 1, 128)[:, :, 0, :]
        return {"sdf_pred": self.fc_out(att).squeeze(-1)}
