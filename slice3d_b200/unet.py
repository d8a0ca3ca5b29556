"""Parameter container for the slice generator / plane encoder.

Mirrors the *state_dict layout* of the reference ``UNet``
(reference: reg_slices/src/unet_custom.py:5-32, reg_slices/src/unet_parts.py:8-84):
a VGG16-BN trunk cut into ``down1..down5_`` keeping torchvision's original
``features`` indices, a 1x1 ``trans_c``, four ``Up`` stages (ConvTranspose2d
2x2/s2 + DoubleConv), four 1x1 skip adapters ``trans_up*``, ``outc`` and the
slice embedding ``emds``.

This module owns parameters only.  Inference arithmetic runs in the CUDA
library (``csrc/encoder.cu``); the train-mode arithmetic (batch-stat BN,
autograd) is ``forward_train`` below, written with torch ops.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

# torchvision VGG16 configuration "D"; 'M' is a 2x2 max-pool.
_VGG16_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]


class _IndexedSequential(nn.Sequential):
    """nn.Sequential whose children keep caller-chosen (non-contiguous) names.

    The reference slices ``vgg.features[a:b]``; nn.Sequential slicing keeps the
    original integer keys, so e.g. ``down2`` has children "4".."10"
    (reference: unet_custom.py:15-20).  We rebuild that without torchvision.
    """

    def __init__(self, named):
        super().__init__()
        for name, mod in named:
            self.add_module(str(name), mod)


def _vgg16_bn_feature_list():
    layers = []
    cin = 3
    for v in _VGG16_CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(cin, v, kernel_size=3, padding=1), nn.BatchNorm2d(v), nn.ReLU(inplace=True)]
            cin = v
    return layers  # 44 entries, indices as in torchvision.models.vgg16_bn().features


class DoubleConv(nn.Module):
    """(conv3x3 no-bias -> BN -> ReLU) x2 (reference: unet_parts.py:8-25)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
            nn.Conv2d(cout, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.double_conv(x)


class Up(nn.Module):
    """ConvTranspose2d(2,2) then cat([skip, up]) then DoubleConv (reference: unet_parts.py:42-75)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.up = nn.ConvTranspose2d(cin, cin // 2, kernel_size=2, stride=2)
        self.conv = DoubleConv(cin, cout)

    def forward(self, x1, x2):
        x1 = self.up(x1)
        dy = x2.size(2) - x1.size(2)
        dx = x2.size(3) - x1.size(3)
        x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        return self.conv(torch.cat([x2, x1], dim=1))


class OutConv(nn.Module):
    """1x1 conv + tanh (reference: unet_parts.py:78-84)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=1)
        self.act = nn.Tanh()

    def forward(self, x):
        return self.act(self.conv(x))


class UNet(nn.Module):
    """Slice generator.  ``n_slices`` is a constructor argument here (the reference
    hard-codes 12, unet_custom.py:9); the default keeps the reference layout."""

    def __init__(self, n_channels=3, n_slices=12):
        super().__init__()
        self.n_channels = n_channels
        self.n_slices = n_slices
        self.dim_embed = 128
        feats = _vgg16_bn_feature_list()
        cut = lambda a, b: _IndexedSequential([(i, feats[i]) for i in range(a, b)])
        self.down1 = cut(0, 4)
        self.down2 = cut(4, 11)
        self.down3 = cut(11, 21)
        self.down4 = cut(21, 31)
        self.down5 = cut(31, 41)
        self.down5_ = cut(41, 44)
        self.trans_c = nn.Conv2d(512 + self.dim_embed, 512, 1)
        self.up1 = Up(512, 256)
        self.trans_up1 = nn.Conv2d(512, 256, 1)
        self.up2 = Up(256, 128)
        self.trans_up2 = nn.Conv2d(256, 128, 1)
        self.up3 = Up(128, 64)
        self.trans_up3 = nn.Conv2d(128, 64, 1)
        self.up4 = Up(64, 32)
        self.trans_up4 = nn.Conv2d(64, 32, 1)
        self.outc = OutConv(32, 3)
        self.emds = nn.Embedding(self.n_slices, self.dim_embed)

    def expand_bs(self, x):
        b, c, h, w = x.shape
        return x.view(b, 1, c, h, w).expand(-1, self.n_slices, -1, -1, -1).reshape(b * self.n_slices, c, h, w)

    def forward_train(self, x):
        """Autograd-capable arithmetic of UNet.forward (reference: unet_custom.py:40-69).

        Used in train/val mode only (batch-statistics BN, gradients).  ``down5_``
        is evaluated and discarded exactly like the reference (:48) so that its
        BN running statistics evolve identically in train mode.
        """
        x1 = self.down1(x)
        x2 = self.down2(x1)
        x3 = self.down3(x2)
        x4 = self.down4(x3)
        x5 = self.down5(x4)
        _ = self.down5_(x5)
        b, _, h, w = x5.shape
        k = self.n_slices
        emb = self.emds.weight.view(1, k, self.dim_embed, 1, 1).expand(b, k, self.dim_embed, h, w)
        emb = emb.reshape(b * k, self.dim_embed, h, w)
        latent = self.trans_c(torch.cat([self.expand_bs(x5), emb], 1))
        feats = [latent]
        y = self.up1(latent, self.trans_up1(self.expand_bs(x4)))
        feats.append(y)
        y = self.up2(y, self.trans_up2(self.expand_bs(x3)))
        feats.append(y)
        y = self.up3(y, self.trans_up3(self.expand_bs(x2)))
        feats.append(y)
        y = self.up4(y, self.trans_up4(self.expand_bs(x1)))
        feats.append(y)
        return feats, self.outc(y)
