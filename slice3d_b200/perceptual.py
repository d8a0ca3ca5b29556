"""Frozen VGG19 perceptual loss (train-time loss term and checkpoint layout).

State-dict layout follows the reference ``VGGPerceptualLoss`` / ``VGG19Feats``
(reference: reg_slices/src/vgg_perceptual_loss.py:6-71): buffers ``mean``/``std``
and 28 frozen tensors ``vgg.slice{1..5}.<torchvision idx>.{weight,bias}``.
The five taps are the outputs of conv1_2, conv2_2, conv3_2, conv4_2, conv5_2
(features[0:3], [3:8], [8:13], [13:22], [22:31]).  Because torchvision's VGG uses
``ReLU(inplace=True)`` and every slice but the last is followed by one, taps 1-4 are
rectified in place by the next slice before the loss reads them (so they are effectively
post-ReLU) while conv5_2 stays pre-ReLU; the in-place ReLUs are kept here so the
arithmetic -- and its gradient -- is the reference's.

Not on the inference hot path: the reference evaluates and discards it on every
test-time chunk (models.py:90-92); here it is evaluated in test mode only when
``Slices3DRegModel.test_time_vgg_loss`` is set.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

_VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]
_SLICE_RANGES = [(0, 3), (3, 8), (8, 13), (13, 22), (22, 31)]
_TAP_WEIGHTS = [1.0 / 2.6, 1.0 / 4.8, 1.0 / 3.7, 1.0 / 5.6, 10.0 / 1.5]


def _vgg19_feature_list():
    layers, cin = [], 3
    for v in _VGG19_CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(2, 2))
        else:
            layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(inplace=True)]
            cin = v
    return layers


class VGG19Feats(nn.Module):
    def __init__(self):
        super().__init__()
        feats = _vgg19_feature_list()
        for n, (a, b) in enumerate(_SLICE_RANGES, start=1):
            seq = nn.Sequential()
            for i in range(a, b):
                seq.add_module(str(i), feats[i])
            setattr(self, f"slice{n}", seq)
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, img):
        taps = []
        for n in range(1, 6):
            img = getattr(self, f"slice{n}")(img)
            taps.append(img)
        return taps


class VGGPerceptualLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.vgg = VGG19Feats()
        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))

    def forward(self, input_img, target_img):
        a = ((input_img + 1) / 2.0 - self.mean) / self.std
        b = ((target_img + 1) / 2.0 - self.mean) / self.std
        fa, fb = self.vgg(a), self.vgg(b)
        loss = 0.0
        for w, x, y in zip(_TAP_WEIGHTS, fa, fb):
            loss = loss + w * F.l1_loss(x, y)
        return {"pt_c_loss": loss, "pt_s_loss": 0.0}
