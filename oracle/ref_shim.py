"""ORACLE support -- import the UNMODIFIED reference from /root/reference (build container only).

Three shims, no reference file touched:
  1. torchvision.models.vgg16_bn / vgg19 ignore ``pretrained=True`` (no network);
  2. torch.Tensor.cuda -> identity on CPU-only hosts (models.py:31 hard-codes .cuda());
  3. sys.path gets /root/reference/reg_slices.
Never imported by tests marked gpu, smoke() or bench.py: /root/reference does not
exist on the GPU box.
"""
import os
import sys

import torch
import torchvision

REF_ROOT = "/root/reference/reg_slices"


def available():
    return os.path.isdir(REF_ROOT)


def import_reference():
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    v16, v19 = torchvision.models.vgg16_bn, torchvision.models.vgg19
    torchvision.models.vgg16_bn = lambda pretrained=False, **k: v16(weights=None)
    torchvision.models.vgg19 = lambda pretrained=False, **k: v19(weights=None)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from src.models import Slices3DRegModel  # noqa: E402
    from src_convonet.common import make_3d_grid  # noqa: E402
    return Slices3DRegModel, make_3d_grid


def build_reference_model(img_size, mode, state_dict, n_slices=12):
    """Reference module loaded with ``state_dict`` (strict).  For n_slices != 12 the
    two hard-coded attributes are patched as described in SURVEY.md section 0."""
    Model, _ = import_reference()
    m = Model(img_size=img_size, n_slices=n_slices, mode=mode)
    if n_slices != 12:
        m.slices_generator.n_slices = n_slices
        m.slices_generator.emds = torch.nn.Embedding(n_slices, 128)
    m.load_state_dict(state_dict, strict=True)
    return m.eval()
