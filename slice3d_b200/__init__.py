"""slice3d_b200 -- B200-native implementation of Slice3D's slice-to-3D hot path.

Public surface mirrors the reference (reg_slices/src/models.py, reg_slices/reconstruct.py):
``Slices3DRegModel`` and ``Generator3D``; the arithmetic lives in the CUDA library behind the
C ABI declared in ``include/slice3d_b200.h``.
"""
from .models import Slices3DRegModel  # noqa: F401
from .model_gt import Slices3DGTModel  # noqa: F401
from .generator import Generator3D  # noqa: F401
from .mcubes import Mesh, marching_cubes  # noqa: F401
from .mise import MISE  # noqa: F401
from . import inputs  # noqa: F401
from .datasets import Slice3DDataset  # noqa: F401
from .synth import make_3d_grid  # noqa: F401
from .train import (cal_acc, cal_loss_pred, cal_loss_pred_gt, fit, latest_checkpoint, save_checkpoint,  # noqa: F401
                    train_step, train_step_gt, val_step, val_step_gt, wrap_ddp)

__all__ = ["Slices3DRegModel", "Slices3DGTModel", "Generator3D", "MISE", "Mesh", "marching_cubes", "make_3d_grid", "train_step", "val_step",
           "cal_loss_pred", "cal_acc", "wrap_ddp", "train_step_gt", "val_step_gt", "cal_loss_pred_gt", "Slice3DDataset", "fit", "save_checkpoint",
           "latest_checkpoint"]
