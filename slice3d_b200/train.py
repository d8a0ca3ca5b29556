"""Training-loop pieces of the reference's ``train.py`` for ``Slices3DRegModel`` (reference: reg_slices/train.py:21-53,
70-93, 131-136) and of ``train_gt.py`` for ``Slices3DGTModel`` (reg_slices/train_gt.py:21-71: ``*_gt``), plus the
one-process-per-GPU wrapper BASELINE configs[4] asks for.

``train_step`` keeps the reference's signature and return values (the three loss terms and the sign accuracy as Python
floats).  With CUDA tensors the decoder's forward and backward run in the CUDA library (``slice3d_b200.train_ops``);
the U-Net / VGG19 convolutions run through torch autograd.  ``wrap_ddp`` replaces the reference's
``torch.nn.DataParallel`` (train.py:131-132) by DistributedDataParallel: per-replica BatchNorm statistics exactly like
DataParallel, gradients averaged over ranks with bucketed NCCL all-reduces.  14 parameters never receive a gradient
(``att_layer.*`` is the unused original of the deep-copied encoder layers, ``down5_.41.*`` feeds a discarded tensor;
SURVEY.md section 3.3), hence ``find_unused_parameters``.
"""
import torch
import torch.nn.functional as F


def cal_acc(x, gt, pred_type="sdf"):
    """train.py:21-27."""
    if pred_type == "occ":
        acc = ((x["occ_pred"].sigmoid() > 0.5) == (gt["occ"] > 0.5)).float().sum(dim=-1) / x["occ_pred"].shape[1]
    else:
        acc = ((x["sdf_pred"] >= 0) == (gt["sdf"] >= 0)).float().sum(dim=-1) / x["sdf_pred"].shape[1]
    return acc.mean(-1)


def cal_loss_pred(x, gt, pred_type="sdf"):
    """train.py:29-39."""
    if pred_type == "occ":
        loss_pred = F.binary_cross_entropy_with_logits(x["occ_pred"], gt["occ"])
    else:
        loss_pred = F.l1_loss(x["sdf_pred"], gt["sdf"])
    return loss_pred, F.l1_loss(x["slices_rec"], gt["img_slices"]), x["vgg_loss"]


def train_step(batch, model, opt, args=None, device=None):
    """train.py:41-53.  ``args`` only needs ``pred_type``; ``device`` defaults to the model's."""
    pred_type = getattr(args, "pred_type", "sdf")
    device = device or next(model.parameters()).device
    for key in batch:
        batch[key] = batch[key].to(device, non_blocking=True)
    opt.zero_grad()
    x = model(batch)
    loss_pred, loss_img, loss_img_vgg = cal_loss_pred(x, batch, pred_type)
    loss = loss_pred + loss_img + loss_img_vgg
    loss.backward()
    opt.step()
    with torch.no_grad():
        acc = cal_acc(x, batch, pred_type)
    return loss_pred.item(), loss_img.item(), loss_img_vgg.item(), acc.item()


@torch.no_grad()
def val_step(model, val_loader, pred_type="sdf", device=None):
    """train.py:70-93 without the PNG contact sheet: average L1(sdf) and sign accuracy over the loader, plus the last
    batch's image loss (what the reference returns and writes into the checkpoint name)."""
    device = device or next(model.parameters()).device
    avg_loss_pred, avg_acc, ni, loss_img = 0.0, 0.0, 0, None
    for batch in val_loader:
        for key in batch:
            batch[key] = batch[key].to(device, non_blocking=True)
        x = model(batch)
        loss_pred, loss_img, _ = cal_loss_pred(x, batch, pred_type)
        avg_loss_pred += loss_pred.item()
        avg_acc += cal_acc(x, batch, pred_type).item()
        ni += 1
    return avg_loss_pred / max(ni, 1), avg_acc / max(ni, 1), loss_img


# ---------------------------------------------------------------------- Slices3DGTModel (reference: reg_slices/train_gt.py)
def cal_loss_pred_gt(x, gt, pred_type="sdf"):
    """train_gt.py:29-36: the GT-slices model has no image head, so the loss is the prediction term alone."""
    if pred_type == "occ":
        return F.binary_cross_entropy_with_logits(x["occ_pred"], gt["occ"])
    return F.l1_loss(x["sdf_pred"], gt["sdf"])


def train_step_gt(batch, model, opt, args=None, device=None):
    """train_gt.py:38-50 for ``Slices3DGTModel``: returns (loss_pred, acc) as Python floats."""
    pred_type = getattr(args, "pred_type", "sdf")
    device = device or next(model.parameters()).device
    for key in batch:
        batch[key] = batch[key].to(device, non_blocking=True)
    opt.zero_grad()
    x = model(batch)
    loss_pred = cal_loss_pred_gt(x, batch, pred_type)
    loss_pred.backward()
    opt.step()
    with torch.no_grad():
        acc = cal_acc(x, batch, pred_type)
    return loss_pred.item(), acc.item()


@torch.no_grad()
def val_step_gt(model, val_loader, pred_type="sdf", device=None):
    """train_gt.py:53-71: average prediction loss and sign accuracy over the loader (an empty loader gives zeros where
    the reference divides by zero)."""
    device = device or next(model.parameters()).device
    avg_loss_pred, avg_acc, ni = 0.0, 0.0, 0
    for batch in val_loader:
        for key in batch:
            batch[key] = batch[key].to(device, non_blocking=True)
        x = model(batch)
        avg_loss_pred += cal_loss_pred_gt(x, batch, pred_type).item()
        avg_acc += cal_acc(x, batch, pred_type).item()
        ni += 1
    return avg_loss_pred / max(ni, 1), avg_acc / max(ni, 1)


# ---------------------------------------------------------------------- the epoch loop, checkpoints, resume
def latest_checkpoint(dir_ckpt):
    """train.py:138-140: the most recently created file of the checkpoint directory (None when there is none)."""
    import glob
    import os
    names = [n for n in glob.glob(os.path.join(dir_ckpt, "*")) if n.endswith(".ckpt")]
    return max(names, key=os.path.getctime) if names else None


def save_checkpoint(dir_ckpt, model, opt, n_epoch, n_iter, val_metrics):
    """train.py:166-169 / train_gt.py: ``{'model', 'opt', 'n_epoch', 'n_iter'}`` of the UNWRAPPED module under the
    reference's file name ``<epoch>_<iter>_<metric:.4>_..._.ckpt`` (the validation figures, four significant digits)."""
    import os
    os.makedirs(dir_ckpt, exist_ok=True)
    inner = model.module if hasattr(model, "module") else model
    name = "_".join([str(n_epoch), str(n_iter)] + [f"{float(v):.4}" for v in val_metrics]) + ".ckpt"
    path = os.path.join(dir_ckpt, name)
    torch.save({"model": inner.state_dict(), "opt": opt.state_dict(), "n_epoch": n_epoch, "n_iter": n_iter}, path)
    return path


def fit(args, model, opt, train_loader, val_loader, dir_ckpt, step_fn=None, val_fn=None, log=print, is_main=None):
    """The epoch loop of train.py:136-183 (and train_gt.py's, with ``step_fn=train_step_gt, val_fn=val_step_gt``):
    optional resume from the latest checkpoint (``args.resume``), ``train_step`` per batch with a log line every
    ``args.freq_log`` iterations, validation + checkpoint every ``args.freq_ckpt`` epochs, learning rate times
    ``args.weight_decay`` every ``args.freq_decay`` epochs (the reference's naming: it is an lr decay factor).
    Under torch.distributed only rank 0 writes checkpoints (``is_main``); every rank resumes from the same file.
    Returns (n_epoch, n_iter)."""
    step_fn = step_fn or train_step
    val_fn = val_fn or val_step
    if is_main is None:
        is_main = not (torch.distributed.is_available() and torch.distributed.is_initialized()) or \
            torch.distributed.get_rank() == 0
    inner = model.module if hasattr(model, "module") else model
    n_epoch = n_iter = 0
    if getattr(args, "resume", False):
        path = latest_checkpoint(dir_ckpt)
        if path is None:
            raise FileNotFoundError(f"--resume: no checkpoint in {dir_ckpt}")
        ckpt = torch.load(path, map_location=next(inner.parameters()).device)
        inner.load_state_dict(ckpt["model"])
        opt.load_state_dict(ckpt["opt"])
        n_epoch, n_iter = ckpt["n_epoch"] + 1, ckpt["n_iter"]
    for _ in range(n_epoch, args.n_epochs):
        model.train()
        for batch in train_loader:
            out = step_fn(batch, model, opt, args)
            if n_iter % args.freq_log == 0:
                log("[train] epoch:", n_epoch, ", iter:", n_iter, " losses / acc:", out)
            n_iter += 1
        if n_epoch % args.freq_ckpt == 0:
            model.eval()
            metrics = val_fn(model, val_loader, getattr(args, "pred_type", "sdf"))
            log("[val] epoch:", n_epoch, ", iter:", n_iter, " metrics:", metrics)
            if is_main:
                save_checkpoint(dir_ckpt, model, opt, n_epoch, n_iter, [m for m in metrics if m is not None])
        if n_epoch > 0 and n_epoch % args.freq_decay == 0:
            for g in opt.param_groups:
                g["lr"] = g["lr"] * args.weight_decay
        n_epoch += 1
    return n_epoch, n_iter


def wrap_ddp(model, device=None):
    """One process per GPU (torchrun): DistributedDataParallel in place of train.py:131-132's DataParallel."""
    from torch.nn.parallel import DistributedDataParallel as DDP
    if device is not None and torch.device(device).type == "cuda":
        return DDP(model, device_ids=[torch.device(device).index], find_unused_parameters=True)
    return DDP(model, find_unused_parameters=True)
