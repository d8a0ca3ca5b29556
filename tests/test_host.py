"""CPU-side checks: checkpoint layout, C-ABI exports, autograd (training) arithmetic against
the reference golden, grid/slab logic, and the world_size-2 gloo path of the slab all-gather."""
import ctypes
import json
import os
import re
import subprocess
import sys

import pytest
import torch

from slice3d_b200 import Slices3DRegModel, _native, dist as s3d_dist, synth
from slice3d_b200.generator import Generator3D
from tests import helpers

ROOT = helpers.ROOT


def test_state_dict_layout_matches_reference_manifest():
    """244 keys, same order/shape/dtype as the reference module (manifest generated from
    the reference in the build container: SURVEY.md section 8b)."""
    manifest = json.load(open(os.path.join(helpers.GOLDEN, "state_dict_manifest.json")))
    sd = Slices3DRegModel(128, 12, "test").state_dict()
    assert list(sd.keys()) == list(manifest.keys())
    assert len(sd) == 244
    for k, (shape, dtype) in manifest.items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dtype, k


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "slice3d_b200.h")).read()
    declared = set(re.findall(r"\b(s3d_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.SYMBOLS)
    lib = ctypes.CDLL(_native.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s
    lib.s3d_abi_version.restype = ctypes.c_int
    assert lib.s3d_abi_version() == _native.ABI_VERSION


def test_inference_without_cuda_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m = Slices3DRegModel(128, 12, "test").eval()
    feed = synth.synthetic_inputs(128)
    feed["qry_norot"] = torch.zeros(1, 8, 3)
    with torch.no_grad(), pytest.raises(_native.NativeError):
        m(feed)


def test_autograd_path_matches_reference_val_golden():
    """Training arithmetic (torch ops) in eval mode == reference forward (rotation branch)."""
    case = helpers.load_case("k12_s128_val_rot")
    m, sd = helpers.case_weights(case)
    m.load_state_dict(sd, strict=True)
    m.eval()
    feed = helpers.case_feed(case, batch=2)
    feed["qry_norot"] = torch.from_numpy(case["qry"]).clone()
    feed["obj_rot_mat"] = torch.from_numpy(case["obj_rot_mat"])
    with torch.no_grad():
        ret = m._forward_autograd(feed)
    assert helpers.maxabs(ret["sdf_pred"], case["sdf"]) < 2e-5
    assert helpers.maxabs(ret["vgg_loss"], case["vgg_loss"]) < 1e-6
    assert helpers.maxabs(ret["slices_rec"][:, :, ::8, ::8], case["slices_rec_sub_b"]) < 2e-5


def test_train_step_runs_and_unused_params_get_no_grad():
    """reg_slices/train.py:41-53 semantics on a tiny batch: loss = L1(sdf)+L1(img)+vgg."""
    torch.manual_seed(0)
    m = Slices3DRegModel(32, 12, "train").train()
    feed = synth.synthetic_inputs(32, batch=2)
    feed["qry_norot"] = torch.rand(2, 16, 3) - 0.5
    feed["sdf"] = torch.randn(2, 16) * 0.1
    opt = torch.optim.Adam(m.parameters(), lr=3e-4)
    x = m(feed)
    loss = (torch.nn.functional.l1_loss(x["sdf_pred"], feed["sdf"]) +
            torch.nn.functional.l1_loss(x["slices_rec"], feed["img_slices"]) + x["vgg_loss"])
    loss.backward()
    opt.step()
    assert torch.isfinite(loss)
    no_grad = [n for n, p in m.named_parameters() if p.requires_grad and p.grad is None]
    # att_layer.* (12 tensors) and down5_.41.{weight,bias} never get a gradient
    assert len(no_grad) == 14 and all(n.startswith(("att_layer.", "slices_generator.down5_.")) for n in no_grad)


def test_make_3d_grid_order_and_axes():
    g = synth.make_3d_grid((-0.5,) * 3, (0.5,) * 3, (3, 4, 5))
    assert g.shape == (60, 3)
    ax = torch.linspace(-0.5, 0.5, 5)
    assert torch.equal(g[:5, 2], ax) and torch.all(g[:5, 0] == -0.5)  # z fastest
    gen = Generator3D(model=None, upsampling_steps=0, resolution0=5)
    assert torch.equal(gen.grid_axes(5, "cpu"), ax)


@pytest.mark.parametrize("nx,world", [(256, 8), (128, 2), (65, 4), (7, 8), (1, 2)])
def test_slab_bounds_partition(nx, world):
    b = s3d_dist.slab_bounds(nx, world)
    assert b[0] == 0 and b[-1] == nx and len(b) == world + 1
    sizes = [b[i + 1] - b[i] for i in range(world)]
    assert all(s >= 0 for s in sizes) and max(sizes) - min(sizes) <= 1
    assert [s3d_dist.slab_range(nx, r, world) for r in range(world)] == [(b[r], b[r + 1]) for r in range(world)]


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from slice3d_b200 import dist as sd
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
ok = True
for nx in (8, 7):
    full = torch.arange(nx * 3 * 2, dtype=torch.float32)
    lo, hi = sd.slab_range(nx, dist.get_rank(), 2)
    vol = torch.full_like(full, -1.0)
    vol[lo * 6:hi * 6] = full[lo * 6:hi * 6]
    sd.all_gather_slabs(vol, nx)
    ok = ok and torch.equal(vol, full)
# rate-balanced (ragged) slabs: rank 1 is 3x as fast as rank 0
b = sd.proportional_bounds(8, [1.0, 3.0])
ok = ok and b == [0, 2, 8]
full = torch.arange(8 * 6, dtype=torch.float32)
vol = torch.full_like(full, -1.0)
r = dist.get_rank()
vol[b[r] * 6:b[r + 1] * 6] = full[b[r] * 6:b[r + 1] * 6]
sd.all_gather_slabs(vol, 8, bounds=b)
ok = ok and torch.equal(vol, full)
# row-granular slabs (Generator3D.slab_unit = "row"): boundaries in rows of nz values, a slab ends inside a plane
nx = 6
full = torch.arange(nx ** 3, dtype=torch.float32)
for rates in ([1.0, 1.3], [2.0, 1.0], [1.0, 1.0]):
    b = sd.proportional_bounds(nx * nx, rates)
    vol = torch.full_like(full, -1.0)
    for first, count, whole in sd.split_at_planes(b[r] * nx, b[r + 1] * nx, nx * nx):
        vol[first:first + count] = full[first:first + count]
    sd.all_gather_slabs(vol, nx, bounds=b, units=nx * nx)
    ok = ok and torch.equal(vol, full)
dist.destroy_process_group()
sys.exit(0 if ok else 3)
"""


def test_proportional_bounds():
    assert s3d_dist.proportional_bounds(256, [1.0] * 8) == s3d_dist.slab_bounds(256, 8)
    b = s3d_dist.proportional_bounds(256, [1, 1, 1, 1, 1, 1, 1, 0.955])
    assert b[0] == 0 and b[-1] == 256 and b[8] - b[7] == 31 and all(b[i + 1] - b[i] in (32, 33) for i in range(7))
    assert s3d_dist.proportional_bounds(3, [1, 100, 1]) == [0, 1, 2, 3]  # every rank keeps a plane
    assert s3d_dist.proportional_bounds(2, [1, 1, 1]) == s3d_dist.slab_bounds(2, 3)  # fewer planes than ranks


@pytest.mark.parametrize("plane", [1, 4, 36, 65536])
def test_split_at_planes_covers_the_range_once(plane):
    """Row-granular slabs: a rank's query range becomes <= 3 launch segments -- rows before the first whole plane, whole
    planes (marked: decoded in locality order), rows after -- that tile the range exactly, in order."""
    import random
    rnd = random.Random(plane)
    total = plane * 9
    cases = [(0, total), (0, 0), (5 % total, 5 % total), (plane, 2 * plane), (1 % total, total)]
    cases += [tuple(sorted((rnd.randrange(total + 1), rnd.randrange(total + 1)))) for _ in range(200)]
    for q0, q1 in cases:
        segs = s3d_dist.split_at_planes(q0, q1, plane)
        assert len(segs) <= 3
        pos = q0
        for first, count, whole in segs:
            assert first == pos and count > 0
            assert whole == (first % plane == 0 and count % plane == 0)
            pos += count
        assert pos == max(q0, q1) if q1 > q0 else segs == []
        if q1 > q0:
            assert sum(w for _, _, w in segs) <= 1  # one launch of whole planes at most
            inner = (q1 // plane) - (-(-q0 // plane))  # whole planes inside the range
            assert sum(c for _, c, w in segs if w) == max(inner, 0) * plane


def test_generator_slab_units_and_row_plan():
    from slice3d_b200 import Generator3D
    g = Generator3D(model=None, upsampling_steps=0, resolution0=256)
    assert g._slab_units(256, 1) == 256 and g._slab_units(256, 2) == 256  # 128 planes per rank: planes are fine enough
    assert g._slab_units(256, 4) == 256                                   # 64 planes per rank
    assert g._slab_units(256, 8) == 256 * 256                             # 32 planes per rank: one plane is 3 % -> rows
    g.slab_unit = "row"
    assert g._slab_units(256, 2) == 256 * 256
    g.slab_unit = "plane"
    assert g._slab_units(256, 8) == 256
    g.slab_unit = "bogus"
    with pytest.raises(ValueError):
        g._slab_units(256, 8)
    # rates 1 % apart at 8 ranks: plane slabs can only answer with 31 / 32 / 33 planes (3 % steps), row slabs follow them
    rates = [1.0, 1.01, 0.99, 1.0, 1.005, 0.995, 1.0, 1.0]
    rows = s3d_dist.proportional_bounds(256 * 256, rates)
    t_rows = max((rows[i + 1] - rows[i]) / rates[i] for i in range(8))
    planes = s3d_dist.proportional_bounds(256, rates)
    t_planes = max((planes[i + 1] - planes[i]) * 256 / rates[i] for i in range(8))
    ideal = 256 * 256 / sum(rates)
    assert t_rows / ideal < 1.0002 and t_planes / ideal > 1.005


def test_slab_plan_averages_the_last_four_calls(monkeypatch):
    """Generator3D._slab_plan: this rank's rate is planes / ms summed over its last four launches at the same size (a
    single call scatters by ~0.4 %); a new size starts a new history.  Events and the all-gather are faked (world 2, the
    other rank always reports 128 planes in 640 ms per call)."""
    import torch.distributed as dist
    from slice3d_b200 import Generator3D

    class Ev:
        def __init__(self, t):
            self.t = t

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return other.t - self.t

    g = Generator3D(model=None, upsampling_steps=0, resolution0=256)
    g.slab_unit = "row"

    def fake_all_gather(out, inp, group=None):
        n = len(g._rate_hist)
        out[0] = inp
        out[1] = torch.tensor([128.0 * n, 640.0 * n], dtype=torch.float64)

    monkeypatch.setattr(dist, "all_gather_into_tensor", fake_all_gather)
    b, units = g._slab_plan(256, 0, 2, None, "cpu")
    assert units == 65536 and b == [0, 32768, 65536]  # no history: equal plane-aligned slabs, in rows
    g._last_dec = (Ev(0.0), Ev(630.0), 128.0, 256, 2)  # this rank was 1.6 % faster than the other
    b, _ = g._slab_plan(256, 0, 2, None, "cpu")
    assert b[1] == 33026  # round(65536 * (128/630) / (128/630 + 128/640))
    g._last_dec = (Ev(0.0), Ev(650.0), 128.0, 256, 2)  # ... then 1.6 % slower: the two calls average out
    b, _ = g._slab_plan(256, 0, 2, None, "cpu")
    assert b[1] == 32768 and len(g._rate_hist) == 2
    for _ in range(5):
        g._last_dec = (Ev(0.0), Ev(640.0), 128.0, 256, 2)
        b, _ = g._slab_plan(256, 0, 2, None, "cpu")
    assert len(g._rate_hist) == 4 and b[1] == 32768
    g._last_dec = (Ev(0.0), Ev(100.0), 64.0, 128, 2)  # another grid size: the history starts again
    b, units = g._slab_plan(128, 0, 2, None, "cpu")
    assert units == 128 * 128 and len(g._rate_hist) == 1 and b[0] == 0 and b[-1] == units


def test_slab_all_gather_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)]) for r in range(2)]
    codes = [p.wait(timeout=120) for p in procs]
    assert codes == [0, 0]


_DDP_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from slice3d_b200 import Slices3DRegModel, synth
rank = int(sys.argv[3])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=2)
torch.manual_seed(0)
torch.set_num_threads(2)
m = Slices3DRegModel(32, 12, "train").train()
for mod in m.modules():  # dropout off: the two runs below must see the same arithmetic
    if isinstance(mod, torch.nn.Dropout):
        mod.p = 0.0
    if isinstance(mod, torch.nn.MultiheadAttention):
        mod.dropout = 0.0
feed = synth.synthetic_inputs(32, batch=2)
g = torch.Generator().manual_seed(1)
feed["qry_norot"] = torch.rand(2, 16, 3, generator=g) - 0.5
feed["sdf"] = torch.randn(2, 16, generator=g) * 0.1
def loss_of(model, f):
    x = model(f)
    return (torch.nn.functional.l1_loss(x["sdf_pred"], f["sdf"]) +
            torch.nn.functional.l1_loss(x["slices_rec"], f["img_slices"]) + x["vgg_loss"])
mine = {k: v[rank:rank + 1].clone() for k, v in feed.items()}
# single-process gradients of each rank's sample, for the expected average
ref = {}
for r in range(2):
    m.zero_grad()
    loss_of(m, {k: v[r:r + 1].clone() for k, v in feed.items()}).backward()
    for n, p in m.named_parameters():
        if p.grad is not None:
            ref[n] = ref.get(n, 0) + p.grad.detach().clone() / 2
m.zero_grad()
ddp = torch.nn.parallel.DistributedDataParallel(m, find_unused_parameters=True)  # 14 parameters never get a gradient
loss_of(ddp, mine).backward()
worst = 0.0
for n, p in m.named_parameters():
    if p.grad is not None and n in ref:
        worst = max(worst, float((p.grad - ref[n]).abs().max() / (ref[n].abs().max() + 1e-12)))
dist.destroy_process_group()
sys.exit(0 if worst < 1e-4 else 3)
"""


def test_ddp_gradient_allreduce_world2_gloo(tmp_path):
    """BASELINE configs[4] on CPU: one process per replica, batch split across ranks, bucketed gradient all-reduce
    (DDP); per-replica BatchNorm statistics as in the reference's DataParallel (train.py:132).  The all-reduced
    gradients equal the average of the per-sample single-process gradients."""
    script = tmp_path / "ddp.py"
    script.write_text(_DDP_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)]) for r in range(2)]
    codes = [p.wait(timeout=600) for p in procs]
    assert codes == [0, 0]


def test_load_pretrained_vgg_maps_torchvision_keys():
    """the reference starts from torchvision's pretrained VGG16-BN / VGG19; this package takes their
    state_dicts explicitly (no download) and maps features.<i>.* into the trunk blocks / perceptual slices."""
    import warnings
    import torch
    from slice3d_b200 import Slices3DRegModel, synth
    torchvision = pytest.importorskip("torchvision")
    m = Slices3DRegModel(32, 12, "train")
    v16 = torchvision.models.vgg16_bn(weights=None).state_dict()
    v19 = torchvision.models.vgg19(weights=None).state_dict()
    assert m.load_pretrained_vgg(v16, v19) == 13 * 2 + 13 * 5 + 14 * 2
    sd = m.state_dict()
    assert torch.equal(sd["slices_generator.down1.0.weight"], v16["features.0.weight"])
    assert torch.equal(sd["slices_generator.down5_.41.running_var"], v16["features.41.running_var"])
    assert torch.equal(sd["vggptlossfunc.vgg.slice5.30.bias"], v19["features.30.bias"])
    fresh = Slices3DRegModel(32, 12, "train").train()
    feed = synth.synthetic_train_batch(32, 12, batch=1, n_qry=8, seed=0)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        fresh(feed)
    assert any("pretrained" in str(x.message) for x in w)


def test_auto_precision_selection_policy():
    """precision='auto' (the package default): fp16f8 only when the probe against the fp32 path stays within AUTO_TOL,
    otherwise fp16x3 with a warning; resolved once per packed-weight handle; explicit modes pass through untouched.  (The
    probe itself needs a GPU: tests/test_gpu_parity.py::test_decoder_error_margin_under_weight_scale_stress.)"""
    import warnings

    def handle():
        nm = object.__new__(_native.NativeModel)  # no CUDA handle: only the selection logic is exercised
        nm._auto, nm.auto_info, nm._h = None, None, None
        return nm

    calls = []

    def probe(err):
        def run(prec):
            calls.append(prec)
            return torch.zeros(64) if prec == "fp32" else torch.full((64,), err)
        return run

    nm = handle()
    assert nm.resolve_precision("fp16x3") == "fp16x3" and nm.resolve_precision("fp32", probe(1.0)) == "fp32" and not calls
    with pytest.raises(_native.NativeError):
        nm.resolve_precision("auto")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert nm.resolve_precision("auto", probe(0.5 * _native.AUTO_TOL)) == "fp16f8" and not w
    assert calls == ["fp32", "fp16f8"] and nm.auto_info["selected"] == "fp16f8"
    assert nm.resolve_precision("auto", probe(1.0)) == "fp16f8" and len(calls) == 2  # resolved once per handle
    nm = handle()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert nm.resolve_precision("auto", probe(2.0 * _native.AUTO_TOL)) == "fp16x3"
    assert any("fp16x3" in str(x.message) for x in w) and nm.auto_info["fp16f8_max_abs_vs_fp32"] > _native.AUTO_TOL
    nm = handle()
    with warnings.catch_warnings(record=True):
        warnings.simplefilter("always")
        assert nm.resolve_precision("auto", probe(float("nan"))) == "fp16x3"  # a NaN probe never selects the fast mode
    assert Slices3DRegModel(64, 12, "test").precision == "auto"


def test_fit_loop_checkpoints_resume_and_lr_decay(tmp_path):
    """The epoch loop of train.py:136-183: checkpoint dict and file name, resume from the latest file, lr decay schedule."""
    import time
    from types import SimpleNamespace
    from slice3d_b200 import train as T

    def make():
        torch.manual_seed(0)
        m = torch.nn.Linear(3, 1)
        return m, torch.optim.Adam(m.parameters(), lr=1e-2)

    def step(batch, model, opt, args):
        opt.zero_grad()
        loss = (model(batch["x"]) - batch["y"]).abs().mean()
        loss.backward()
        opt.step()
        return loss.item(), 0.5

    def val(model, loader, pred_type):
        time.sleep(0.02)  # distinct creation times for latest_checkpoint
        return 0.123456, 0.98765, torch.tensor(0.4567891)

    g = torch.Generator().manual_seed(1)
    loader = [{"x": torch.randn(4, 3, generator=g), "y": torch.randn(4, 1, generator=g)} for _ in range(5)]
    args = SimpleNamespace(n_epochs=5, freq_log=2, freq_ckpt=2, freq_decay=2, weight_decay=0.5, resume=False, pred_type="sdf")
    d = str(tmp_path / "ckpt")
    logs = []
    m, opt = make()
    assert T.fit(args, m, opt, loader, loader, d, step, val, log=lambda *a: logs.append(a)) == (5, 25)
    # validation + checkpoint at epochs 0, 2, 4; the reference's file name (train.py:169)
    assert sorted(os.listdir(d)) == ["0_5_0.1235_0.9877_0.4568.ckpt", "2_15_0.1235_0.9877_0.4568.ckpt",
                                     "4_25_0.1235_0.9877_0.4568.ckpt"]
    ck = torch.load(T.latest_checkpoint(d))
    assert set(ck) == {"model", "opt", "n_epoch", "n_iter"} and (ck["n_epoch"], ck["n_iter"]) == (4, 25)
    assert all(torch.equal(ck["model"][k], v) for k, v in m.state_dict().items())
    # lr halves after epochs 2 and 4 (n_epoch > 0 and n_epoch % freq_decay == 0); the checkpoint of epoch 4 was written before
    assert opt.param_groups[0]["lr"] == pytest.approx(1e-2 * 0.25)
    assert ck["opt"]["param_groups"][0]["lr"] == pytest.approx(1e-2 * 0.5)
    assert sum(1 for l in logs if l[0] == "[train] epoch:") == 13 and sum(1 for l in logs if l[0] == "[val] epoch:") == 3
    # resume: weights, Adam state and counters continue from the latest file (epoch 5 onwards)
    m2, opt2 = make()
    args2 = SimpleNamespace(**{**vars(args), "resume": True, "n_epochs": 7})
    assert T.fit(args2, m2, opt2, loader, loader, d, step, val, log=lambda *a: None) == (7, 35)
    assert "6_35_0.1235_0.9877_0.4568.ckpt" in os.listdir(d)
    m3, opt3 = make()  # the same 7 epochs without interruption
    d3 = str(tmp_path / "ckpt3")
    T.fit(SimpleNamespace(**{**vars(args), "n_epochs": 7}), m3, opt3, loader, loader, d3, step, val, log=lambda *a: None)
    # (the resumed run restarts from the epoch-4 optimizer state: lr 0.5e-2 at epoch 5, as the reference's resume does,
    #  while the uninterrupted run had already decayed to 0.25e-2 -- the reference's own resume semantics, kept)
    assert opt2.param_groups[0]["lr"] == pytest.approx(1e-2 * 0.5 * 0.5) and opt3.param_groups[0]["lr"] == pytest.approx(1e-2 * 0.125)
    with pytest.raises(FileNotFoundError):
        T.fit(args2, m2, opt2, loader, loader, str(tmp_path / "none"), step, val)
    # a DDP / DataParallel wrapper is unwrapped for the checkpoint
    wrapped = SimpleNamespace(module=m)
    p = T.save_checkpoint(str(tmp_path / "w"), wrapped, opt, 1, 2, [0.5])
    assert os.path.basename(p) == "1_2_0.5.ckpt" and set(torch.load(p)["model"]) == set(m.state_dict())


def test_abi_size_queries_and_argument_errors_need_no_gpu():
    """Host-only entry points of include/slice3d_b200.h: the size queries a caller allocates from, and the argument checks
    that run before any CUDA call (negative S3D_ERR_* + a message from s3d_last_error, nothing thrown, nothing touched)."""
    L = _native.lib()
    # projected planes: per image (K, R_s, R_s, 128) fp32 for R_s = S/16 * 2^s, s = 0..4
    for B, K, S in [(1, 12, 256), (2, 4, 128), (3, 1, 64)]:
        assert L.s3d_planes_bytes(B, K, S) == B * K * 128 * 4 * sum((S // 16 * 2 ** s) ** 2 for s in range(5))
    assert L.s3d_planes_bytes(1, 12, 256) == 536346624
    # workspaces: positive, non-decreasing in the problem size, the decoder's bounded by its persistent grid
    assert 0 < L.s3d_encoder_workspace_bytes(1, 12, 128) < L.s3d_encoder_workspace_bytes(1, 12, 256) \
        <= L.s3d_encoder_workspace_bytes(2, 12, 256)
    for prec in _native.PRECISIONS.values():
        a, b = L.s3d_decoder_workspace_bytes(3000, prec), L.s3d_decoder_workspace_bytes(256 ** 3, prec)
        assert 0 < a <= b < (1 << 33)
    assert L.s3d_decoder_workspace_bytes(256 ** 3, _native.PREC_FP16F8) == L.s3d_decoder_workspace_bytes(128 ** 3, _native.PREC_FP16F8)
    assert L.s3d_sparse_scratch_bytes(32, 3, 1 << 20) > 12 * (1 << 20)  # at least the compacted points of one round
    assert L.s3d_mise_scratch_ints(32, 3) >= (32 * 8 + 1) ** 3 // 8
    assert L.s3d_scan_scratch_bytes(10 ** 6) > 0 and L.s3d_vgg_loss_workspace_bytes(12, 128) > 0
    assert L.s3d_preprocess_workspace_bytes(13, 137, 128) >= 13 * 137 * 128 * 3  # the 8-bit intermediate of the two passes
    # argument errors
    h = ctypes.c_void_p()
    assert L.s3d_model_create(ctypes.byref(h), None, 0, 12, 0, None) == -1 and not h.value  # S3D_ERR_BAD_ARG
    assert b"model_create" in L.s3d_last_error()
    assert L.s3d_decoder_fwd(None, None, 256, None, 10, None, None, 0, 1.0, None, 3, None, 0, None) < 0
    assert b"decoder" in L.s3d_last_error()
    assert L.s3d_encoder_fwd(None, None, 1, 256, None, None, None, None, 0, None) < 0
    assert b"encoder" in L.s3d_last_error()
    assert L.s3d_model_n_slices(None) == 0
    L.s3d_model_destroy(None)  # destroying nothing is a no-op
