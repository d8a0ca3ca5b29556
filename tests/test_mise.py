"""MISE octree refinement (slice3d_b200/mise.py, Generator3D.generate_sparse_grid) against golden vectors made with the
reference's own MISE class (oracle/make_golden_mise.py) and, when oracle/_ref holds the compiled reference, against it
directly.  Integer bookkeeping: point sets and the dense volume are compared exactly."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

from slice3d_b200.mise import MISE
from tests import helpers, mise_fields

REF_DIR = os.path.join(helpers.ROOT, "oracle", "_ref")


def _ref_mise():
    if not glob.glob(os.path.join(REF_DIR, "mise*.so")):
        return None
    sys.path.insert(0, REF_DIR)
    try:
        from mise import MISE as RefMISE
        return RefMISE
    except ImportError:
        return None
    finally:
        sys.path.remove(REF_DIR)


def _run(m, field, res):
    rounds = []
    p = m.query()
    while p.shape[0] != 0:
        pn = p.cpu().numpy() if torch.is_tensor(p) else p
        rounds.append(pn.shape[0])
        v = field(pn, res)
        m.update(p, torch.from_numpy(v) if torch.is_tensor(p) else v)
        p = m.query()
    d = m.to_dense()
    return rounds, (d.cpu().numpy() if torch.is_tensor(d) else d)


@pytest.mark.parametrize("name", list(mise_fields.CASES))
def test_mise_matches_reference_golden(name):
    res0, depth, thr, field = mise_fields.CASES[name]
    gold = helpers.load_case(f"mise_{name}")
    rounds, dense = _run(MISE(res0, depth, thr), field, res0 << depth)
    assert rounds == gold["rounds"].tolist()
    assert dense.dtype == np.float64 and np.array_equal(dense, gold["dense"])


def test_mise_known_answer_of_the_reference_test():
    """libmise/test.py: MISE(1, 2, 0.) with v = 2 * (x+y+z > 2) - 1 -> 3 rounds, 5^3 dense, sum 105.0."""
    res0, depth, thr, field = mise_fields.CASES["known_answer_r1_d2"]
    rounds, dense = _run(MISE(res0, depth, thr), field, 4)
    assert rounds == [8, 19, 61] and dense.shape == (5, 5, 5) and dense.sum() == 105.0


@pytest.mark.parametrize("name", ["sphere_r8_d2", "rough_r8_d3"])
def test_mise_matches_compiled_reference(name):
    RefMISE = _ref_mise()
    if RefMISE is None:
        pytest.skip("oracle/_ref/mise*.so not built (python oracle/build_ref_mise.py)")
    res0, depth, thr, field = mise_fields.CASES[name]
    ref, mine = RefMISE(res0, depth, thr), MISE(res0, depth, thr)
    res = res0 << depth
    p_ref, p = ref.query(), mine.query()
    while p_ref.shape[0] != 0:
        pn = p.numpy()
        # same point SET every round (the order differs: insertion order there, flat-index order here)
        key = lambda a: np.sort((a[:, 0] * (res + 1) + a[:, 1]) * (res + 1) + a[:, 2])
        assert np.array_equal(key(p_ref), key(pn))
        ref.update(p_ref, field(p_ref, res))
        mine.update(p, torch.from_numpy(field(pn, res)))
        p_ref, p = ref.query(), mine.query()
    assert p.shape[0] == 0
    assert np.array_equal(ref.to_dense(), mine.to_dense().numpy())


def test_mise_partial_state_and_errors():
    m = MISE(2, 1, 0.0)
    p = m.query()
    assert p.shape == (27, 3) and p.dtype == torch.int64
    with pytest.raises(ValueError):
        m.update(torch.tensor([[1, 0, 0]]), torch.tensor([1.0]))  # not a grid point yet
    with pytest.raises(ValueError):
        m.update(p, torch.zeros(3, dtype=torch.float64))
    # unknown points read as 0.0 in to_dense (mise.pyx:338-346: value=0., known=False) and the lattice is forward-filled
    m.update(p[:1], torch.tensor([5.0], dtype=torch.float64))
    d = m.to_dense().numpy()
    assert d.shape == (5, 5, 5) and d[0, 0, 0] == 5.0 and d[1, 0, 0] == 5.0 and d[2, 0, 0] == 0.0 and d[0, 0, 1] == 5.0
    # depth 0: nothing to refine, one round
    m0 = MISE(3, 0, 0.0)
    q = m0.query()
    m0.update(q, torch.arange(q.shape[0], dtype=torch.float64) - 30.0)
    assert m0.query().shape[0] == 0 and m0.to_dense().shape == (4, 4, 4)


def _gather_worker(rank, world, port, n, out):
    import torch.distributed as dist
    from slice3d_b200 import dist as s3d_dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = s3d_dist.slab_range(n, rank, world)
    flat = torch.full((n,), -1.0)
    flat[lo:hi] = torch.arange(lo, hi, dtype=torch.float32)
    s3d_dist.all_gather_ranges(flat, n)
    out[rank] = bool(torch.equal(flat, torch.arange(n, dtype=torch.float32)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7])
def test_point_shares_all_gather_world2(n):
    """The N > 1 path of generate_sparse_grid: contiguous shares of a round's points, even and ragged."""
    import torch.multiprocessing as mp
    port = 29600 + n
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gather_worker, args=(2, port, n, out), nprocs=2, join=True)
        assert out[0] and out[1]


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sphere_r4_d4", "rough_r8_d3"])
def test_mise_on_device_matches_golden(name):
    res0, depth, thr, field = mise_fields.CASES[name]
    gold = helpers.load_case(f"mise_{name}")
    rounds, dense = _run(MISE(res0, depth, thr, device="cuda:0"), field, res0 << depth)
    assert rounds == gold["rounds"].tolist()
    assert np.array_equal(dense, gold["dense"])


@pytest.mark.gpu
def test_generate_sparse_grid_matches_reference_loop():
    """Generator3D's MISE branch (reconstruct.py:147-167) on the CUDA path against the golden made from the
    reference MISE class + the CPU oracle.  Values within 1e-4; the octree may only differ where a value lies within
    the arithmetic tolerance of the threshold (a sign flip there changes which voxels are refined)."""
    from slice3d_b200 import Generator3D
    case = helpers.load_case("k12_s128_g128")
    gold = helpers.load_case("sparse_k12_s128_r8_d2")
    m, sd = helpers.case_weights(case)
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda:0").eval()
    feed = helpers.case_feed(case)
    gen = Generator3D(m, resolution0=int(gold["resolution0"]), upsampling_steps=int(gold["depth"]), pred_type="sdf")
    stats = {}
    with torch.no_grad():
        grid = gen.generate_sparse_grid(feed, precision="fp32", stats=stats)
        grid3 = gen.generate_sparse_grid(feed, precision="bf16x3")
    assert grid.shape == gold["dense"].shape and grid.dtype == np.float64
    diff = np.abs(grid - gold["dense"])
    print(f"sparse grid: points/round {stats['points_per_round']} (reference {gold['rounds'].tolist()}), "
          f"max-abs {diff.max():.3e}, bf16x3 vs fp32 {np.abs(grid3 - grid).max():.3e}")
    assert (diff < 1e-4).mean() > 0.999
    if stats["points_per_round"] == gold["rounds"].tolist():
        assert diff.max() < 1e-4
    # device-resident rounds (s3d_sparse_rounds: compaction + decoder with a device-side query count + update, several
    # rounds per host synchronisation) against the host-driven loop with the same decoder arithmetic: identical volume
    # and identical points per round
    with torch.no_grad():
        for rounds_per_sync in (6, 1, 40):
            gen.ROUNDS_PER_SYNC = rounds_per_sync
            s_dev, s_host = {}, {}
            gen.device_rounds = True
            g_dev = gen.generate_sparse_grid(feed, precision="fp16x3", stats=s_dev)
            gen.device_rounds = False
            g_host = gen.generate_sparse_grid(feed, precision="fp16x3", stats=s_host)
            assert s_dev["points_per_round"] == s_host["points_per_round"], rounds_per_sync
            assert np.array_equal(g_dev, g_host)
    gen.device_rounds = True
    # every lattice value of the sparse volume is the model's value at that point or a forward fill of one:
    # re-evaluate the whole lattice densely with the same kernels and compare at the evaluated points
    R = grid.shape[0] - 1
    idx = torch.nonzero(torch.ones(R + 1, R + 1, R + 1, dtype=torch.bool))
    pts = ((idx.double() / R - 0.5).float()).to("cuda:0")
    nat = m.native()
    planes = m.encode(feed["img_input"].to("cuda:0"))
    dense = nat.decode(planes, 0, pts.contiguous(), feed["trans_mat_wo_rot_tp"][0].to("cuda:0"), None, True, -1.0,
                       "fp32").double().view(R + 1, R + 1, R + 1).cpu().numpy()
    same = grid == dense
    assert same.mean() > 0.15  # the evaluated lattice points (the rest is forward fill)
    assert same[::4, ::4, ::4].all()  # the initial resolution0 lattice is always evaluated
