"""Marching cubes (slice3d_b200/mcubes.py, Generator3D.extract_mesh) against golden vectors made with the reference's
own marching cubes core (oracle/make_golden_mcubes.py) and, when oracle/_ref holds it, the compiled reference directly.

Parity bar: BOTH arrays bit for bit -- the vertex array (float64 values AND order) and the face array (int64 indices,
triangle order inside a cell and vertex order inside a triangle included)."""
import os

import numpy as np
import pytest
import torch

from slice3d_b200 import mcubes
from tests import helpers, mc_volumes


def _polygons(tris, cell_of_tri):
    """Set of canonical oriented polygon loops: inside one cell, opposite directed edges (the diagonals) cancel."""
    out, start, n = set(), 0, len(tris)
    while start < n:
        end = start
        while end < n and cell_of_tri[end] == cell_of_tri[start]:
            end += 1
        edges = {}
        for t in tris[start:end]:
            for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                if (b, a) in edges:
                    del edges[(b, a)]
                else:
                    edges[(a, b)] = 1
        nxt = dict(edges.keys())
        seen = set()
        for s in list(nxt):
            if s in seen:
                continue
            loop, cur = [s], nxt[s]
            seen.add(s)
            while cur != s:
                loop.append(cur)
                seen.add(cur)
                cur = nxt[cur]
            k = loop.index(min(loop))
            out.add(tuple(int(x) for x in loop[k:] + loop[:k]))
        start = end
    return out


def _check(vol, iso, ref_v, ref_t, device="cpu"):
    v, t = mcubes.marching_cubes(torch.from_numpy(vol).to(device), iso)
    v, t = v.cpu().numpy(), t.cpu().numpy()
    assert v.dtype == np.float64 and v.shape == ref_v.shape and np.array_equal(v, ref_v)  # values and order
    assert t.dtype == np.int64 and np.array_equal(t, ref_t.astype(np.int64))


@pytest.mark.parametrize("name", list(mc_volumes.cases()))
def test_marching_cubes_matches_reference_golden(name):
    vol, iso = mc_volumes.cases()[name]
    gold = helpers.load_case(f"mcubes_{name}")
    _check(vol, iso, gold["vertices"], gold["triangles"])


def test_marching_cubes_matches_compiled_reference_random():
    from oracle import build_ref_mcubes
    if build_ref_mcubes.build() is None:
        pytest.skip("/root/reference not available: oracle/_ref/libref_mcubes.so cannot be built")
    rng = np.random.RandomState(11)
    for shape in [(6, 5, 9), (2, 2, 2), (12, 12, 12), (3, 17, 4)]:
        vol = rng.randn(*shape)
        rv, rt = build_ref_mcubes.run(vol, 0.05)
        _check(vol, 0.05, rv, rt.astype(np.int64))


def test_table_constant_agrees_with_cube_geometry():
    """mc_table.TRI_ROWS (the reference's constant) against the table derived from the cube's geometry: for every one
    of the 256 configurations the same oriented polygons, so a corrupted constant cannot pass; plus closure checks."""
    gen_table, gen_count = mcubes._build_tables()
    for c in range(256):
        tris = [tuple(t) for t in mcubes._TABLE[c] if t[0] >= 0]
        gen = np.array([t for t in gen_table[c] if t[0] >= 0], dtype=np.int64).reshape(-1, 3)
        assert len(gen) == gen_count[c] == len(tris)
        assert _polygons(np.array(tris, dtype=np.int64).reshape(-1, 3), np.zeros(len(tris), dtype=np.int64)) == \
            _polygons(gen, np.zeros(len(gen), dtype=np.int64))
        assert len(tris) == mcubes._COUNT[c] <= 5
        crossed = {e for e, (a, b) in enumerate(mcubes._EDGES) if ((c >> a) & 1) != ((c >> b) & 1)}
        assert {e for t in tris for e in t} == crossed
        loops = _polygons(np.array(tris, dtype=np.int64).reshape(-1, 3), np.zeros(len(tris), dtype=np.int64))
        assert sum(len(l) for l in loops) == len(crossed)
        # complementary configurations cut the same polygons with the opposite orientation only when no face is ambiguous;
        # triangle counts follow polygon sizes
        assert sum(len(l) - 2 for l in loops) == len(tris)


def test_extract_mesh_transform_and_export(tmp_path):
    """Generator3D.extract_mesh (reconstruct.py:175-223): padding, cell-centre shift, unit-box normalisation."""
    from slice3d_b200 import Generator3D
    vol, iso = mc_volumes.cases()["blob_17"]
    gold = helpers.load_case("mcubes_extract_blob_17")
    gen = Generator3D(None, threshold=0.5, upsampling_steps=0, pred_type="sdf")
    stats = {}
    mesh = gen.extract_mesh(vol, stats_dict=stats)
    assert np.array_equal(mesh.vertices, gold["vertices"]) and np.array_equal(mesh.faces, gold["triangles"])
    assert stats["n_vertices"] == len(mesh.vertices)
    assert stats["time (marching cubes)"] >= 0.0  # the reference's stats key (reconstruct.py:193)
    # closed surface: every undirected edge is shared by exactly two faces, with opposite directions
    d = {}
    for t in mesh.faces:
        for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
            d[(a, b)] = d.get((a, b), 0) + 1
    assert all(n == 1 and d.get((b, a), 0) == 1 for (a, b), n in d.items())
    for ext in ("obj", "off", "ply"):
        p = mesh.export(os.path.join(tmp_path, "m." + ext))
        assert os.path.getsize(p) > 1000
    # generate_mesh's plumbing (reconstruct.py:104-173) with the value grid supplied: same mesh, the reference's stats keys
    gen.generate_grid = lambda data, as_numpy=True: np.asarray(vol)
    mesh2, stats2 = gen.generate_mesh({})
    assert np.array_equal(mesh2.vertices, mesh.vertices) and np.array_equal(mesh2.faces, mesh.faces)
    assert {"time (eval points)", "time (marching cubes)", "n_vertices", "n_faces"} <= set(stats2)
    assert gen.generate_mesh({}, return_stats=False).faces.shape == mesh.faces.shape
    with pytest.raises(NotImplementedError):
        Generator3D(None, with_normals=True, pred_type="sdf").extract_mesh(vol)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_marching_cubes_on_device_and_generate_mesh():
    for name in ("noise_9x7x6", "border_7"):
        vol, iso = mc_volumes.cases()[name]
        gold = helpers.load_case(f"mcubes_{name}")
        _check(vol, iso, gold["vertices"], gold["triangles"], device="cuda:0")
    # end to end: encoder -> MISE-refined value grid -> mesh, the whole reconstruct.py generate_mesh path
    from slice3d_b200 import Generator3D
    case = helpers.load_case("k12_s128_g128")
    m, sd = helpers.case_weights(case)
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda:0").eval()
    gen = Generator3D(m, resolution0=8, upsampling_steps=2, pred_type="sdf")
    with torch.no_grad():
        mesh, stats = gen.generate_mesh(helpers.case_feed(case))
        grid = gen.generate_sparse_grid(helpers.case_feed(case))
    # the same mesh as marching cubes on the CPU over the same value grid
    ref = Generator3D(None, resolution0=8, upsampling_steps=2, pred_type="sdf").extract_mesh(grid)
    assert np.array_equal(mesh.vertices, ref.vertices) and np.array_equal(mesh.faces, ref.faces)
    assert len(mesh.faces) > 0 and np.abs(mesh.vertices).max() <= 0.5 + 1.5 / 32
    print(f"generate_mesh: {len(mesh.vertices)} vertices, {len(mesh.faces)} faces")
