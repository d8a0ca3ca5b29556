"""Marching cubes on the device (torch tensor ops, vectorised over all cells), the step after the value grid
(reference: reg_slices/reconstruct.py:175-243 -> libmcubes.marching_cubes =
reg_slices/src_convonet/utils/libmcubes/marchingcubes.h:22-196, pywrapper.cpp:90-128).

Both output arrays equal libmcubes' bit for bit:

* the **vertex array**: same vertices, same float64 coordinates, same ORDER as the reference's sequential scan
  (cells x-major; inside a cell the three "owned" edges 6, 5, 10 first, then the edges that are only created on the
  low boundaries -- including the reference's duplicated vertices on the i = 0 / j = 0 / k = 0 faces).  The running
  vertex counter of the sequential code becomes an exclusive prefix sum over the per-cell counts.
* the **face array**: cells in scan order, and inside a cell the triangles of the classic Lorensen-Cline / Bourke
  table (`mc_table.TRI_ROWS`, the public-domain constant the reference consumes at marchingcubes.h:185-196) in the
  table's order, so `np.array_equal(faces, reference_faces)` holds.  `_build_tables` still derives the polygons of
  every configuration from the cube's geometry; tests/test_mcubes.py uses it to check that the constant cuts every
  configuration into the same oriented polygons (a typo in the constant cannot go unnoticed).

tests/test_mcubes.py compares both arrays exactly against the reference compiled from /root/reference
(oracle/build_ref_mcubes.py) and against committed goldens.
"""
import numpy as np
import torch

# corner m of a cell at (i, j, k) and the 12 edges (marchingcubes.h:60-64, 74-176)
_CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
_EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
_FACES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)]
# creation order of a cell's vertices and, per edge: (first corner, second corner, axis) of mc_add_vertex -- the vertex
# starts at the first corner and is interpolated towards the second along `axis` (marchingcubes.h:74-176)
_ORDER = [6, 5, 10, 0, 1, 2, 3, 4, 7, 8, 9, 11]
_ADD = {6: (6, 7, 0), 5: (5, 6, 1), 10: (2, 6, 2), 0: (0, 1, 0), 1: (1, 2, 1), 2: (2, 3, 0), 3: (3, 0, 1), 4: (4, 5, 0),
        7: (7, 4, 1), 8: (0, 4, 2), 9: (1, 5, 2), 11: (3, 7, 2)}
# edges not owned by the cell: (needs i > 0, j > 0, k > 0) to be shared, neighbour offset, neighbour's owned edge
_SHARED = {0: ((0, 1, 1), (0, -1, -1), 6), 1: ((0, 0, 1), (0, 0, -1), 5), 2: ((0, 0, 1), (0, 0, -1), 6),
           3: ((1, 0, 1), (-1, 0, -1), 5), 4: ((0, 1, 0), (0, -1, 0), 6), 7: ((1, 0, 0), (-1, 0, 0), 5),
           8: ((1, 1, 0), (-1, -1, 0), 10), 9: ((0, 1, 0), (0, -1, 0), 10), 11: ((1, 0, 0), (-1, 0, 0), 10)}


def _edge_id(a, b):
    for i, (p, q) in enumerate(_EDGES):
        if (p, q) == (a, b) or (p, q) == (b, a):
            return i
    raise KeyError((a, b))


def _build_tables():
    """(256, 5, 3) int64 triangle table (edge ids, -1 padded) and (256,) triangle counts, from the cube's geometry."""
    mid = [(np.array(_CORNERS[a]) + np.array(_CORNERS[b])) / 2.0 for a, b in _EDGES]
    table = -np.ones((256, 5, 3), dtype=np.int64)
    count = np.zeros(256, dtype=np.int64)
    for c in range(256):
        inside = [(c >> m) & 1 for m in range(8)]
        segs = []
        for f in _FACES:
            cr = [i for i in range(4) if inside[f[i]] != inside[f[(i + 1) % 4]]]
            if len(cr) == 2:
                segs.append((_edge_id(f[cr[0]], f[(cr[0] + 1) % 4]), _edge_id(f[cr[1]], f[(cr[1] + 1) % 4])))
            elif len(cr) == 4:  # ambiguous face: every inside corner is cut off on its own
                for i in range(4):
                    if inside[f[i]]:
                        segs.append((_edge_id(f[(i - 1) % 4], f[i]), _edge_id(f[i], f[(i + 1) % 4])))
        adj = {}
        for a, b in segs:
            adj.setdefault(a, []).append(b)
            adj.setdefault(b, []).append(a)
        seen, tris = set(), []
        for s in sorted(adj):
            if s in seen:
                continue
            loop, prev, cur = [s], None, s
            seen.add(s)
            while True:
                nxt = [n for n in adj[cur] if n != prev]
                nxt = nxt[0] if nxt else adj[cur][0]
                if nxt == s or nxt in seen:
                    break
                loop.append(nxt)
                seen.add(nxt)
                prev, cur = cur, nxt
            # orientation: normals point to the inside (value <= isovalue) corners, like the reference's table.  Vote over
            # the loop's vertices: local normal (next - cur) x (prev - cur) against the outside -> inside direction of
            # the cube edge the vertex sits on (the polygons are not planar, a global normal is not reliable).
            p = [mid[e] for e in loop]
            vote = 0.0
            for i, e in enumerate(loop):
                a, b = _EDGES[e]
                d = np.array(_CORNERS[a if inside[a] else b], dtype=float) - np.array(_CORNERS[b if inside[a] else a], dtype=float)
                n_loc = np.cross(p[(i + 1) % len(p)] - p[i], p[i - 1] - p[i])
                vote += float(np.dot(n_loc, d))
            if vote < 0:
                loop = [loop[0]] + loop[:0:-1]
            tris += [(loop[0], loop[i], loop[i + 1]) for i in range(1, len(loop) - 1)]
        count[c] = len(tris)
        for t, tri in enumerate(tris):
            table[c, t] = tri
    return table, count


def _canonical_tables():
    """(256, 5, 3) / (256,) arrays from the hex rows of mc_table.TRI_ROWS."""
    from .mc_table import TRI_ROWS
    table = -np.ones((256, 5, 3), dtype=np.int64)
    count = np.zeros(256, dtype=np.int64)
    for c, row in enumerate(TRI_ROWS):
        ids = [int(ch, 16) for ch in row]
        count[c] = len(ids) // 3
        table[c].reshape(-1)[:len(ids)] = ids
    return table, count


_TABLE, _COUNT = _canonical_tables()
_DEV_TABLES = {}


def marching_cubes(volume, isovalue):
    """volume (nx, ny, nz) -> (vertices (n, 3) float64, triangles (m, 3) int64) on the volume's device.
    Coordinates are the reference's (cell-centred: the reference shifts them by +0.5, reconstruct.py:194)."""
    vol = torch.as_tensor(volume).double()
    dev = vol.device
    nx, ny, nz = vol.shape
    cx, cy, cz = nx - 1, ny - 1, nz - 1
    if min(cx, cy, cz) <= 0:
        return torch.zeros((0, 3), dtype=torch.float64, device=dev), torch.zeros((0, 3), dtype=torch.int64, device=dev)
    if vol.is_cuda:
        # on the device: two hand-written kernels around the prefix sums (csrc/mcubes.cu, s3d_mc_count / s3d_mc_emit);
        # the tensor program below is the same algorithm for host tensors (what the CPU tests run)
        from . import _native
        key = str(dev)
        if key not in _DEV_TABLES:
            t15 = np.where(_TABLE.reshape(256, 15) < 0, 0, _TABLE.reshape(256, 15)).astype(np.int8)
            _DEV_TABLES[key] = (torch.as_tensor(t15, device=dev).contiguous(),
                                torch.as_tensor(_COUNT.astype(np.int32), device=dev).contiguous())
        tab, cnt = _DEV_TABLES[key]
        return _native.marching_cubes(vol.contiguous(), isovalue, tab, cnt)
    iso = float(isovalue)
    v = [vol[a:a + cx, b:b + cy, c:c + cz] for a, b, c in _CORNERS]  # corner values per cell
    inside = [x <= iso for x in v]
    cube = torch.zeros((cx, cy, cz), dtype=torch.int64, device=dev)
    for m in range(8):
        cube |= inside[m].long() << m
    crossed = {e: inside[a] != inside[b] for e, (a, b) in enumerate(_EDGES)}
    ii = torch.arange(cx, device=dev).view(-1, 1, 1)
    jj = torch.arange(cy, device=dev).view(1, -1, 1)
    kk = torch.arange(cz, device=dev).view(1, 1, -1)
    pos = (ii > 0, jj > 0, kk > 0)
    true = torch.ones((cx, cy, cz), dtype=torch.bool, device=dev)

    def shared_ok(e):  # the neighbour that owns this geometric edge exists
        need = _SHARED[e][0]
        ok = true
        for ax in range(3):
            if need[ax]:
                ok = ok & pos[ax]
        return ok

    new = {e: crossed[e] & (true if e in (6, 5, 10) else ~shared_ok(e)) for e in _ORDER}
    # running vertex counter of the sequential scan = exclusive prefix sum of the per-cell counts
    per_cell = sum(new[e].long() for e in _ORDER)
    flat = per_cell.reshape(-1)
    base = (torch.cumsum(flat, 0) - flat).view(cx, cy, cz)
    n_vert = int(flat.sum())
    idx, rank = {}, torch.zeros_like(base)
    for e in _ORDER:
        idx[e] = base + rank
        rank = rank + new[e].long()
    for e, (_, off, owner) in _SHARED.items():  # shared edges read the owning neighbour's index
        src = torch.roll(idx[owner], shifts=tuple(-o for o in off), dims=(0, 1, 2))
        idx[e] = torch.where(new[e], idx[e], src)
    # vertices (mc_add_vertex, marchingcubes.cpp:290-326): x = i + 0.5 etc., spacing exactly 1
    verts = torch.empty((n_vert, 3), dtype=torch.float64, device=dev)
    base_xyz = (ii.double() + 0.5, jj.double() + 0.5, kk.double() + 0.5)
    for e in _ORDER:
        a, b, axis = _ADD[e]
        m = new[e]
        if not bool(m.any()):
            continue
        f1, f2 = v[a][m], v[b][m]
        p1 = [(base_xyz[ax] + float(_CORNERS[a][ax])).expand(cx, cy, cz)[m] for ax in range(3)]
        x1 = p1[axis]
        x2 = (base_xyz[axis] + float(_CORNERS[b][axis])).expand(cx, cy, cz)[m]
        t = torch.where(f2 == f1, (x2 + x1) / 2, (x2 - x1) * (iso - f1) / (f2 - f1) + x1)
        out = torch.stack([t if ax == axis else p1[ax] for ax in range(3)], 1)
        verts[idx[e][m]] = out
    # triangles: cells in scan order, the table's order inside a cell
    table = torch.as_tensor(_TABLE, device=dev)
    count = torch.as_tensor(_COUNT, device=dev)[cube].reshape(-1)
    tbase = torch.cumsum(count, 0) - count
    n_tri = int(count.sum())
    tris = torch.empty((n_tri, 3), dtype=torch.int64, device=dev)
    idx_all = torch.stack([idx[e].reshape(-1) for e in range(12)], 1)  # (cells, 12)
    cube_f = cube.reshape(-1)
    for t in range(5):
        sel = torch.nonzero(count > t).squeeze(1)
        if sel.numel() == 0:
            break
        edges = table[cube_f[sel], t]  # (n, 3) edge ids
        tris[tbase[sel] + t] = torch.gather(idx_all[sel], 1, edges)
    return verts, tris


class Mesh(object):
    """Minimal stand-in for the trimesh.Trimesh the reference returns (reconstruct.py:221-243): vertices, faces, export."""

    def __init__(self, vertices, faces, vertex_normals=None):
        self.vertices = np.asarray(vertices, dtype=np.float64)
        self.faces = np.asarray(faces, dtype=np.int64)
        self.vertex_normals = vertex_normals

    def export(self, path):
        ext = str(path).rsplit(".", 1)[-1].lower()
        with open(path, "w") as f:
            if ext == "obj":
                f.writelines("v %.9g %.9g %.9g\n" % tuple(v) for v in self.vertices)
                f.writelines("f %d %d %d\n" % tuple(t + 1) for t in self.faces)
            elif ext == "off":
                f.write("OFF\n%d %d 0\n" % (len(self.vertices), len(self.faces)))
                f.writelines("%.9g %.9g %.9g\n" % tuple(v) for v in self.vertices)
                f.writelines("3 %d %d %d\n" % tuple(t) for t in self.faces)
            elif ext == "ply":
                f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                        "element face %d\nproperty list uchar int vertex_indices\nend_header\n"
                        % (len(self.vertices), len(self.faces)))
                f.writelines("%.9g %.9g %.9g\n" % tuple(v) for v in self.vertices)
                f.writelines("3 %d %d %d\n" % tuple(t) for t in self.faces)
            else:
                raise ValueError("export: .obj, .off or .ply")
        return path
