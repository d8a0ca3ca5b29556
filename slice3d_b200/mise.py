"""Multiresolution IsoSurface Extraction bookkeeping on the device.

Same results as the reference's ``MISE`` class (reference:
reg_slices/src_convonet/utils/libmise/mise.pyx:35-235), which keeps an octree of voxels, a hash of grid
points and python-side numpy round trips on the host.  Here the state is a handful of dense tensors on
the GPU (a 257^3 problem is 17 M cells: a few tens of MB out of 180 GB), every step is a vectorised
scatter / gather, and the points never leave the device between ``query`` and ``update``:

* ``cell_level[x, y, z]``  level of the leaf voxel that contains unit cell (x, y, z)  (int8, R^3);
  a voxel of level l has edge ``2^(depth - l)``, so this one array IS the octree.
* ``exists`` / ``known`` / ``value``  per grid point of the finest lattice ((R+1)^3).

``update`` reproduces mise.pyx:87-104,184-235 exactly: every known grid point marks the leaf voxels
that contain its 8 adjacent unit cells as "next to positive" (value >= threshold) and / or "next to
negative" (value <= threshold) -- this includes hanging nodes on the faces of coarser neighbours --
and every leaf voxel below the maximum depth with both marks is split, which adds the 27 lattice
points of its 2x2x2 children.  ``to_dense`` is mise.pyx:130-164: the existing points' values, then a
forward fill along x, then y, then z.

Only the ORDER of the queried points differs (flat-index order here, insertion order there); the
order has no effect on any value.
"""
import torch


class MISE(object):
    def __init__(self, resolution_0, depth, threshold, device="cpu"):
        self.resolution_0 = int(resolution_0)
        self.depth = int(depth)
        self.threshold = float(threshold)
        self.voxel_size_0 = 1 << self.depth
        self.resolution = self.resolution_0 * self.voxel_size_0
        self.device = torch.device(device)
        R = self.resolution
        self.cell_level = torch.zeros((R, R, R), dtype=torch.int8, device=self.device)
        self.exists = torch.zeros((R + 1,) * 3, dtype=torch.bool, device=self.device)
        self.known = torch.zeros((R + 1,) * 3, dtype=torch.bool, device=self.device)
        self.value = torch.zeros((R + 1,) * 3, dtype=torch.float64, device=self.device)
        s = self.voxel_size_0
        self.exists[::s, ::s, ::s] = True  # initial grid points (mise.pyx:75-85)
        # offsets of the 8 unit cells adjacent to a grid point (mise.pyx:205-207: range(-1, 1))
        o = torch.tensor([-1, 0], device=self.device)
        self._adj = torch.stack(torch.meshgrid(o, o, o, indexing="ij"), -1).reshape(8, 3)
        t = torch.arange(3, device=self.device)
        self._child_pts = torch.stack(torch.meshgrid(t, t, t, indexing="ij"), -1).reshape(27, 3)
        self._flags = None

    # ------------------------------------------------------------------ reference API
    def query(self):
        """(n, 3) int64 lattice coordinates of the grid points whose value is still unknown."""
        return torch.nonzero(self.exists & ~self.known)

    def update(self, points, values, validate=True):
        """Set ``values`` (n,) at ``points`` (n, 3) and split every active voxel.  ``validate=False`` skips the
        "Point not in grid!" check (one device synchronisation) when the points come straight from ``query()``."""
        points = torch.as_tensor(points, device=self.device).long()
        values = torch.as_tensor(values, device=self.device).double()
        if points.shape[0] != values.shape[0] or points.shape[1] != 3:
            raise ValueError("points must be (n, 3) and values (n,)")
        px, py, pz = points.unbind(1)
        if validate and points.numel() and not bool(self.exists[px, py, pz].all()):
            raise ValueError("Point not in grid!")
        self.value[px, py, pz] = values
        self.known[px, py, pz] = True
        self._subdivide_voxels()

    def to_dense(self):
        """(R+1)^3 float64 volume at the highest resolution (mise.pyx:130-164)."""
        out = torch.where(self.exists, self.value, torch.full_like(self.value, float("nan")))
        for axis in range(3):
            out = _forward_fill(out, axis)
        if bool(torch.isnan(out).any()):
            raise AssertionError("to_dense: unfilled grid point")
        return out

    def get_points(self):
        pts = torch.nonzero(self.exists)
        return pts, self.value[pts[:, 0], pts[:, 1], pts[:, 2]]

    def flags(self):
        """Zeroed int32 scratch of the CUDA refinement step (one word per voxel of every level below the maximum depth)."""
        if self._flags is None:
            from . import _native
            self._flags = torch.zeros(_native.mise_scratch_ints(self.resolution_0, self.depth), dtype=torch.int32,
                                      device=self.device)
        return self._flags

    # ------------------------------------------------------------------ internals
    def _subdivide_voxels(self):
        R, depth = self.resolution, self.depth
        if depth == 0:
            return
        if self.device.type == "cuda":
            # on the device: two hand-written kernels (csrc/mise.cu, s3d_mise_subdivide); the tensor program below is
            # the same step for host tensors (what the CPU tests run against the reference)
            from . import _native
            _native.mise_subdivide(self.resolution_0, depth, self.threshold, self.value, self.known, self.cell_level,
                                   self.exists, self.flags())
            return
        pts = torch.nonzero(self.known)
        val = self.value[pts[:, 0], pts[:, 1], pts[:, 2]]
        pos_pt, neg_pt = val >= self.threshold, val <= self.threshold
        # the 8 adjacent unit cells of every known point, those inside the volume
        cells = (pts[:, None, :] + self._adj[None, :, :]).reshape(-1, 3)
        pos_c = pos_pt[:, None].expand(-1, 8).reshape(-1)
        neg_c = neg_pt[:, None].expand(-1, 8).reshape(-1)
        inb = ((cells >= 0) & (cells < R)).all(1)
        cells, pos_c, neg_c = cells[inb], pos_c[inb], neg_c[inb]
        lvl = self.cell_level[cells[:, 0], cells[:, 1], cells[:, 2]].long()
        new_pts = []
        bump = torch.zeros_like(self.cell_level)
        for l in range(depth):  # voxels of the maximum depth are never split
            sel = lvl == l
            if not bool(sel.any()):
                continue
            n_l = self.resolution_0 << l
            shift = depth - l
            v = cells[sel] >> shift
            flat = (v[:, 0] * n_l + v[:, 1]) * n_l + v[:, 2]
            pos = torch.zeros(n_l ** 3, dtype=torch.bool, device=self.device)
            neg = torch.zeros(n_l ** 3, dtype=torch.bool, device=self.device)
            pos[flat[pos_c[sel]]] = True
            neg[flat[neg_c[sel]]] = True
            act = torch.nonzero((pos & neg).view(n_l, n_l, n_l))  # active leaf voxels of level l
            if act.shape[0] == 0:
                continue
            size = 1 << shift
            half = size >> 1
            # children: every unit cell of the voxel moves one level down
            a = (pos & neg).view(n_l, n_l, n_l)
            bump += a.repeat_interleave(size, 0).repeat_interleave(size, 1).repeat_interleave(size, 2).to(torch.int8)
            # the 27 lattice points of the 2x2x2 children (mise.pyx:268-283)
            new_pts.append((act[:, None, :] * size + self._child_pts[None, :, :] * half).reshape(-1, 3))
        self.cell_level += bump
        if new_pts:
            p = torch.cat(new_pts)
            self.exists[p[:, 0], p[:, 1], p[:, 2]] = True


def _forward_fill(x, axis):
    """Along ``axis``: every NaN takes the last non-NaN value before it (leading NaNs stay)."""
    n = x.shape[axis]
    shape = [1, 1, 1]
    shape[axis] = n
    idx = torch.arange(n, device=x.device).view(shape).expand_as(x)
    valid = ~torch.isnan(x)
    last = torch.cummax(torch.where(valid, idx, torch.zeros_like(idx)), dim=axis).values
    return torch.gather(x, axis, last)
