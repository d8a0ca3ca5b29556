"""Multi-GPU sharding of the dense query grid (one process per GPU).

Queries are independent given the planes, so the (nx,ny,nz) value volume is split into
contiguous slabs along axis 0 -- the slowest axis of ``make_3d_grid`` (reference:
reg_slices/src_convonet/common.py:159-162) -- which makes every rank's result one contiguous
range of the flat volume.  Each rank runs the (cheap) encoder redundantly; the only exchange
is one all-gather of the slabs (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def rank_world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def slab_bounds(nx, world):
    """Axis-0 boundaries [b_0=0, ..., b_world=nx]; the first nx % world ranks get one extra plane."""
    base, rem = divmod(nx, world)
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


def slab_range(nx, rank, world):
    b = slab_bounds(nx, world)
    return b[rank], b[rank + 1]


def proportional_bounds(nx, rates):
    """Axis-0 boundaries with slab widths proportional to ``rates`` (planes per unit time of each rank), every rank
    keeping at least one plane; largest-remainder rounding, deterministic for identical input on every rank."""
    world = len(rates)
    if nx < world:
        return slab_bounds(nx, world)
    total = float(sum(rates))
    ideal = [max(1.0, nx * float(r) / total) for r in rates]
    scale = nx / sum(ideal)
    ideal = [x * scale for x in ideal]
    w = [max(1, int(x)) for x in ideal]
    order = sorted(range(world), key=lambda i: (-(ideal[i] - int(ideal[i])), i))
    k = 0
    while sum(w) < nx:
        w[order[k % world]] += 1
        k += 1
    while sum(w) > nx:
        i = max(range(world), key=lambda j: (w[j], -j))
        w[i] -= 1
    b = [0]
    for x in w:
        b.append(b[-1] + x)
    return b


def split_at_planes(q0, q1, plane):
    """A contiguous query range [q0, q1) of the flat volume as at most three launch segments (first, count, whole):
    the rows before the first whole axis-0 plane, the whole planes, the rows after them.  Whole planes are decoded in the
    locality order of the library (csrc/common.cuh:grid_point); partial planes in flat order.  Values do not depend on
    the segmentation (every query is computed independently)."""
    if q1 <= q0:
        return []
    a = -(-q0 // plane) * plane  # first plane boundary >= q0
    b = (q1 // plane) * plane    # last plane boundary <= q1
    if a >= b:  # no whole plane inside the range: one launch in flat order
        return [(q0, q1 - q0, False)]
    segs = []
    if a > q0:
        segs.append((q0, a - q0, False))
    segs.append((a, b - a, True))
    if q1 > b:
        segs.append((b, q1 - b, False))
    return segs


def all_gather_slabs(vol_flat, nx, group=None, bounds=None, units=None):
    """In place: every rank has filled its own slab of ``vol_flat`` (nx^3 values, flat; slab r = units
    [bounds[r], bounds[r+1]) of ``units`` equal parts of the volume -- axis-0 planes by default (units = nx), rows when
    units = nx * nx; equal plane slabs by default); afterwards every rank holds the whole volume."""
    rank, world = rank_world(group)
    if world == 1:
        return vol_flat
    units = nx if units is None else int(units)
    plane = vol_flat.numel() // units
    b = list(bounds) if bounds is not None else slab_bounds(nx, world)
    if all(b[r + 1] - b[r] == b[1] - b[0] for r in range(world)):
        lo, hi = b[rank], b[rank + 1]
        # equal slabs: a single all-gather straight into the volume.  NCCL gathers in place when the input IS the
        # rank's slot of the output (no staging copy); the gloo backend of the CPU tests needs a separate input.
        mine = vol_flat[lo * plane:hi * plane]
        dist.all_gather_into_tensor(vol_flat, mine if vol_flat.is_cuda else mine.clone(), group=group)
    else:
        outs = [vol_flat[b[r] * plane:b[r + 1] * plane] for r in range(world)]
        if hasattr(dist, "all_gather") and all(o.numel() == outs[0].numel() for o in outs):
            dist.all_gather(outs, outs[rank].clone(), group=group)
        else:
            for r in range(world):  # ragged slabs: one broadcast per owner
                dist.broadcast(outs[r], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
    return vol_flat


def all_gather_ranges(flat, n, group=None):
    """In place: rank r has filled ``flat[slab_range(n, r, world)]``; afterwards every rank holds all n values."""
    rank, world = rank_world(group)
    if world == 1:
        return flat
    b = slab_bounds(n, world)
    if n % world == 0:
        lo, hi = b[rank], b[rank + 1]
        dist.all_gather_into_tensor(flat, flat[lo:hi].clone(), group=group)
    else:
        for r in range(world):  # ragged shares: one broadcast per owner
            if b[r + 1] > b[r]:
                dist.broadcast(flat[b[r]:b[r + 1]], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
    return flat
