// MISE octree refinement step on the device state of slice3d_b200/mise.py.
//
// reference: MISE.subdivide_voxels / subdivide_voxel (reg_slices/src_convonet/utils/libmise/mise.pyx:184-283).  The
// reference walks a vector of grid points and, per point, descends the voxel tree for each of the 8 adjacent unit cells;
// then it walks the voxel vector and splits the marked leaves.  With the octree stored as `cell_level` (level of the leaf
// voxel containing each unit cell) both walks are embarrassingly parallel:
//   k_mise_mark    one thread per lattice point: a KNOWN point ORs "next to positive" (value >= threshold) / "next to
//                  negative" (value <= threshold) into the flag word of the leaf voxel of each adjacent unit cell.
//   k_mise_split   one thread per voxel of every level below the maximum depth: a leaf with both marks moves its unit
//                  cells one level down and adds the 27 lattice points of its 2x2x2 children; flag words are cleared.
#include "common.cuh"

namespace s3d {

namespace {

struct MiseDims {
  int R, res0, depth;
  long long level_off[16];  // offset of level l's flag words in the scratch (levels 0 .. depth-1)
};

__global__ void k_mise_mark(MiseDims d, double thr, const double* __restrict__ value, const unsigned char* __restrict__ known,
                            const signed char* __restrict__ cell_level, int* __restrict__ flags) {
  const long long P = d.R + 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * P * P) return;
  if (!known[idx]) return;
  const int z = (int)(idx % P);
  const long long t = idx / P;
  const int y = (int)(t % P), x = (int)(t / P);
  const double v = value[idx];
  const int f = (v >= thr ? 1 : 0) | (v <= thr ? 2 : 0);
#pragma unroll
  for (int a = 0; a < 8; ++a) {  // the 8 adjacent unit cells (mise.pyx:205-207: offsets -1, 0)
    const int cx = x - (a & 1), cy = y - ((a >> 1) & 1), cz = z - ((a >> 2) & 1);
    if (cx < 0 || cy < 0 || cz < 0 || cx >= d.R || cy >= d.R || cz >= d.R) continue;
    const int l = cell_level[((long long)cx * d.R + cy) * d.R + cz];
    if (l >= d.depth) continue;  // voxels of the maximum depth are never split: their marks are never read
    const int sh = d.depth - l, n = d.res0 << l;
    atomicOr(flags + d.level_off[l] + ((long long)(cx >> sh) * n + (cy >> sh)) * n + (cz >> sh), f);
  }
}

__global__ void k_mise_split(MiseDims d, signed char* __restrict__ cell_level, unsigned char* __restrict__ exists,
                             int* __restrict__ flags, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int f = flags[idx];
  if (f == 0) return;
  flags[idx] = 0;  // ready for the next round
  if (f != 3) return;
  int l = 0;
  while (l + 1 < d.depth && idx >= d.level_off[l + 1]) ++l;
  const long long loc = idx - d.level_off[l];
  const int n = d.res0 << l, size = 1 << (d.depth - l), half = size >> 1;
  const int vz = (int)(loc % n), vy = (int)((loc / n) % n), vx = (int)(loc / ((long long)n * n));
  const int x0 = vx * size, y0 = vy * size, z0 = vz * size;
  // (a voxel that is marked at level l is a leaf of level l: marks are made at the level of the containing leaf)
  for (int a = 0; a < size; ++a)
    for (int b = 0; b < size; ++b)
      for (int c = 0; c < size; ++c) cell_level[((long long)(x0 + a) * d.R + (y0 + b)) * d.R + (z0 + c)] = (signed char)(l + 1);
  const long long P = d.R + 1;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      for (int c = 0; c < 3; ++c) exists[((long long)(x0 + a * half) * P + (y0 + b * half)) * P + (z0 + c * half)] = 1;
}

}  // namespace

size_t mise_scratch_ints(int res0, int depth) {
  size_t n = 0;
  for (int l = 0; l < depth; ++l) n += (size_t)(res0 << l) * (res0 << l) * (res0 << l);
  return n ? n : 1;
}

int mise_subdivide(int res0, int depth, double thr, const double* value, const unsigned char* known, signed char* cell_level,
                   unsigned char* exists, int* flags_zeroed, cudaStream_t st) {
  if (res0 < 1 || depth < 0 || depth > 15 || !value || !known || !cell_level || !exists || !flags_zeroed) {
    set_error("mise_subdivide: bad argument");
    return S3D_ERR_BAD_ARG;
  }
  if (depth == 0) return S3D_OK;
  MiseDims d{};
  d.R = res0 << depth;
  d.res0 = res0;
  d.depth = depth;
  long long off = 0;
  for (int l = 0; l < depth; ++l) {
    d.level_off[l] = off;
    off += (long long)(res0 << l) * (res0 << l) * (res0 << l);
  }
  const long long P = d.R + 1, pts = P * P * P;
  k_mise_mark<<<(unsigned)((pts + 255) / 256), 256, 0, st>>>(d, thr, value, known, cell_level, flags_zeroed);
  S3D_LAUNCH_CHECK();
  k_mise_split<<<(unsigned)((off + 255) / 256), 256, 0, st>>>(d, cell_level, exists, flags_zeroed, off);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

}  // namespace s3d
