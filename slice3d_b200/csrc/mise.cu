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

// ---- device-resident query / update (one MISE round without a host round trip) -------------------------------
// MISE.query (mise.pyx:106-128) = the lattice points that exist and have no value yet, in flat-index order.  A
// deterministic three-step compaction: per-block counts, a one-block exclusive scan, per-block ranked writes.
constexpr int QB = 1024;  // lattice points per block (256 threads x 4)

__device__ __forceinline__ bool is_query(const unsigned char* exists, const unsigned char* known, long long i, long long n) {
  return i < n && exists[i] && !known[i];
}
__global__ void __launch_bounds__(256) k_mise_qcount(const unsigned char* __restrict__ exists, const unsigned char* __restrict__ known,
                                                     long long n, int* __restrict__ blk) {
  __shared__ int wsum[8];
  const long long base = (long long)blockIdx.x * QB + threadIdx.x * 4;
  int c = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) c += is_query(exists, known, base + j, n) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += wsum[w];
    blk[blockIdx.x] = t;
  }
}
// in place: blk[i] <- sum of blk[0..i); total -> *count (clamped to cap, overflow flagged) and round_count
__global__ void __launch_bounds__(1024) k_mise_qscan(int* __restrict__ blk, int nblk, int* __restrict__ count, int cap,
                                                     int* __restrict__ round_count, int* __restrict__ overflow) {
  __shared__ int part[1024];
  const int t = threadIdx.x;
  const int per = (nblk + 1023) / 1024;
  const int lo = t * per, hi = min(nblk, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += blk[i];
  part[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan of the 1024 partials
    const int v = t >= o ? part[t - o] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int run = t ? part[t - 1] : 0;
  for (int i = lo; i < hi; ++i) {
    const int c = blk[i];
    blk[i] = run;
    run += c;
  }
  if (t == 1023) {
    const int total = part[1023];
    if (total > cap) *overflow = 1;
    *count = min(total, cap);
    *round_count = total;
  }
}
// pt_idx[rank] = flat lattice index; pts[rank] = box * (p / R - 0.5) per axis, evaluated in float64 and rounded to float32
// like the reference's numpy -> torch.FloatTensor path (reconstruct.py:150-154)
__global__ void __launch_bounds__(256) k_mise_qemit(const unsigned char* __restrict__ exists, const unsigned char* __restrict__ known,
                                                    long long n, int R, double box, const int* __restrict__ blk, int cap,
                                                    int* __restrict__ pt_idx, float* __restrict__ pts) {
  __shared__ int woff[8];
  const long long base = (long long)blockIdx.x * QB + threadIdx.x * 4;
  bool q[4];
  int c = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    q[j] = is_query(exists, known, base + j, n);
    c += q[j] ? 1 : 0;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) woff[warp] = incl;
  __syncthreads();
  int wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += woff[w];
  int pos = blk[blockIdx.x] + wbase + incl - c;
  const long long P = R + 1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!q[j]) continue;
    if (pos < cap) {
      const long long i = base + j;
      const int z = (int)(i % P);
      const long long t = i / P;
      const int y = (int)(t % P), x = (int)(t / P);
      pt_idx[pos] = (int)i;
      pts[3 * pos + 0] = (float)(box * ((double)x / R - 0.5));
      pts[3 * pos + 1] = (float)(box * ((double)y / R - 0.5));
      pts[3 * pos + 2] = (float)(box * ((double)z / R - 0.5));
    }
    ++pos;
  }
}
// MISE.update's value store (mise.pyx:87-104): value[p] = v, known[p] = 1 for the round's points
__global__ void __launch_bounds__(256) k_mise_apply(const int* __restrict__ count, const int* __restrict__ pt_idx,
                                                    const float* __restrict__ vals, double* __restrict__ value,
                                                    unsigned char* __restrict__ known) {
  const int n = *count;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const int p = pt_idx[i];
    value[p] = (double)vals[i];
    known[p] = 1;
  }
}

}  // namespace

int mise_query_device(int R, double box, const unsigned char* exists, const unsigned char* known, int* blk, int* count, int cap,
                      int* round_count, int* overflow, int* pt_idx, float* pts, cudaStream_t st) {
  const long long P = R + 1, n = P * P * P;
  const int nblk = (int)((n + QB - 1) / QB);
  k_mise_qcount<<<nblk, 256, 0, st>>>(exists, known, n, blk);
  S3D_LAUNCH_CHECK();
  k_mise_qscan<<<1, 1024, 0, st>>>(blk, nblk, count, cap, round_count, overflow);
  S3D_LAUNCH_CHECK();
  k_mise_qemit<<<nblk, 256, 0, st>>>(exists, known, n, R, box, blk, cap, pt_idx, pts);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}
size_t mise_query_blocks(int R) {
  const long long P = R + 1, n = P * P * P;
  return (size_t)((n + QB - 1) / QB);
}
int mise_apply_device(const int* count, const int* pt_idx, const float* vals, double* value, unsigned char* known,
                      cudaStream_t st) {
  k_mise_apply<<<296, 256, 0, st>>>(count, pt_idx, vals, value, known);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

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
