// Marching cubes over a float64 value volume, two passes around an exclusive prefix sum.
//
// reference: libmcubes' sequential scan (reg_slices/src_convonet/utils/libmcubes/marchingcubes.h:22-196, called by
// reconstruct.py:190 through pywrapper.cpp:90-128).  The sequential code numbers vertices with a running counter and
// remembers the indices of the three edges a cell "owns" (6, 5, 10) for its +x/+y/+z neighbours.  Here:
//   k_mc_count   one thread per cell: cube configuration, which crossed edges create a NEW vertex in this cell (the owned
//                edges always; the others only on the low boundary faces, where the reference duplicates vertices),
//                per-cell vertex and triangle counts, and the 3-bit "owned edges crossed" mask.
//   (host)       exclusive prefix sums of the two count arrays (the running counters of the sequential scan).
//   k_mc_emit    one thread per cell: writes its new vertices (same float64 expression as mc_add_vertex) at
//                base + rank, resolves shared edges through the owning neighbour's base and mask, writes its triangles.
// The vertex array equals the reference's bit for bit; so does the face array when the caller supplies the classic
// Lorensen-Cline / Bourke table the reference uses (slice3d_b200/mc_table.py).
#include "common.cuh"

namespace s3d {

namespace {

struct McDims {
  int nx, ny, nz, cx, cy, cz;
};

// corner m of cell (i,j,k): offsets as in marchingcubes.h:60-64
__device__ __forceinline__ void load_corners(const double* __restrict__ vol, const McDims& d, int i, int j, int k, double* v) {
  const size_t sy = d.nz, sx = (size_t)d.ny * d.nz;
  const double* p = vol + (size_t)i * sx + (size_t)j * sy + k;
  v[0] = p[0];        v[1] = p[sx];           v[2] = p[sx + sy];      v[3] = p[sy];
  v[4] = p[1];        v[5] = p[sx + 1];       v[6] = p[sx + sy + 1];  v[7] = p[sy + 1];
}

// edge e joins corners EA[e], EB[e]
__constant__ int c_ea[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3};
__constant__ int c_eb[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
// creation order inside a cell and mc_add_vertex's (start corner, end corner, axis) per edge (marchingcubes.h:74-176)
__constant__ int c_order[12] = {6, 5, 10, 0, 1, 2, 3, 4, 7, 8, 9, 11};
__constant__ int c_start[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3};
__constant__ int c_end[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
__constant__ int c_axis[12] = {0, 1, 0, 1, 0, 1, 0, 1, 2, 2, 2, 2};
__constant__ int c_cornx[8] = {0, 1, 1, 0, 0, 1, 1, 0};
__constant__ int c_corny[8] = {0, 0, 1, 1, 0, 0, 1, 1};
__constant__ int c_cornz[8] = {0, 0, 0, 0, 1, 1, 1, 1};

__device__ __forceinline__ unsigned cube_index(const double* v, double iso) {
  unsigned c = 0;
#pragma unroll
  for (int m = 0; m < 8; ++m)
    if (v[m] <= iso) c |= 1u << m;
  return c;
}
__device__ __forceinline__ unsigned crossed_mask(unsigned cube) {
  unsigned m = 0;
#pragma unroll
  for (int e = 0; e < 12; ++e)
    if (((cube >> c_ea[e]) ^ (cube >> c_eb[e])) & 1u) m |= 1u << e;
  return m;
}
// edges that are NOT created here but read from the neighbour that owns them (marchingcubes.h:93-176)
__device__ __forceinline__ unsigned shared_mask(int i, int j, int k) {
  unsigned m = 0;
  if (j > 0 && k > 0) m |= 1u << 0;
  if (k > 0) m |= (1u << 1) | (1u << 2);
  if (i > 0 && k > 0) m |= 1u << 3;
  if (j > 0) m |= (1u << 4) | (1u << 9);
  if (i > 0) m |= (1u << 7) | (1u << 11);
  if (i > 0 && j > 0) m |= 1u << 8;
  return m;
}

__global__ void k_mc_count(const double* __restrict__ vol, McDims d, double iso, const int* __restrict__ tri_count,
                           int* __restrict__ vcount, int* __restrict__ tcount, unsigned char* __restrict__ owned) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= (long long)d.cx * d.cy * d.cz) return;
  const int k = (int)(c % d.cz);
  const long long t = c / d.cz;
  const int j = (int)(t % d.cy), i = (int)(t / d.cy);
  double v[8];
  load_corners(vol, d, i, j, k, v);
  const unsigned cube = cube_index(v, iso);
  const unsigned cr = crossed_mask(cube);
  const unsigned nw = cr & ~shared_mask(i, j, k);
  vcount[c] = __popc(nw);
  tcount[c] = tri_count[cube];
  owned[c] = (unsigned char)(((cr >> 6) & 1u) | (((cr >> 5) & 1u) << 1) | (((cr >> 10) & 1u) << 2));
}

__global__ void k_mc_emit(const double* __restrict__ vol, McDims d, double iso, const signed char* __restrict__ table,
                          const long long* __restrict__ vbase, const long long* __restrict__ tbase,
                          const int* __restrict__ tcount, const unsigned char* __restrict__ owned, double* __restrict__ verts,
                          long long* __restrict__ tris) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= (long long)d.cx * d.cy * d.cz) return;
  const int k = (int)(c % d.cz);
  const long long t = c / d.cz;
  const int j = (int)(t % d.cy), i = (int)(t / d.cy);
  double v[8];
  load_corners(vol, d, i, j, k, v);
  const unsigned cube = cube_index(v, iso);
  const unsigned cr = crossed_mask(cube);
  if (cr == 0) return;
  const unsigned nw = cr & ~shared_mask(i, j, k);
  long long idx[12];
  long long next = vbase[c];
  const double bx = (double)i + 0.5, by = (double)j + 0.5, bz = (double)k + 0.5;  // x = lower + dx*i + dx/2, dx = 1
#pragma unroll
  for (int o = 0; o < 12; ++o) {
    const int e = c_order[o];
    if (!((nw >> e) & 1u)) continue;
    idx[e] = next;
    const int a = c_start[e], b = c_end[e], ax = c_axis[e];
    double pos[3] = {bx + c_cornx[a], by + c_corny[a], bz + c_cornz[a]};
    const double x1 = pos[ax];
    const double x2 = (ax == 0 ? bx + c_cornx[b] : ax == 1 ? by + c_corny[b] : bz + c_cornz[b]);
    const double f1 = v[a], f2 = v[b];
    // mc_isovalue_interpolation (marchingcubes.cpp:290-297), no contraction
    pos[ax] = (f2 == f1) ? __ddiv_rn(__dadd_rn(x2, x1), 2.0)
                         : __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(x2, x1), __dsub_rn(iso, f1)), __dsub_rn(f2, f1)), x1);
    verts[3 * next + 0] = pos[0];
    verts[3 * next + 1] = pos[1];
    verts[3 * next + 2] = pos[2];
    ++next;
  }
  // shared edges: index of the owning neighbour's edge 6 / 5 / 10 = its base + rank among its crossed owned edges
  const long long sz = 1, sy = d.cz, sx = (long long)d.cy * d.cz;
  auto owner_index = [&](long long n, int which) -> long long {  // which: 0 = edge 6, 1 = edge 5, 2 = edge 10
    const unsigned m = owned[n];
    const int rank = (which >= 1 ? (m & 1u) : 0) + (which >= 2 ? ((m >> 1) & 1u) : 0);
    return vbase[n] + rank;
  };
  const unsigned sh = cr & shared_mask(i, j, k);
  if ((sh >> 0) & 1u) idx[0] = owner_index(c - sy - sz, 0);
  if ((sh >> 1) & 1u) idx[1] = owner_index(c - sz, 1);
  if ((sh >> 2) & 1u) idx[2] = owner_index(c - sz, 0);
  if ((sh >> 3) & 1u) idx[3] = owner_index(c - sx - sz, 1);
  if ((sh >> 4) & 1u) idx[4] = owner_index(c - sy, 0);
  if ((sh >> 7) & 1u) idx[7] = owner_index(c - sx, 1);
  if ((sh >> 8) & 1u) idx[8] = owner_index(c - sx - sy, 2);
  if ((sh >> 9) & 1u) idx[9] = owner_index(c - sy, 2);
  if ((sh >> 11) & 1u) idx[11] = owner_index(c - sx, 2);
  const int nt = tcount[c];
  const signed char* row = table + cube * 15;
  long long* out = tris + 3 * tbase[c];
  for (int q = 0; q < nt; ++q) {
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      const int e = row[3 * q + w];
      long long val = 0;
#pragma unroll
      for (int s = 0; s < 12; ++s)
        if (s == e) val = idx[s];  // (register-resident select instead of a dynamically indexed local array)
      out[3 * q + w] = val;
    }
  }
}

}  // namespace

int mc_count(const double* vol, int nx, int ny, int nz, double iso, const int* tri_count, int* vcount, int* tcount,
             unsigned char* owned, cudaStream_t st) {
  if (!vol || nx < 2 || ny < 2 || nz < 2 || !tri_count || !vcount || !tcount || !owned) {
    set_error("marching cubes: bad argument (the volume needs at least 2 samples per axis)");
    return S3D_ERR_BAD_ARG;
  }
  McDims d{nx, ny, nz, nx - 1, ny - 1, nz - 1};
  const long long cells = (long long)d.cx * d.cy * d.cz;
  k_mc_count<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(vol, d, iso, tri_count, vcount, tcount, owned);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

int mc_emit(const double* vol, int nx, int ny, int nz, double iso, const signed char* table, const long long* vbase,
            const long long* tbase, const int* tcount, const unsigned char* owned, double* verts, long long* tris,
            cudaStream_t st) {
  if (!vol || nx < 2 || ny < 2 || nz < 2 || !table || !vbase || !tbase || !tcount || !owned) {
    set_error("marching cubes: bad argument");
    return S3D_ERR_BAD_ARG;
  }
  McDims d{nx, ny, nz, nx - 1, ny - 1, nz - 1};
  const long long cells = (long long)d.cx * d.cy * d.cz;
  k_mc_emit<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(vol, d, iso, table, vbase, tbase, tcount, owned, verts, tris);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

// ---- exclusive prefix sum int32 -> int64 (the sequential scan's running counters) --------------------------------------
namespace {
constexpr int SCB = 2048;  // elements per block (256 threads x 8)

__global__ void __launch_bounds__(256) k_scan_block_sums(const int* __restrict__ in, long long n, long long* __restrict__ sums) {
  __shared__ long long w[8];
  const long long base = (long long)blockIdx.x * SCB + threadIdx.x * 8;
  long long s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += base + j < n ? in[base + j] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int i = 0; i < 8; ++i) t += w[i];
    sums[blockIdx.x] = t;
  }
}
// in place: sums[i] <- sums[0] + ... + sums[i-1]; *total = everything.  One block.
__global__ void __launch_bounds__(1024) k_scan_sums(long long* __restrict__ sums, int nb, long long* __restrict__ total) {
  __shared__ long long part[1024];
  const int t = threadIdx.x, per = (nb + 1023) / 1024, lo = t * per, hi = min(nb, lo + per);
  long long s = 0;
  for (int i = lo; i < hi; ++i) s += sums[i];
  part[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const long long v = t >= o ? part[t - o] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  long long run = t ? part[t - 1] : 0;
  for (int i = lo; i < hi; ++i) {
    const long long c = sums[i];
    sums[i] = run;
    run += c;
  }
  if (t == 1023) *total = part[1023];
}
__global__ void __launch_bounds__(256) k_scan_write(const int* __restrict__ in, long long n, const long long* __restrict__ sums,
                                                    long long* __restrict__ out) {
  __shared__ long long w[8];
  const long long base = (long long)blockIdx.x * SCB + threadIdx.x * 8;
  int v[8];
  long long s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = base + j < n ? in[base + j] : 0;
    s += v[j];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) w[warp] = incl;
  __syncthreads();
  long long run = sums[blockIdx.x] + incl - s;
  for (int i = 0; i < warp; ++i) run += w[i];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (base + j < n) out[base + j] = run;
    run += v[j];
  }
}
}  // namespace

size_t scan_scratch_bytes(long long n) { return (size_t)((n + SCB - 1) / SCB + 1) * sizeof(long long); }

int exclusive_scan_i32_i64(const int* in, long long n, long long* out, long long* total, void* scratch, cudaStream_t st) {
  if (n == 0) {
    S3D_CUDA(cudaMemsetAsync(total, 0, sizeof(long long), st));
    return S3D_OK;
  }
  const int nb = (int)((n + SCB - 1) / SCB);
  long long* sums = static_cast<long long*>(scratch);
  k_scan_block_sums<<<nb, 256, 0, st>>>(in, n, sums);
  S3D_LAUNCH_CHECK();
  k_scan_sums<<<1, 1024, 0, st>>>(sums, nb, total);
  S3D_LAUNCH_CHECK();
  k_scan_write<<<nb, 256, 0, st>>>(in, n, sums, out);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

}  // namespace s3d
