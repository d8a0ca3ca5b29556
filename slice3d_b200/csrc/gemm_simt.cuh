// fp32 CUDA-core implicit GEMM used by the encoder (3x3 / 1x1 / transposed convolutions on
// NHWC activations) and by the validation-grade fp32 decoder.
//
//   C[m][n] = sum_k A(m,k) * W[k][n]        W row-major [Kpad][N], rows >= K are zero.
//
// A(m,k) is produced by a loader functor (plain row-major matrix, or an on-the-fly im2col of
// up to two NHWC sources = the U-Net's channel concat, the first one optionally broadcast over
// the K slices of an image).  The epilogue functor receives 4 consecutive columns of one row.
//
// Tile 128 x BN x 16, 256 threads, 8 x (BN/16) outputs per thread, double-buffered smem with
// register prefetch.  This is the exact-fp32 path; the tensor-core paths live in *_tc.cu.
#pragma once
#include "common.cuh"

namespace s3d {

constexpr int GM_BM = 128;
constexpr int GM_BK = 16;
constexpr int GM_THREADS = 256;

// ---------------------------------------------------------------- loaders
// Every loader exposes: int M, K;  struct Row;  Row row(int m);  float4 load(const Row&, int k)
// with k a multiple of 4 (and k < Kpad).

struct LoadPlain {  // A is [M][lda] row-major, K % 4 == 0
  const float* a;
  int M, K, lda;
  struct Row {
    const float* p;  // null when m >= M
  };
  __device__ __forceinline__ Row row(int m) const { return Row{m < M ? a + (size_t)m * lda : nullptr}; }
  __device__ __forceinline__ float4 load(const Row& r, int k) const {
    if (r.p == nullptr || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
    return __ldg(reinterpret_cast<const float4*>(r.p + k));
  }
};

// 3x3 (pad 1) or 1x1 convolution over NHWC input(s).  Row m = (n, y, x) of the output
// (same H x W as the input).  k = tap * Cin + ci, tap = ky*3+kx; channels [0,C0) come from
// src0 whose image index is n / bcast0 (bcast0 = K when the skip tensor is shared by the K
// slices of one input view, reference unet_custom.py:35-38 expand_bs), channels [C0,Cin)
// from src1 (image index n).  C0, C1 multiples of 4.
struct LoadConv {
  const float* src0;
  const float* src1;
  int M, K;  // M = N*H*W, K = taps*Cin
  int H, W, C0, C1, bcast0, ks;
  struct Row {
    int n, y, x;  // n < 0 => out of range
  };
  __device__ __forceinline__ Row row(int m) const {
    if (m >= M) return Row{-1, 0, 0};
    int x = m % W;
    int t = m / W;
    return Row{t / H, t % H, x};
  }
  __device__ __forceinline__ float4 load(const Row& r, int k) const {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r.n < 0 || k >= K) return z;
    const int Cin = C0 + C1;
    int tap = 0, ci = k;
    int yy = r.y, xx = r.x;
    if (ks == 3) {
      tap = k / Cin;
      ci = k - tap * Cin;
      int ky = tap / 3;
      yy += ky - 1;
      xx += (tap - ky * 3) - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) return z;
    }
    const float* p;
    if (ci < C0)
      p = src0 + (((size_t)(r.n / bcast0) * H + yy) * W + xx) * C0 + ci;
    else
      p = src1 + (((size_t)r.n * H + yy) * W + xx) * C1 + (ci - C0);
    return __ldg(reinterpret_cast<const float4*>(p));
  }
};

// ---------------------------------------------------------------- epilogues
// void operator()(int m, int n, const float v[4])  -- columns n..n+3 of row m (m < M, n < N).

struct EpiAffine {  // out[m*ldc+n] = act(v*scale[n] + shift[n]); scale/shift may be null
  float* out;
  const float* scale;
  const float* shift;
  int ldc, relu;
  __device__ __forceinline__ void operator()(int m, int n, const float* v) const {
    float4 r;
    float* rr = reinterpret_cast<float*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t = v[i];
      if (scale) t *= __ldg(scale + n + i);
      if (shift) t += __ldg(shift + n + i);
      if (relu) t = fmaxf(t, 0.f);
      rr[i] = t;
    }
    *reinterpret_cast<float4*>(out + (size_t)m * ldc + n) = r;
  }
};

// ConvTranspose2d(k=2, s=2) written as a GEMM with N = 4*Cout, n = (dy*2+dx)*Cout + co
// (reference unet_parts.py:53,66): out[(img, 2y+dy, 2x+dx, co)] = v + bias[co].
struct EpiShuffle2x {
  float* out;
  const float* bias;
  int H, W, Cout;  // input H, W
  __device__ __forceinline__ void operator()(int m, int n, const float* v) const {
    int x = m % W;
    int t = m / W;
    int y = t % H, img = t / H;
    int q = n / Cout, co = n - q * Cout;
    int dy = q >> 1, dx = q & 1;
    float4 r = make_float4(v[0] + __ldg(bias + co), v[1] + __ldg(bias + co + 1), v[2] + __ldg(bias + co + 2),
                           v[3] + __ldg(bias + co + 3));
    size_t o = (((size_t)img * (2 * H) + (2 * y + dy)) * (2 * W) + (2 * x + dx)) * Cout + co;
    *reinterpret_cast<float4*>(out + o) = r;
  }
};

// Split-fp16 variants (hi = fp16(x) saturated, lo = fp16(x - hi); x = hi + lo + O(2^-22 |x|)): producers of the
// tensor-core convolutions' inputs (conv_tc.cu).  ldc = channel pitch of the split tensor.
__device__ __forceinline__ void store_split4(__half* hi, __half* lo, size_t o, const float* r) {
  float c[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = fminf(fmaxf(r[i], -65504.f), 65504.f);
  const __half2 h01 = __floats2half2_rn(c[0], c[1]), h23 = __floats2half2_rn(c[2], c[3]);
  const __half2 l01 = __floats2half2_rn(r[0] - __low2float(h01), r[1] - __high2float(h01));
  const __half2 l23 = __floats2half2_rn(r[2] - __low2float(h23), r[3] - __high2float(h23));
  uint2 hv, lv;
  hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
  lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
  *reinterpret_cast<uint2*>(hi + o) = hv;
  *reinterpret_cast<uint2*>(lo + o) = lv;
}

struct EpiAffineSplit {
  __half* hi;
  __half* lo;
  const float* scale;
  const float* shift;
  int ldc, relu;
  __device__ __forceinline__ void operator()(int m, int n, const float* v) const {
    float r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t = v[i];
      if (scale) t *= __ldg(scale + n + i);
      if (shift) t += __ldg(shift + n + i);
      if (relu) t = fmaxf(t, 0.f);
      r[i] = t;
    }
    store_split4(hi, lo, (size_t)m * ldc + n, r);
  }
};

struct EpiShuffle2xSplit {  // EpiShuffle2x writing the split format with channel pitch ldc >= Cout
  __half* hi;
  __half* lo;
  const float* bias;
  int H, W, Cout, ldc;
  __device__ __forceinline__ void operator()(int m, int n, const float* v) const {
    int x = m % W;
    int t = m / W;
    int y = t % H, img = t / H;
    int q = n / Cout, co = n - q * Cout;
    int dy = q >> 1, dx = q & 1;
    const float r[4] = {v[0] + __ldg(bias + co), v[1] + __ldg(bias + co + 1), v[2] + __ldg(bias + co + 2),
                        v[3] + __ldg(bias + co + 3)};
    size_t o = (((size_t)img * (2 * H) + (2 * y + dy)) * (2 * W) + (2 * x + dx)) * ldc + co;
    store_split4(hi, lo, o, r);
  }
};

// ---------------------------------------------------------------- kernel
template <int BN, class Loader, class Epi>
__global__ void __launch_bounds__(GM_THREADS) gemm_simt_kernel(Loader L, const float* __restrict__ Wt, int N, int Kpad,
                                                               Epi E) {
  constexpr int TN = BN / 16;  // columns per thread: 8 (BN=128), 4 (BN=64), 2 (BN=32)
  __shared__ __align__(16) float As[2][GM_BK][GM_BM + 4];
  __shared__ __align__(16) float Bs[2][GM_BK][BN];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * GM_BM;
  const int n0 = blockIdx.y * BN;
  const int ty = tid >> 4, tx = tid & 15;

  // A staging: 128 rows x 4 float4 per k-tile -> 2 float4 per thread.
  const int a_row = tid >> 2, a_kq = (tid & 3) * 4;
  typename Loader::Row r0 = L.row(m0 + a_row);
  typename Loader::Row r1 = L.row(m0 + a_row + 64);
  // B staging: 16 x BN floats = 4*BN float4 ... per thread (BN*16/4)/256 float4.
  constexpr int B_F4 = (GM_BK * BN / 4 + GM_THREADS - 1) / GM_THREADS;  // 2, 1, 1(half idle)

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 pa0, pa1, pb[B_F4];
  auto fetch = [&](int k0) {
    pa0 = L.load(r0, k0 + a_kq);
    pa1 = L.load(r1, k0 + a_kq);
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
      int f = tid + i * GM_THREADS;  // float4 index inside the [16][BN] tile
      int kr = f / (BN / 4), nc = (f % (BN / 4)) * 4;
      if (kr < GM_BK && n0 + nc < N)
        pb[i] = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)(k0 + kr) * N + n0 + nc));
      else
        pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stage = [&](int buf) {
    As[buf][a_kq + 0][a_row] = pa0.x;
    As[buf][a_kq + 1][a_row] = pa0.y;
    As[buf][a_kq + 2][a_row] = pa0.z;
    As[buf][a_kq + 3][a_row] = pa0.w;
    As[buf][a_kq + 0][a_row + 64] = pa1.x;
    As[buf][a_kq + 1][a_row + 64] = pa1.y;
    As[buf][a_kq + 2][a_row + 64] = pa1.z;
    As[buf][a_kq + 3][a_row + 64] = pa1.w;
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
      int f = tid + i * GM_THREADS;
      int kr = f / (BN / 4), nc = (f % (BN / 4)) * 4;
      if (kr < GM_BK) *reinterpret_cast<float4*>(&Bs[buf][kr][nc]) = pb[i];
    }
  };

  const int nk = Kpad / GM_BK;
  fetch(0);
  stage(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * GM_BK);
#pragma unroll
    for (int kk = 0; kk < GM_BK; ++kk) {
      float a[8], b[TN];
      float4 t0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 t1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      a[0] = t0.x; a[1] = t0.y; a[2] = t0.z; a[3] = t0.w;
      a[4] = t1.x; a[5] = t1.y; a[6] = t1.z; a[7] = t1.w;
      if constexpr (TN == 8) {
        float4 u0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        float4 u1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
        b[0] = u0.x; b[1] = u0.y; b[2] = u0.z; b[3] = u0.w;
        b[4] = u1.x; b[5] = u1.y; b[6] = u1.z; b[7] = u1.w;
      } else if constexpr (TN == 4) {
        float4 u0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        b[0] = u0.x; b[1] = u0.y; b[2] = u0.z; b[3] = u0.w;
      } else {
        float2 u0 = *reinterpret_cast<const float2*>(&Bs[buf][kk][tx * 2]);
        b[0] = u0.x; b[1] = u0.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      stage(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: rows ty*4+i (i<4) and 64+ty*4+(i-4); columns per TN layout above
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if constexpr (TN == 2) {
      // pair lanes (tx even, tx odd) to hand the epilogue 4 contiguous columns; the shuffles
      // run before any row guard so that every lane takes part
      float v[4];
      float o0 = __shfl_xor_sync(0xffffffffu, acc[i][0], 1);
      float o1 = __shfl_xor_sync(0xffffffffu, acc[i][1], 1);
      if ((tx & 1) == 0 && m < L.M) {
        v[0] = acc[i][0]; v[1] = acc[i][1]; v[2] = o0; v[3] = o1;
        int na = n0 + tx * 2;
        if (na < N) E(m, na, v);
      }
      continue;
    }
    if (m >= L.M) continue;
    if constexpr (TN == 8) {
      int na = n0 + tx * 4, nb = n0 + 64 + tx * 4;
      if (na < N) E(m, na, &acc[i][0]);
      if (nb < N) E(m, nb, &acc[i][4]);
    } else if constexpr (TN == 4) {
      int na = n0 + tx * 4;
      if (na < N) E(m, na, &acc[i][0]);
    }
  }
}

template <class Loader, class Epi>
int launch_gemm(const Loader& L, const float* Wt, int N, int Kpad, const Epi& E, cudaStream_t st) {
  if (L.M <= 0) return S3D_OK;
  dim3 grid((L.M + GM_BM - 1) / GM_BM, 1, 1);
  if (N > 64) {
    grid.y = (N + 127) / 128;
    gemm_simt_kernel<128, Loader, Epi><<<grid, GM_THREADS, 0, st>>>(L, Wt, N, Kpad, E);
  } else if (N > 32) {
    gemm_simt_kernel<64, Loader, Epi><<<grid, GM_THREADS, 0, st>>>(L, Wt, N, Kpad, E);
  } else {
    gemm_simt_kernel<32, Loader, Epi><<<grid, GM_THREADS, 0, st>>>(L, Wt, N, Kpad, E);
  }
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

}  // namespace s3d
