// Train-mode decoder: forward AND backward of the per-query path of Slices3DRegModel.forward in training
// (reference: reg_slices/src/models.py:57-84 under model.train(), driven by reg_slices/train.py:41-53) as hand-written
// fp32 CUDA kernels behind s3d_train_decoder_fwd / s3d_train_decoder_bwd.
//
//   project_coord -> 5 x grid_sample(bilinear, zeros, align_corners) of the NCHW feature planes -> fc_s / fc_p ->
//   3 x nn.TransformerEncoderLayer(d_model 128, 4 heads, FFN 2048, ReLU, post-norm, dropout p) over K+1 tokens ->
//   fc_out on token 0
//
// and the gradient of sdf_pred with respect to the five feature planes and the 42 parameter tensors.  The U-Net and
// the VGG19 loss keep their autograd on the host side (torch); this file covers rows a3-a9 of SURVEY.md section 8 in train mode.
//
// Everything is exact fp32 (CUDA-core FMA, fp32 accumulate): the gradients are compared with the reference's own
// autograd gradients (tests/golden/train_*.npz).  Layout conventions are PyTorch's: Linear weights are (out, in)
// row-major and are used in place, no repacking -- the optimizer updates them every step.
//
//   GEMM flavours (one tiled kernel, 128 x 128 x 16, 256 threads, 8 x 8 per thread):
//     NT  Y[M,N]  = X[M,K] . W[N,K]^T (+ bias, ReLU, dropout)          forward of every Linear
//     NN  dX[M,K] = dY[M,N] . W[N,K]  (+=, ReLU/dropout mask)          input gradients
//     TN  dW[N,K] = dY[M,N]^T . X[M,K]   split over M, partials reduced in a second kernel (deterministic)
//   Dropout masks are a counter-based hash of (seed, site, element): regenerated in backward, never stored.
#include "common.cuh"

namespace s3d {

namespace {

constexpr int D = 128;      // d_model
constexpr int NH = 4;       // heads
constexpr int HD = 32;      // head dim
constexpr int FF = 2048;    // FFN width
constexpr int CAGG = 992;   // 512 + 256 + 128 + 64 + 32

// ---------------------------------------------------------------- dropout
// keep(seed, site, i): uniform 24-bit hash >= p * 2^24.  splitmix64 finaliser.
__device__ __forceinline__ uint32_t hash24(unsigned long long seed, unsigned site, unsigned long long i) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (i + ((unsigned long long)site << 44) + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 40);
}
struct Drop {
  unsigned long long seed;
  uint32_t thresh;  // p * 2^24; 0 = no dropout
  float scale;      // 1 / (1 - p)
  unsigned site;
  __device__ __forceinline__ float apply(float v, unsigned long long i) const {
    if (thresh == 0) return v;
    return hash24(seed, site, i) >= thresh ? v * scale : 0.f;
  }
};

// ---------------------------------------------------------------- GEMM
// C[i][j] = sum_k A(i,k) B(k,j), i < M, j < N, k in [k0, k1).
//   A_IC = false: A(i,k) = a[i*lda + k]  (k contiguous)      A_IC = true: A(i,k) = a[k*lda + i]  (i contiguous)
//   B_JC = true : B(k,j) = b[k*ldb + j]  (j contiguous)      B_JC = false: B(k,j) = b[j*ldb + k] (k contiguous)
// All leading dimensions and the contiguous extents are multiples of 4 (checked on the host).
constexpr int TM = 128, TN = 128, TK = 16, GT = 256;

enum { EPI_STORE = 0, EPI_BIAS, EPI_BIAS_RELU_DROP, EPI_ACC, EPI_MASK, EPI_PARTIAL };
struct Epi {
  int mode;
  float* c;
  int ldc;
  const float* bias;   // EPI_BIAS*, per column
  const float* mask;   // EPI_MASK: dH = (Hd > 0) ? v * scale : 0, same layout as c
  float scale;
  Drop drop;           // EPI_BIAS_RELU_DROP
  size_t split_stride;  // EPI_PARTIAL: c + blockIdx.z * split_stride
};

template <bool A_IC, bool B_JC>
__global__ void __launch_bounds__(GT) k_sgemm(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                              int M, int N, int Kd, int ksplit, Epi E) {
  __shared__ __align__(16) float As[2][TK][TM + 4];
  __shared__ __align__(16) float Bs[2][TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int kbeg = blockIdx.z * ksplit, kend = min(Kd, kbeg + ksplit);
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // staging: 128 x 16 elements per operand = 512 float4, 2 per thread
  float4 pa[2], pb[2];
  auto ld4 = [&](const float* p, bool ok) -> float4 {
    return ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto fetch = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int f = tid + h * GT;
      if (A_IC) {  // float4 along i: f -> (k = f / 32, i4 = (f % 32) * 4)
        const int k = k0 + (f >> 5), i = m0 + (f & 31) * 4;
        pa[h] = ld4(a + (size_t)k * lda + i, k < kend && i < M);
      } else {  // float4 along k: f -> (i = f / 4, k4 = (f % 4) * 4)
        const int i = m0 + (f >> 2), k = k0 + (f & 3) * 4;
        pa[h] = ld4(a + (size_t)i * lda + k, i < M && k < kend);
      }
      if (B_JC) {
        const int k = k0 + (f >> 5), j = n0 + (f & 31) * 4;
        pb[h] = ld4(b + (size_t)k * ldb + j, k < kend && j < N);
      } else {
        const int j = n0 + (f >> 2), k = k0 + (f & 3) * 4;
        pb[h] = ld4(b + (size_t)j * ldb + k, j < N && k < kend);
      }
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int f = tid + h * GT;
      if (A_IC) {
        *reinterpret_cast<float4*>(&As[buf][f >> 5][(f & 31) * 4]) = pa[h];
      } else {
        const int i = f >> 2, k = (f & 3) * 4;
        As[buf][k][i] = pa[h].x; As[buf][k + 1][i] = pa[h].y; As[buf][k + 2][i] = pa[h].z; As[buf][k + 3][i] = pa[h].w;
      }
      if (B_JC) {
        *reinterpret_cast<float4*>(&Bs[buf][f >> 5][(f & 31) * 4]) = pb[h];
      } else {
        const int j = f >> 2, k = (f & 3) * 4;
        Bs[buf][k][j] = pb[h].x; Bs[buf][k + 1][j] = pb[h].y; Bs[buf][k + 2][j] = pb[h].z; Bs[buf][k + 3][j] = pb[h].w;
      }
    }
  };

  // the k range of a split is a multiple of 4 wide except possibly the last; partial float4s never straddle kend because
  // Kd and ksplit are multiples of 4 (host check)
  const int nk = (kend - kbeg + TK - 1) / TK;
  if (nk > 0) {
    fetch(kbeg);
    stage(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch(kbeg + (kt + 1) * TK);
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      stage(buf ^ 1);
      __syncthreads();
    }
  }

  float* cbase = E.c + (E.mode == EPI_PARTIAL ? (size_t)blockIdx.z * E.split_stride : 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int n = n0 + half * 64 + tx * 4;
      if (n >= N) continue;
      float v[4] = {acc[i][half * 4], acc[i][half * 4 + 1], acc[i][half * 4 + 2], acc[i][half * 4 + 3]};
      float* dst = cbase + (size_t)m * E.ldc + n;
      if (E.mode == EPI_BIAS || E.mode == EPI_BIAS_RELU_DROP) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(E.bias + n));
        v[0] += b4.x; v[1] += b4.y; v[2] += b4.z; v[3] += b4.w;
        if (E.mode == EPI_BIAS_RELU_DROP) {
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = E.drop.apply(fmaxf(v[q], 0.f), (unsigned long long)m * E.ldc + n + q);
        }
      } else if (E.mode == EPI_ACC) {
        const float4 o = *reinterpret_cast<const float4*>(dst);
        v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
      } else if (E.mode == EPI_MASK) {
        const float4 h = __ldg(reinterpret_cast<const float4*>(E.mask + (size_t)m * E.ldc + n));
        v[0] = h.x > 0.f ? v[0] * E.scale : 0.f;
        v[1] = h.y > 0.f ? v[1] * E.scale : 0.f;
        v[2] = h.z > 0.f ? v[2] * E.scale : 0.f;
        v[3] = h.w > 0.f ? v[3] * E.scale : 0.f;
      }
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// out[i] = sum_s part[s][i]  (i < n), optionally out[i] += ...
__global__ void k_reduce_parts(const float* __restrict__ part, int splits, size_t n, float* __restrict__ out, int accumulate) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < splits; ++p) s += part[(size_t)p * n + i];
  out[i] = accumulate ? out[i] + s : s;
}

// column sums of X[M][N] over rows: partial sums per row chunk, then k_reduce_parts.
__global__ void __launch_bounds__(256) k_colsum_part(const float* __restrict__ x, int M, int N, int rows_per_block,
                                                     float* __restrict__ part) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= N) return;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += x[(size_t)r * N + c];
  part[(size_t)blockIdx.y * N + c] = s;
}

struct Ctx {
  cudaStream_t st;
  float* scratch;       // split-K partials / column-sum partials
  size_t scratch_floats;
};

int gemm(const Ctx& cx, bool a_ic, bool b_jc, const float* a, int lda, const float* b, int ldb, int M, int N, int Kd, Epi E,
         int splits = 1) {
  if (M <= 0 || N <= 0 || Kd <= 0) return S3D_OK;
  if ((lda & 3) || (ldb & 3) || (E.ldc & 3) || (N & 3) || (Kd & 3) || (a_ic && (M & 3))) {
    set_error("train decoder: GEMM extents must be multiples of 4");
    return S3D_ERR_BAD_ARG;
  }
  int ksplit = Kd;
  if (splits > 1) {
    ksplit = ((Kd + splits - 1) / splits + 15) / 16 * 16;
    splits = (Kd + ksplit - 1) / ksplit;
  }
  dim3 grid((M + TM - 1) / TM, (N + TN - 1) / TN, splits);
  if (a_ic && b_jc) k_sgemm<true, true><<<grid, GT, 0, cx.st>>>(a, lda, b, ldb, M, N, Kd, ksplit, E);
  else if (!a_ic && b_jc) k_sgemm<false, true><<<grid, GT, 0, cx.st>>>(a, lda, b, ldb, M, N, Kd, ksplit, E);
  else if (!a_ic && !b_jc) k_sgemm<false, false><<<grid, GT, 0, cx.st>>>(a, lda, b, ldb, M, N, Kd, ksplit, E);
  else k_sgemm<true, false><<<grid, GT, 0, cx.st>>>(a, lda, b, ldb, M, N, Kd, ksplit, E);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

// Y[R,N] = X[R,K] . W[N,K]^T + bias  (optionally ReLU + dropout)
int linear_fwd(const Ctx& cx, const float* x, const float* w, const float* bias, int R, int N, int Kd, float* y,
               const Drop* relu_drop) {
  Epi E{};
  E.mode = relu_drop ? EPI_BIAS_RELU_DROP : EPI_BIAS;
  E.c = y; E.ldc = N; E.bias = bias;
  if (relu_drop) E.drop = *relu_drop;
  return gemm(cx, false, false, x, Kd, w, Kd, R, N, Kd, E);
}
// dX[R,K] (=, +=, masked) dY[R,N] . W[N,K]
int linear_bwd_x(const Ctx& cx, const float* dy, const float* w, int R, int N, int Kd, float* dx, int mode,
                 const float* mask = nullptr, float scale = 1.f) {
  Epi E{};
  E.mode = mode; E.c = dx; E.ldc = Kd; E.mask = mask; E.scale = scale;
  return gemm(cx, false, true, dy, N, w, Kd, R, Kd, N, E);
}
// dW[N,K] = dY[R,N]^T . X[R,K] ; db[N] = column sums of dY.   Split over R, deterministic two-stage reduction.
int linear_bwd_w(const Ctx& cx, const float* dy, const float* x, int R, int N, int Kd, float* dw, float* db) {
  const int tiles = ((N + TM - 1) / TM) * ((Kd + TN - 1) / TN);
  int splits = (296 + tiles - 1) / tiles;  // ~2 CTAs per SM
  const int max_by_rows = (R + 255) / 256;
  if (splits > max_by_rows) splits = max_by_rows;
  if (splits < 1) splits = 1;
  const size_t n = (size_t)N * Kd;
  while (splits > 1 && (size_t)splits * n > cx.scratch_floats) --splits;
  int ksplit = ((R + splits - 1) / splits + 15) / 16 * 16;
  splits = (R + ksplit - 1) / ksplit;
  Epi E{};
  E.ldc = Kd;
  if (splits == 1) {
    E.mode = EPI_STORE; E.c = dw;
    S3D_TRY(gemm(cx, true, true, dy, N, x, Kd, N, Kd, R, E));
  } else {
    E.mode = EPI_PARTIAL; E.c = cx.scratch; E.split_stride = n;
    S3D_TRY(gemm(cx, true, true, dy, N, x, Kd, N, Kd, R, E, splits));
    k_reduce_parts<<<(unsigned)((n + 255) / 256), 256, 0, cx.st>>>(cx.scratch, splits, n, dw, 0);
    S3D_LAUNCH_CHECK();
  }
  if (db) {
    int rpb = 256;
    int nb = (R + rpb - 1) / rpb;
    while ((size_t)nb * N > cx.scratch_floats) { rpb *= 2; nb = (R + rpb - 1) / rpb; }
    k_colsum_part<<<dim3((N + 255) / 256, nb), 256, 0, cx.st>>>(dy, R, N, rpb, cx.scratch);
    S3D_LAUNCH_CHECK();
    k_reduce_parts<<<(N + 255) / 256, 256, 0, cx.st>>>(cx.scratch, nb, N, db, 0);
    S3D_LAUNCH_CHECK();
  }
  return S3D_OK;
}

// ---------------------------------------------------------------- sampling (models.py:28-46, 69-78)
struct Feats {
  const float* p[5];
  float* g[5];  // gradients (backward)
  int C[5], R[5];
};
__device__ __forceinline__ void project(const float* T, float x, float y, float z, float& gu, float& gv) {
  const float pu = x * T[0] + y * T[3] + z * T[6] + T[9];
  const float pv = x * T[1] + y * T[4] + z * T[7] + T[10];
  const float pw = x * T[2] + y * T[5] + z * T[8] + T[11];
  gu = fminf(fmaxf(2.f * (pu / pw - 0.5f), -1.f), 1.f);
  gv = fminf(fmaxf(2.f * (pv / pw - 0.5f), -1.f), 1.f);
}
// One block per (query q = b*Mq + m, slice k): AGG[(q*K + k)][992] = cat_s grid_sample(feat_s[b*K + k], uv(q)).
// BWD: dfeat_s[b*K + k][c][tap] += w_tap * dAGG[...][c]  (atomics: several queries share texels).
template <bool BWD>
__global__ void __launch_bounds__(256) k_sample(Feats F, const float* __restrict__ qry, const float* __restrict__ T, int Mq,
                                                int K, float* __restrict__ agg) {
  __shared__ Taps taps[5];
  const int row = blockIdx.x;  // q*K + k
  const int q = row / K, k = row - q * K;
  const int b = q / Mq;
  if (threadIdx.x < 5) {
    float gu, gv;
    project(T + 12 * b, qry[3 * q], qry[3 * q + 1], qry[3 * q + 2], gu, gv);
    taps[threadIdx.x] = make_taps(gu, gv, F.R[threadIdx.x]);
  }
  __syncthreads();
  float* arow = agg + (size_t)row * CAGG;
  for (int c = threadIdx.x; c < CAGG; c += 256) {
    int s = 0, c0 = 0;
    while (c >= c0 + F.C[s]) { c0 += F.C[s]; ++s; }
    const Taps t = taps[s];
    const size_t r2 = (size_t)F.R[s] * F.R[s];
    const size_t base = ((size_t)(b * K + k) * F.C[s] + (c - c0)) * r2;
    if (!BWD) {
      const float* P = F.p[s] + base;
      arow[c] = __ldg(P + t.o00) * t.w00 + __ldg(P + t.o01) * t.w01 + __ldg(P + t.o10) * t.w10 + __ldg(P + t.o11) * t.w11;
    } else {
      float* G = F.g[s] + base;
      const float d = arow[c];
      if (t.w00 != 0.f) atomicAdd(G + t.o00, d * t.w00);
      if (t.w01 != 0.f) atomicAdd(G + t.o01, d * t.w01);
      if (t.w10 != 0.f) atomicAdd(G + t.o10, d * t.w10);
      if (t.w11 != 0.f) atomicAdd(G + t.o11, d * t.w11);
    }
  }
}

// X[q*L + 0] = fc_p(qry[q])  (models.py:79); slice tokens are written by the fc_s GEMM into a compact [Q*K][128]
// buffer and scattered to rows q*L + 1 + k here.
__global__ void __launch_bounds__(128) k_tokens_assemble(const float* __restrict__ qry, const float* __restrict__ wp,
                                                         const float* __restrict__ bp, const float* __restrict__ toks, int K,
                                                         float* __restrict__ X) {
  const int q = blockIdx.x, c = threadIdx.x, L = K + 1;
  const float x = qry[3 * q], y = qry[3 * q + 1], z = qry[3 * q + 2];
  X[((size_t)q * L) * D + c] = bp[c] + x * wp[3 * c] + y * wp[3 * c + 1] + z * wp[3 * c + 2];
  for (int k = 0; k < K; ++k) X[((size_t)q * L + 1 + k) * D + c] = toks[((size_t)q * K + k) * D + c];
}
// backward of the assembly: dtoks <- rows 1..K of dX; fc_p gradients from rows 0 (per-block partials over queries)
__global__ void __launch_bounds__(128) k_tokens_split(const float* __restrict__ dX, int K, float* __restrict__ dtoks) {
  const int q = blockIdx.x, c = threadIdx.x, L = K + 1;
  for (int k = 0; k < K; ++k) dtoks[((size_t)q * K + k) * D + c] = dX[((size_t)q * L + 1 + k) * D + c];
}
// dWp[c][j] = sum_q dX[q*L][c] * qry[q][j], dbp[c] = sum_q dX[q*L][c]: grid = chunks of queries, partials [chunk][4][128]
__global__ void __launch_bounds__(128) k_fcp_grad_part(const float* __restrict__ dX, const float* __restrict__ qry, int Q, int L,
                                                       int q_per_block, float* __restrict__ part) {
  const int c = threadIdx.x;
  const int q0 = blockIdx.x * q_per_block, q1 = min(Q, q0 + q_per_block);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, sb = 0.f;
  for (int q = q0; q < q1; ++q) {
    const float d = dX[((size_t)q * L) * D + c];
    s0 = fmaf(d, qry[3 * q], s0);
    s1 = fmaf(d, qry[3 * q + 1], s1);
    s2 = fmaf(d, qry[3 * q + 2], s2);
    sb += d;
  }
  float* p = part + (size_t)blockIdx.x * 512;
  p[3 * c] = s0; p[3 * c + 1] = s1; p[3 * c + 2] = s2; p[384 + c] = sb;
}

// ---------------------------------------------------------------- attention (nn.MultiheadAttention, 4 heads of 32)
// One block (128 threads) per query.  QKV [Q*L][384]; P [Q][4][L][L] = softmax probabilities (saved); O [Q*L][128].
__global__ void __launch_bounds__(128) k_attn_fwd(const float* __restrict__ QKV, float* __restrict__ P, float* __restrict__ O,
                                                  int L, Drop drop) {
  extern __shared__ float sm[];
  float* qkv = sm;                    // [L][384]
  float* sc = sm + (size_t)L * 384;   // [4][L][L]
  const int i = blockIdx.x;
  const float* src = QKV + (size_t)i * L * 384;
  for (int t = threadIdx.x; t < L * 96; t += 128)
    reinterpret_cast<float4*>(qkv)[t] = __ldg(reinterpret_cast<const float4*>(src) + t);
  __syncthreads();
  const float scale = 0.17677669529663687f;  // 1/sqrt(32)
  for (int t = threadIdx.x; t < NH * L * L; t += 128) {
    const int h = t / (L * L), r = t % (L * L), a = r / L, b = r % L;
    const float* qa = qkv + a * 384 + h * HD;
    const float* kb = qkv + b * 384 + D + h * HD;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < HD; ++c) s = fmaf(qa[c] * scale, kb[c], s);
    sc[t] = s;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < NH * L; t += 128) {
    float* row = sc + (size_t)t * L;
    float mx = row[0];
    for (int b = 1; b < L; ++b) mx = fmaxf(mx, row[b]);
    float sum = 0.f;
    for (int b = 0; b < L; ++b) {
      const float e = expf(row[b] - mx);
      row[b] = e;
      sum += e;
    }
    const float inv = 1.f / sum;
    float* prow = P + ((size_t)i * NH * L + t) * L;
    for (int b = 0; b < L; ++b) {
      const float pv = row[b] * inv;
      prow[b] = pv;  // saved: post-softmax, pre-dropout
      row[b] = drop.apply(pv, ((unsigned long long)i * NH * L + t) * L + b);
    }
  }
  __syncthreads();
  const int c = threadIdx.x, h = c >> 5;
  for (int a = 0; a < L; ++a) {
    const float* p = sc + ((size_t)h * L + a) * L;
    float o = 0.f;
    for (int b = 0; b < L; ++b) o = fmaf(p[b], qkv[b * 384 + 2 * D + c], o);
    O[((size_t)i * L + a) * D + c] = o;
  }
}

// dQKV from dO: dPd = dO V^T, dV = Pd^T dO, dP = mask(dPd), dS = P (dP - sum_j dP P), dQ = dS K / sqrt(32), dK = dS^T Q / sqrt(32)
__global__ void __launch_bounds__(128) k_attn_bwd(const float* __restrict__ QKV, const float* __restrict__ P,
                                                  const float* __restrict__ dO, float* __restrict__ dQKV, int L, Drop drop) {
  extern __shared__ float sm[];
  float* qkv = sm;                         // [L][384]
  float* dov = qkv + (size_t)L * 384;      // [L][128]
  float* pd = dov + (size_t)L * D;         // [4][L][L] dropped probabilities, then reused
  float* ds = pd + (size_t)NH * L * L;     // [4][L][L] dS
  const int i = blockIdx.x;
  const float* src = QKV + (size_t)i * L * 384;
  for (int t = threadIdx.x; t < L * 96; t += 128)
    reinterpret_cast<float4*>(qkv)[t] = __ldg(reinterpret_cast<const float4*>(src) + t);
  for (int t = threadIdx.x; t < L * 32; t += 128)
    reinterpret_cast<float4*>(dov)[t] = __ldg(reinterpret_cast<const float4*>(dO + (size_t)i * L * D) + t);
  const float* Pq = P + (size_t)i * NH * L * L;
  for (int t = threadIdx.x; t < NH * L * L; t += 128)
    pd[t] = drop.apply(Pq[t], (unsigned long long)i * NH * L * L + t);
  __syncthreads();
  // dV[b][c] = sum_a Pd[h][a][b] dO[a][c]
  {
    const int c = threadIdx.x, h = c >> 5;
    for (int b = 0; b < L; ++b) {
      float s = 0.f;
      for (int a = 0; a < L; ++a) s = fmaf(pd[(h * L + a) * L + b], dov[a * D + c], s);
      dQKV[((size_t)i * L + b) * 384 + 2 * D + c] = s;
    }
  }
  // dPd[h][a][b] = sum_c dO[a][h*32+c] V[b][h*32+c] ; through dropout ; softmax backward
  for (int t = threadIdx.x; t < NH * L * L; t += 128) {
    const int h = t / (L * L), r = t % (L * L), a = r / L, b = r % L;
    const float* da = dov + a * D + h * HD;
    const float* vb = qkv + b * 384 + 2 * D + h * HD;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < HD; ++c) s = fmaf(da[c], vb[c], s);
    // d(dropout): keep -> scale, dropped -> 0 (same hash as forward)
    float g = s;
    if (drop.thresh) g = hash24(drop.seed, drop.site, (unsigned long long)i * NH * L * L + t) >= drop.thresh ? s * drop.scale : 0.f;
    ds[t] = g;  // dP
  }
  __syncthreads();
  for (int t = threadIdx.x; t < NH * L; t += 128) {
    const float* prow = Pq + (size_t)t * L;
    float* drow = ds + (size_t)t * L;
    float dot = 0.f;
    for (int b = 0; b < L; ++b) dot = fmaf(drow[b], prow[b], dot);
    for (int b = 0; b < L; ++b) drow[b] = prow[b] * (drow[b] - dot);
  }
  __syncthreads();
  const float scale = 0.17677669529663687f;
  {
    const int c = threadIdx.x, h = c >> 5;
    for (int a = 0; a < L; ++a) {  // dQ[a][c] = scale * sum_b dS[h][a][b] K[b][c]
      float s = 0.f;
      for (int b = 0; b < L; ++b) s = fmaf(ds[(h * L + a) * L + b], qkv[b * 384 + D + c], s);
      dQKV[((size_t)i * L + a) * 384 + c] = s * scale;
    }
    for (int b = 0; b < L; ++b) {  // dK[b][c] = scale * sum_a dS[h][a][b] Q[a][c]
      float s = 0.f;
      for (int a = 0; a < L; ++a) s = fmaf(ds[(h * L + a) * L + b], qkv[a * 384 + c], s);
      dQKV[((size_t)i * L + b) * 384 + D + c] = s * scale;
    }
  }
}

// ---------------------------------------------------------------- residual + dropout + LayerNorm
// pre = X + drop(Y); out = LN(pre) * w + b.  One warp per row; `pre` is saved for the backward pass.
__global__ void __launch_bounds__(256) k_add_ln_fwd(const float* __restrict__ X, const float* __restrict__ Y, Drop drop,
                                                    const float* __restrict__ w, const float* __restrict__ b, long long rows,
                                                    float* __restrict__ pre, float* __restrict__ out) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4 a = *reinterpret_cast<const float4*>(X + row * D + lane * 4);
  const float4 t = *reinterpret_cast<const float4*>(Y + row * D + lane * 4);
  const unsigned long long e0 = (unsigned long long)row * D + lane * 4;
  float v[4] = {a.x + drop.apply(t.x, e0), a.y + drop.apply(t.y, e0 + 1), a.z + drop.apply(t.z, e0 + 2),
                a.w + drop.apply(t.w, e0 + 3)};
  *reinterpret_cast<float4*>(pre + row * D + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
  float s = v[0] + v[1] + v[2] + v[3];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / D);
  float d2 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j] -= mean;
    d2 += v[j] * v[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
  const float rstd = rsqrtf(d2 * (1.f / D) + 1e-5f);
  const float4 ww = *reinterpret_cast<const float4*>(w + lane * 4);
  const float4 bb = *reinterpret_cast<const float4*>(b + lane * 4);
  *reinterpret_cast<float4*>(out + row * D + lane * 4) =
      make_float4(v[0] * rstd * ww.x + bb.x, v[1] * rstd * ww.y + bb.y, v[2] * rstd * ww.z + bb.z, v[3] * rstd * ww.w + bb.w);
}
// Given dOut (gradient of the LayerNorm output) and the saved pre-norm sum: dPre (= gradient of the residual input X),
// dY = dropout-masked dPre (gradient of the branch output), and per-block partial sums of dw, db.
// Each block loops over a strided set of rows (8 warps); partials [block][2][128].
__global__ void __launch_bounds__(256) k_add_ln_bwd(const float* __restrict__ dOut, const float* __restrict__ pre, Drop drop,
                                                    const float* __restrict__ w, long long rows, float* __restrict__ dPre,
                                                    float* __restrict__ dY, float* __restrict__ part) {
  __shared__ float red[8][2][D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 ww = *reinterpret_cast<const float4*>(w + lane * 4);
  float gw[4] = {0.f, 0.f, 0.f, 0.f}, gb[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
    const float4 p4 = *reinterpret_cast<const float4*>(pre + row * D + lane * 4);
    const float4 d4 = *reinterpret_cast<const float4*>(dOut + row * D + lane * 4);
    float v[4] = {p4.x, p4.y, p4.z, p4.w};
    const float dout4[4] = {d4.x, d4.y, d4.z, d4.w};
    float s = v[0] + v[1] + v[2] + v[3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / D);
    float d2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] -= mean;
      d2 += v[j] * v[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    const float rstd = rsqrtf(d2 * (1.f / D) + 1e-5f);
    const float wv[4] = {ww.x, ww.y, ww.z, ww.w};
    float g[4], xh[4], sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      xh[j] = v[j] * rstd;
      g[j] = dout4[j] * wv[j];
      sg += g[j];
      sgx += g[j] * xh[j];
      gw[j] += dout4[j] * xh[j];
      gb[j] += dout4[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
    }
    const float mg = sg * (1.f / D), mgx = sgx * (1.f / D);
    float dp[4], dy[4];
    const unsigned long long e0 = (unsigned long long)row * D + lane * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dp[j] = rstd * (g[j] - mg - xh[j] * mgx);
      dy[j] = drop.apply(dp[j], e0 + j);  // same mask and scale as the forward dropout of the branch
    }
    *reinterpret_cast<float4*>(dPre + row * D + lane * 4) = make_float4(dp[0], dp[1], dp[2], dp[3]);
    *reinterpret_cast<float4*>(dY + row * D + lane * 4) = make_float4(dy[0], dy[1], dy[2], dy[3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[warp][0][lane * 4 + j] = gw[j];
    red[warp][1][lane * 4 + j] = gb[j];
  }
  __syncthreads();
  const int t = threadIdx.x;  // 256 threads: (which, column)
  float s = 0.f;
#pragma unroll
  for (int wq = 0; wq < 8; ++wq) s += red[wq][t >> 7][t & 127];
  part[(size_t)blockIdx.x * 256 + t] = s;
}

// out[a] += b[a]
__global__ void k_axpy(float* __restrict__ a, const float* __restrict__ b, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += b[i];
}

// ---------------------------------------------------------------- head (models.py:83-84)
__global__ void __launch_bounds__(256) k_head_fwd(const float* __restrict__ X, const float* __restrict__ w,
                                                  const float* __restrict__ b, float* __restrict__ out, int Q, int L) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= Q) return;
  const int lane = threadIdx.x & 31;
  const float4 a = *reinterpret_cast<const float4*>(X + (size_t)i * L * D + lane * 4);
  const float4 ww = *reinterpret_cast<const float4*>(w + lane * 4);
  float s = a.x * ww.x + a.y * ww.y + a.z * ww.z + a.w * ww.w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[i] = s + b[0];
}
// dX[(q, 0)] = dsdf[q] * w, other token rows 0; per-block partials of dw (128) and db (1): [block][129]
__global__ void __launch_bounds__(128) k_head_bwd(const float* __restrict__ X, const float* __restrict__ w,
                                                  const float* __restrict__ dsdf, int Q, int L, int q_per_block,
                                                  float* __restrict__ dX, float* __restrict__ part) {
  const int c = threadIdx.x;
  const int q0 = blockIdx.x * q_per_block, q1 = min(Q, q0 + q_per_block);
  const float wc = w[c];
  float sw = 0.f, sb = 0.f;
  for (int q = q0; q < q1; ++q) {
    const float d = dsdf[q];
    dX[((size_t)q * L) * D + c] = d * wc;
    for (int t = 1; t < L; ++t) dX[((size_t)q * L + t) * D + c] = 0.f;
    sw = fmaf(d, X[((size_t)q * L) * D + c], sw);
    sb += d;
  }
  part[(size_t)blockIdx.x * 129 + c] = sw;
  if (c == 0) part[(size_t)blockIdx.x * 129 + 128] = sb;
}

// ---------------------------------------------------------------- buffers
// Saved activations (forward -> backward), per layer, in floats; R = Q * L token rows.
struct Saved {
  float *agg, *x0;
  struct Layer {
    float *qkv, *p, *o, *pre1, *x1, *hd, *pre2, *x2;
  } l[3];
};
size_t carve(Saved& s, float* base, long long Q, int K) {
  const long long L = K + 1, R = Q * L;
  size_t off = 0;
  auto take = [&](size_t n) {
    float* p = base ? base + off : nullptr;
    off += (n + 63) / 64 * 64;
    return p;
  };
  s.agg = take((size_t)Q * K * CAGG);
  s.x0 = take((size_t)R * D);
  for (int i = 0; i < 3; ++i) {
    s.l[i].qkv = take((size_t)R * 3 * D);
    s.l[i].p = take((size_t)Q * NH * L * L);
    s.l[i].o = take((size_t)R * D);
    s.l[i].pre1 = take((size_t)R * D);
    s.l[i].x1 = take((size_t)R * D);
    s.l[i].hd = take((size_t)R * FF);
    s.l[i].pre2 = take((size_t)R * D);
    s.l[i].x2 = take((size_t)R * D);
  }
  return off;
}
constexpr size_t SCRATCH_FLOATS = (size_t)16 << 20;  // 64 MB of split-K / reduction partials

struct Params {  // the 42 tensors, PyTorch layouts
  const float *wp, *bp, *ws, *bs;
  struct Layer {
    const float *win, *bin, *wo, *bo, *w1, *b1, *w2, *b2, *n1w, *n1b, *n2w, *n2b;
  } l[3];
  const float *wout, *bout;
};
template <class P, class T>
void unpack(P& p, T* const* a) {
  int i = 0;
  p.wp = a[i++]; p.bp = a[i++]; p.ws = a[i++]; p.bs = a[i++];
  for (int l = 0; l < 3; ++l) {
    p.l[l].win = a[i++]; p.l[l].bin = a[i++]; p.l[l].wo = a[i++]; p.l[l].bo = a[i++];
    p.l[l].w1 = a[i++]; p.l[l].b1 = a[i++]; p.l[l].w2 = a[i++]; p.l[l].b2 = a[i++];
    p.l[l].n1w = a[i++]; p.l[l].n1b = a[i++]; p.l[l].n2w = a[i++]; p.l[l].n2b = a[i++];
  }
  p.wout = a[i++]; p.bout = a[i++];
}
struct GradParams {
  float *wp, *bp, *ws, *bs;
  struct Layer {
    float *win, *bin, *wo, *bo, *w1, *b1, *w2, *b2, *n1w, *n1b, *n2w, *n2b;
  } l[3];
  float *wout, *bout;
};

Drop make_drop(float p, unsigned long long seed, unsigned site) {
  Drop d{};
  d.seed = seed;
  d.site = site;
  if (p > 0.f) {
    d.thresh = (uint32_t)(p * 16777216.0f);
    d.scale = 1.f / (1.f - p);
  } else {
    d.thresh = 0;
    d.scale = 1.f;
  }
  return d;
}

int check_cfg(const s3d_train_cfg* c) {
  if (!c || c->B < 1 || c->n_qry < 1 || c->K < 1 || c->K > 12 || c->S < 32 || c->S % 16 || c->dropout_p < 0.f ||
      c->dropout_p >= 1.f) {
    set_error("train decoder: bad configuration");
    return S3D_ERR_BAD_ARG;
  }
  return S3D_OK;
}
void fill_feats(Feats& F, const s3d_train_cfg* c, const float* const* feats, float* const* grads) {
  for (int s = 0; s < 5; ++s) {
    F.p[s] = feats ? feats[s] : nullptr;
    F.g[s] = grads ? grads[s] : nullptr;
    F.C[s] = kPlaneC[s];
    F.R[s] = plane_res(c->S, s);
  }
}

}  // namespace

size_t train_decoder_saved_bytes(const s3d_train_cfg* c) {
  if (check_cfg(c) != S3D_OK) return 0;
  Saved s;
  return (carve(s, nullptr, (long long)c->B * c->n_qry, c->K) + SCRATCH_FLOATS) * sizeof(float);
}

size_t train_decoder_bwd_workspace_bytes(const s3d_train_cfg* c) {
  if (check_cfg(c) != S3D_OK) return 0;
  const size_t R = (size_t)c->B * c->n_qry * (c->K + 1);
  // dX, dPre, dBranch (128 each), dQKV (384), dH (2048), dtoks (128 per slice row) + dAGG (992 per slice row)
  return (R * (3 * D + 3 * D + FF) + (size_t)c->B * c->n_qry * c->K * (D + CAGG) + 1024) * sizeof(float);
}

int train_decoder_fwd(const s3d_train_cfg* c, const float* const* feats, const float* qry, const float* T,
                      const float* const* params, float* sdf, void* saved, size_t saved_bytes, cudaStream_t st) {
  S3D_TRY(check_cfg(c));
  if (!feats || !qry || !T || !params || !sdf || !saved || saved_bytes < train_decoder_saved_bytes(c)) {
    set_error("train decoder fwd: null pointer or saved buffer too small");
    return S3D_ERR_BAD_ARG;
  }
  const int K = c->K, L = K + 1;
  const long long Q = (long long)c->B * c->n_qry, R = Q * L;
  Saved s;
  const size_t used = carve(s, static_cast<float*>(saved), Q, K);
  Ctx cx{st, static_cast<float*>(saved) + used, SCRATCH_FLOATS};
  Params P;
  unpack(P, params);
  Feats F;
  fill_feats(F, c, feats, nullptr);
  const float p = c->dropout_p;

  k_sample<false><<<(unsigned)(Q * K), 256, 0, st>>>(F, qry, T, c->n_qry, K, s.agg);
  S3D_LAUNCH_CHECK();
  float* toks = s.l[0].o;  // free until layer 0's attention output is written (after the assembly)
  S3D_TRY(linear_fwd(cx, s.agg, P.ws, P.bs, (int)(Q * K), D, CAGG, toks, nullptr));
  k_tokens_assemble<<<(unsigned)Q, 128, 0, st>>>(qry, P.wp, P.bp, toks, K, s.x0);
  S3D_LAUNCH_CHECK();
  const float* x = s.x0;
  const size_t attn_smem = ((size_t)L * 384 + NH * L * L) * sizeof(float);
  float* branch = cx.scratch;  // [R][128] branch outputs (out-proj / linear2) live only until the add+LN
  Ctx cx2{st, cx.scratch + (size_t)R * D, SCRATCH_FLOATS - (size_t)R * D};
  if ((size_t)R * D > SCRATCH_FLOATS / 2) {
    set_error("train decoder: batch too large for the scratch area");
    return S3D_ERR_WORKSPACE;
  }
  for (int l = 0; l < 3; ++l) {
    const Params::Layer& W = P.l[l];
    const Saved::Layer& A = s.l[l];
    S3D_TRY(linear_fwd(cx2, x, W.win, W.bin, (int)R, 3 * D, D, A.qkv, nullptr));
    k_attn_fwd<<<(unsigned)Q, 128, attn_smem, st>>>(A.qkv, A.p, A.o, L, make_drop(p, c->seed, 4 * l + 0));
    S3D_LAUNCH_CHECK();
    S3D_TRY(linear_fwd(cx2, A.o, W.wo, W.bo, (int)R, D, D, branch, nullptr));
    k_add_ln_fwd<<<(unsigned)((R + 7) / 8), 256, 0, st>>>(x, branch, make_drop(p, c->seed, 4 * l + 1), W.n1w, W.n1b, R, A.pre1,
                                                          A.x1);
    S3D_LAUNCH_CHECK();
    const Drop dffn = make_drop(p, c->seed, 4 * l + 2);
    S3D_TRY(linear_fwd(cx2, A.x1, W.w1, W.b1, (int)R, FF, D, A.hd, &dffn));
    S3D_TRY(linear_fwd(cx2, A.hd, W.w2, W.b2, (int)R, D, FF, branch, nullptr));
    k_add_ln_fwd<<<(unsigned)((R + 7) / 8), 256, 0, st>>>(A.x1, branch, make_drop(p, c->seed, 4 * l + 3), W.n2w, W.n2b, R, A.pre2,
                                                          A.x2);
    S3D_LAUNCH_CHECK();
    x = A.x2;
  }
  k_head_fwd<<<(unsigned)((Q + 7) / 8), 256, 0, st>>>(x, P.wout, P.bout, sdf, (int)Q, L);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

int train_decoder_bwd(const s3d_train_cfg* c, const float* qry, const float* T, const float* const* params, const float* dsdf,
                      void* saved, size_t saved_bytes, float* const* dfeats, float* const* dparams, void* ws, size_t ws_bytes,
                      cudaStream_t st) {
  S3D_TRY(check_cfg(c));
  if (!qry || !T || !params || !dsdf || !saved || !dfeats || !dparams || !ws ||
      saved_bytes < train_decoder_saved_bytes(c) || ws_bytes < train_decoder_bwd_workspace_bytes(c)) {
    set_error("train decoder bwd: null pointer or buffer too small");
    return S3D_ERR_BAD_ARG;
  }
  const int K = c->K, L = K + 1;
  const long long Q = (long long)c->B * c->n_qry, R = Q * L;
  Saved s;
  const size_t used = carve(s, static_cast<float*>(saved), Q, K);
  Ctx cx{st, static_cast<float*>(saved) + used, SCRATCH_FLOATS};
  Params P;
  unpack(P, params);
  GradParams G;
  unpack(G, dparams);
  Feats F;
  fill_feats(F, c, nullptr, dfeats);
  const float p = c->dropout_p;
  float* w = static_cast<float*>(ws);
  float* dX = w;                      w += (size_t)R * D;
  float* dPre = w;                    w += (size_t)R * D;
  float* dBr = w;                     w += (size_t)R * D;
  float* dQKV = w;                    w += (size_t)R * 3 * D;
  float* dH = w;                      w += (size_t)R * FF;
  float* dtoks = w;                   w += (size_t)Q * K * D;
  float* dAGG = w;

  const int QPB = 64;
  const int nqb = (int)((Q + QPB - 1) / QPB);
  // ---- head
  const float* xlast = s.l[2].x2;
  k_head_bwd<<<nqb, 128, 0, st>>>(xlast, P.wout, dsdf, (int)Q, L, QPB, dX, cx.scratch);
  S3D_LAUNCH_CHECK();
  {
    // partial layout [block][129]: dw = first 128, db = last -> reduce with stride 129
    k_reduce_parts<<<1, 256, 0, st>>>(cx.scratch, nqb, 129, cx.scratch + (size_t)nqb * 129, 0);
    S3D_LAUNCH_CHECK();
    S3D_CUDA(cudaMemcpyAsync(G.wout, cx.scratch + (size_t)nqb * 129, 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    S3D_CUDA(cudaMemcpyAsync(G.bout, cx.scratch + (size_t)nqb * 129 + 128, sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  const size_t attn_smem = ((size_t)L * 384 + (size_t)L * D + 2 * NH * L * L) * sizeof(float);
  const int LNB = 296;
  auto ln_bwd = [&](const float* dout, const float* pre, const Drop& dr, const float* nw, float* dpre, float* dbranch,
                    float* gw, float* gb) -> int {
    k_add_ln_bwd<<<LNB, 256, 0, st>>>(dout, pre, dr, nw, R, dpre, dbranch, cx.scratch);
    S3D_LAUNCH_CHECK();
    float* red = cx.scratch + (size_t)LNB * 256;
    k_reduce_parts<<<1, 256, 0, st>>>(cx.scratch, LNB, 256, red, 0);
    S3D_LAUNCH_CHECK();
    S3D_CUDA(cudaMemcpyAsync(gw, red, 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    S3D_CUDA(cudaMemcpyAsync(gb, red + 128, 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return S3D_OK;
  };
  for (int l = 2; l >= 0; --l) {
    const Params::Layer& W = P.l[l];
    const GradParams::Layer& GW = G.l[l];
    const Saved::Layer& A = s.l[l];
    const float* xin = l == 0 ? s.x0 : s.l[l - 1].x2;
    // ---- LayerNorm 2: dX (grad of x2) -> dPre (grad of x1 through the residual), dBr (grad of linear2's output)
    S3D_TRY(ln_bwd(dX, A.pre2, make_drop(p, c->seed, 4 * l + 3), W.n2w, dPre, dBr, GW.n2w, GW.n2b));
    // ---- FFN
    S3D_TRY(linear_bwd_w(cx, dBr, A.hd, (int)R, D, FF, GW.w2, GW.b2));
    const Drop dffn = make_drop(p, c->seed, 4 * l + 2);
    S3D_TRY(linear_bwd_x(cx, dBr, W.w2, (int)R, D, FF, dH, EPI_MASK, A.hd, dffn.scale));  // through dropout and ReLU
    S3D_TRY(linear_bwd_w(cx, dH, A.x1, (int)R, FF, D, GW.w1, GW.b1));
    S3D_TRY(linear_bwd_x(cx, dH, W.w1, (int)R, FF, D, dPre, EPI_ACC));  // dPre = total gradient of x1
    // ---- LayerNorm 1: dPre -> dX (grad of the layer input through the residual), dBr (grad of out-proj's output)
    S3D_TRY(ln_bwd(dPre, A.pre1, make_drop(p, c->seed, 4 * l + 1), W.n1w, dX, dBr, GW.n1w, GW.n1b));
    // ---- attention
    S3D_TRY(linear_bwd_w(cx, dBr, A.o, (int)R, D, D, GW.wo, GW.bo));
    S3D_TRY(linear_bwd_x(cx, dBr, W.wo, (int)R, D, D, dPre, EPI_STORE));  // dPre reused as dO
    k_attn_bwd<<<(unsigned)Q, 128, attn_smem, st>>>(A.qkv, A.p, dPre, dQKV, L, make_drop(p, c->seed, 4 * l + 0));
    S3D_LAUNCH_CHECK();
    S3D_TRY(linear_bwd_w(cx, dQKV, xin, (int)R, 3 * D, D, GW.win, GW.bin));
    S3D_TRY(linear_bwd_x(cx, dQKV, W.win, (int)R, 3 * D, D, dX, EPI_ACC));  // dX = total gradient of the layer input
  }
  // ---- tokens: fc_p from the query rows, fc_s from the slice rows
  k_fcp_grad_part<<<nqb, 128, 0, st>>>(dX, qry, (int)Q, L, QPB, cx.scratch);
  S3D_LAUNCH_CHECK();
  {
    float* red = cx.scratch + (size_t)nqb * 512;
    k_reduce_parts<<<2, 256, 0, st>>>(cx.scratch, nqb, 512, red, 0);
    S3D_LAUNCH_CHECK();
    S3D_CUDA(cudaMemcpyAsync(G.wp, red, 384 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    S3D_CUDA(cudaMemcpyAsync(G.bp, red + 384, 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  k_tokens_split<<<(unsigned)Q, 128, 0, st>>>(dX, K, dtoks);
  S3D_LAUNCH_CHECK();
  S3D_TRY(linear_bwd_w(cx, dtoks, s.agg, (int)(Q * K), D, CAGG, G.ws, G.bs));
  S3D_TRY(linear_bwd_x(cx, dtoks, P.ws, (int)(Q * K), D, CAGG, dAGG, EPI_STORE));
  // ---- grid_sample backward into the (zero-initialised) feature-plane gradients
  k_sample<true><<<(unsigned)(Q * K), 256, 0, st>>>(F, qry, T, c->n_qry, K, dAGG);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

}  // namespace s3d
