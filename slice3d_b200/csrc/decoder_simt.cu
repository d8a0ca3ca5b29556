// Validation-grade fp32 decoder (S3D_PREC_FP32): the reference's per-query path
// (reg_slices/src/models.py:53-84) as a chain of plain CUDA-core kernels with global-memory
// intermediates.  Exact fp32 arithmetic, no tensor cores; it exists to pin the tensor-core
// decoder (decoder_tc.cu) on the device and to serve the <=1e-4 parity tests at any size.
#include "gemm_simt.cuh"

namespace s3d {

namespace {

constexpr int TOK = 128;  // d_model
constexpr int QCHUNK = 8192;

// One block per query, 13 warps: warp 0 builds the query token fc_p(q) (models.py:79), warp
// k+1 the slice token of slice k = fc_s bias + sum over the 5 scales of the bilinear sample of
// the fc_s-projected plane (models.py:71-80; fc_s commutes with the bilinear interpolation).
__global__ void __launch_bounds__(32 * 13) k_tokens(QueryCtx q, long long i0, int n, const float* __restrict__ planes,
                                                    int S, int K, const float* __restrict__ fcp_wt,
                                                    const float* __restrict__ fcp_b, const float* __restrict__ fcs_b,
                                                    float* __restrict__ X) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp > K) return;
  float x, y, z, gu, gv;
  planes += (size_t)load_query(q, i0 + i, x, y, z, gu, gv) * q.plane_stride;
  float* dst = X + ((size_t)i * (K + 1) + warp) * TOK + lane * 4;
  if (warp == 0) {
    float4 r;
    float* rr = reinterpret_cast<float*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = lane * 4 + j;
      rr[j] = fcp_b[c] + x * fcp_wt[c] + y * fcp_wt[TOK + c] + z * fcp_wt[2 * TOK + c];
    }
    *reinterpret_cast<float4*>(dst) = r;
    return;
  }
  const int k = warp - 1;
  float4 acc = *reinterpret_cast<const float4*>(fcs_b + lane * 4);
  size_t off = 0;
  for (int s = 0; s < 5; ++s) {
    const int R = plane_res(S, s);
    const float* P = planes + off + (size_t)k * R * R * TOK + lane * 4;
    Taps t = make_taps(gu, gv, R);
    float4 a = __ldg(reinterpret_cast<const float4*>(P + (size_t)t.o00 * TOK));
    float4 b = __ldg(reinterpret_cast<const float4*>(P + (size_t)t.o01 * TOK));
    float4 c = __ldg(reinterpret_cast<const float4*>(P + (size_t)t.o10 * TOK));
    float4 d = __ldg(reinterpret_cast<const float4*>(P + (size_t)t.o11 * TOK));
    acc.x += a.x * t.w00 + b.x * t.w01 + c.x * t.w10 + d.x * t.w11;
    acc.y += a.y * t.w00 + b.y * t.w01 + c.y * t.w10 + d.y * t.w11;
    acc.z += a.z * t.w00 + b.z * t.w01 + c.z * t.w10 + d.z * t.w11;
    acc.w += a.w * t.w00 + b.w * t.w01 + c.w * t.w10 + d.w * t.w11;
    off += (size_t)K * R * R * TOK;
  }
  *reinterpret_cast<float4*>(dst) = acc;
}

// Ready tokens (Slices3DGTModel, gt.cu): X[i][0] = tok_query[i], X[i][1 + k] = tok_slice[i][k].
__global__ void __launch_bounds__(128) k_assemble_tokens(const float* __restrict__ tq, const float* __restrict__ ts, int K,
                                                         float* __restrict__ X) {
  const int i = blockIdx.x, c = threadIdx.x, L = K + 1;
  X[((size_t)i * L) * TOK + c] = tq[(size_t)i * TOK + c];
  for (int k = 0; k < K; ++k) X[((size_t)i * L + 1 + k) * TOK + c] = ts[((size_t)i * K + k) * TOK + c];
}

// Self-attention of one query's L = K+1 tokens, 4 heads of 32 (nn.MultiheadAttention inside
// nn.TransformerEncoderLayer, models.py:18).  One block (128 threads) per query.
__global__ void __launch_bounds__(128) k_attention(const float* __restrict__ QKV, float* __restrict__ O, int n, int L) {
  extern __shared__ float sm[];
  float* qkv = sm;                // [L][384]
  float* sc = sm + (size_t)L * 384;  // [4][L][L]
  const int i = blockIdx.x;
  if (i >= n) return;
  const float* src = QKV + (size_t)i * L * 384;
  for (int t = threadIdx.x; t < L * 96; t += 128)
    reinterpret_cast<float4*>(qkv)[t] = __ldg(reinterpret_cast<const float4*>(src) + t);
  __syncthreads();
  const float scale = 0.17677669529663687f;  // 1/sqrt(32)
  for (int t = threadIdx.x; t < 4 * L * L; t += 128) {
    int h = t / (L * L), r = t % (L * L), a = r / L, b = r % L;
    const float* qa = qkv + a * 384 + h * 32;
    const float* kb = qkv + b * 384 + 128 + h * 32;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) s = fmaf(qa[c] * scale, kb[c], s);
    sc[t] = s;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 4 * L; t += 128) {
    float* row = sc + (size_t)t * L;
    float mx = row[0];
    for (int b = 1; b < L; ++b) mx = fmaxf(mx, row[b]);
    float sum = 0.f;
    for (int b = 0; b < L; ++b) {
      float e = expf(row[b] - mx);
      row[b] = e;
      sum += e;
    }
    float inv = 1.f / sum;
    for (int b = 0; b < L; ++b) row[b] *= inv;
  }
  __syncthreads();
  const int c = threadIdx.x, h = c >> 5;
  for (int a = 0; a < L; ++a) {
    const float* p = sc + ((size_t)h * L + a) * L;
    float o = 0.f;
    for (int b = 0; b < L; ++b) o = fmaf(p[b], qkv[b * 384 + 256 + c], o);
    O[((size_t)i * L + a) * TOK + c] = o;
  }
}

// X[row] = LayerNorm(X[row] + T[row]) * w + b  (post-norm residual, eps 1e-5).  One warp per row.
__global__ void __launch_bounds__(256) k_add_ln(float* __restrict__ X, const float* __restrict__ T,
                                                const float* __restrict__ w, const float* __restrict__ b,
                                                long long rows) {
  long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float4 a = *reinterpret_cast<const float4*>(X + row * TOK + lane * 4);
  float4 t = *reinterpret_cast<const float4*>(T + row * TOK + lane * 4);
  float v[4] = {a.x + t.x, a.y + t.y, a.z + t.z, a.w + t.w};
  float s = v[0] + v[1] + v[2] + v[3];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / TOK);
  float d2 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j] -= mean;
    d2 += v[j] * v[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
  const float rstd = rsqrtf(d2 * (1.f / TOK) + 1e-5f);
  float4 ww = *reinterpret_cast<const float4*>(w + lane * 4);
  float4 bb = *reinterpret_cast<const float4*>(b + lane * 4);
  float4 r = make_float4(v[0] * rstd * ww.x + bb.x, v[1] * rstd * ww.y + bb.y, v[2] * rstd * ww.z + bb.z,
                         v[3] * rstd * ww.w + bb.w);
  *reinterpret_cast<float4*>(X + row * TOK + lane * 4) = r;
}

// out[i] = out_scale * (fc_out(token 0 of query i))  (models.py:83-84; reconstruct.py:97 negates).
__global__ void __launch_bounds__(256) k_head(const float* __restrict__ X, const float* __restrict__ w,
                                              const float* __restrict__ b, float out_scale, float* __restrict__ out,
                                              int n, int L) {
  int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  float4 a = *reinterpret_cast<const float4*>(X + (size_t)i * L * TOK + lane * 4);
  float4 ww = *reinterpret_cast<const float4*>(w + lane * 4);
  float s = a.x * ww.x + a.y * ww.y + a.z * ww.z + a.w * ww.w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[i] = out_scale * (s + b[0]);
}

int linear(const ConvW& w, const float* a, long long M, float* out, int relu, cudaStream_t st) {
  LoadPlain L{a, (int)M, w.k, w.k};
  EpiAffine E{out, nullptr, w.shift, w.ncols, relu};
  return launch_gemm(L, w.w, w.ncols, w.kpad, E, st);
}

}  // namespace

size_t decoder_simt_workspace_bytes(int64_t n) {
  int64_t c = n < QCHUNK ? n : QCHUNK;
  if (c < 1) c = 1;
  // X, T, O (128 each), QKV (384), H (2048) per token row; 13 rows per query is the reference
  // K = 12; sized for K <= 12.
  return (size_t)c * 13 * (3 * 128 + 384 + 2048) * sizeof(float) + 1024;
}

int decoder_simt(const s3d_model* m, const float* planes, int S, const QueryCtx& q, int64_t n, float out_scale,
                 float* out, float* debug_tokens, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int K = m->K, L = K + 1;
  if (K > 12) {
    set_error("decoder(fp32): n_slices > 12 unsupported");
    return S3D_ERR_UNSUPPORTED;
  }
  if (n <= 0) return S3D_OK;
  if (ws == nullptr || ws_bytes < decoder_simt_workspace_bytes(n)) {
    set_error("decoder(fp32): workspace too small");
    return S3D_ERR_WORKSPACE;
  }
  const DecF32& d = m->dec32;
  float* X = static_cast<float*>(ws);
  float* T = X + (size_t)QCHUNK * 13 * 128;
  float* O = T + (size_t)QCHUNK * 13 * 128;
  float* QKV = O + (size_t)QCHUNK * 13 * 128;
  float* Hh = QKV + (size_t)QCHUNK * 13 * 384;
  if (n < QCHUNK) {  // compact layout for small calls (matches decoder_simt_workspace_bytes)
    T = X + (size_t)n * 13 * 128;
    O = T + (size_t)n * 13 * 128;
    QKV = O + (size_t)n * 13 * 128;
    Hh = QKV + (size_t)n * 13 * 384;
  }
  const size_t attn_smem = ((size_t)L * 384 + 4 * L * L) * sizeof(float);
  for (int64_t i0 = 0; i0 < n; i0 += QCHUNK) {
    const int c = (int)((n - i0) < QCHUNK ? (n - i0) : QCHUNK);
    const long long rows = (long long)c * L;
    if (q.tok_slice)
      k_assemble_tokens<<<c, 128, 0, st>>>(q.tok_query + (size_t)i0 * TOK, q.tok_slice + (size_t)i0 * K * TOK, K, X);
    else
      k_tokens<<<c, 32 * 13, 0, st>>>(q, i0, c, planes, S, K, d.fcp_wt, d.fcp_b, d.fcs_b, X);
    S3D_LAUNCH_CHECK();
    if (debug_tokens)
      S3D_CUDA(cudaMemcpyAsync(debug_tokens + (size_t)i0 * L * 128, X, rows * 128 * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
    for (int l = 0; l < 3; ++l) {
      const DecLayerF32& W = d.L[l];
      S3D_TRY(linear(W.in_proj, X, rows, QKV, 0, st));
      k_attention<<<c, 128, attn_smem, st>>>(QKV, O, c, L);
      S3D_LAUNCH_CHECK();
      S3D_TRY(linear(W.out_proj, O, rows, T, 0, st));
      k_add_ln<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(X, T, W.n1_w, W.n1_b, rows);
      S3D_LAUNCH_CHECK();
      S3D_TRY(linear(W.lin1, X, rows, Hh, 1, st));
      S3D_TRY(linear(W.lin2, Hh, rows, T, 0, st));
      k_add_ln<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(X, T, W.n2_w, W.n2_b, rows);
      S3D_LAUNCH_CHECK();
      if (debug_tokens)
        S3D_CUDA(cudaMemcpyAsync(debug_tokens + ((size_t)(l + 1) * n + i0) * L * 128, X, rows * 128 * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st));
    }
    k_head<<<(c + 7) / 8, 256, 0, st>>>(X, d.fco_w, d.fco_b, out_scale, out + i0, c, L);
    S3D_LAUNCH_CHECK();
  }
  return S3D_OK;
}

}  // namespace s3d
