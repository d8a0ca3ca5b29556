// Slices3DGTModel's per-query path (reg_slices/src/model_gt.py:84-104) in front of the shared transformer decoder:
//
//   slice token (q, k) = ReLU(W2 . ReLU(b1 + sum_s bilinear(plane_s[k], uv(q))) + b2)      fc_local, first Linear hoisted
//   query token (q)    = pts_feat_extractor(q): 3 -> 32 -> 64 -> 128 with ReLUs
//
//   k_gt_tokens   one block per query: warp 0 runs the query MLP, warp k + 1 gathers slice k's sample (5 scales x 4 taps
//                 x 512 B, like the regression model's token build), adds b1, rectifies and writes the row in the
//                 split-fp16 format of the tensor-core GEMM
//   conv_tc       W2 (128 x 128) as a 1x1 "convolution" over the (queries x K) rows, bias + ReLU in the epilogue (tcgen05)
//   decoder       decoder_tc / decoder_simt with ready tokens (QueryCtx::tok_query / tok_slice): the decoders skip their
//                 own gather.
// Token traffic: 13 x 512 B written and read once per query (6.6 KB: ~0.1 % of the decoder's time at HBM rate).
#include "gemm_simt.cuh"

namespace s3d {

namespace {

constexpr int TOK = 128;
constexpr int GT_CHUNK = 32768;  // queries per pass (token buffers: ~420 MB)

__global__ void __launch_bounds__(32 * 13) k_gt_tokens(QueryCtx q, long long i0, int n, const float* __restrict__ planes, int S,
                                                       int K, const float* __restrict__ b1, const float* __restrict__ w0,
                                                       const float* __restrict__ c0, const float* __restrict__ w1,
                                                       const float* __restrict__ c1, const float* __restrict__ w2,
                                                       const float* __restrict__ c2, float* __restrict__ tok_query,
                                                       __half* __restrict__ pre_hi, __half* __restrict__ pre_lo) {
  __shared__ float h0[32], h1[64];
  const int i = blockIdx.x;
  if (i >= n) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp > K) return;
  float x, y, z, gu, gv;
  planes += (size_t)load_query(q, i0 + i, x, y, z, gu, gv) * q.plane_stride;
  if (warp == 0) {  // pts_feat_extractor (model_gt.py:24-31): PyTorch (out, in) weights
    h0[lane] = fmaxf(c0[lane] + w0[3 * lane] * x + w0[3 * lane + 1] * y + w0[3 * lane + 2] * z, 0.f);
    __syncwarp();
#pragma unroll
    for (int o = lane; o < 64; o += 32) {
      float s = c1[o];
      for (int j = 0; j < 32; ++j) s = fmaf(w1[o * 32 + j], h0[j], s);
      h1[o] = fmaxf(s, 0.f);
    }
    __syncwarp();
    float4 r;
    float* rr = reinterpret_cast<float*>(&r);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int o = lane * 4 + jj;
      float s = c2[o];
      for (int j = 0; j < 64; ++j) s = fmaf(w2[o * 64 + j], h1[j], s);
      rr[jj] = fmaxf(s, 0.f);
    }
    *reinterpret_cast<float4*>(tok_query + (size_t)i * TOK + lane * 4) = r;
    return;
  }
  const int k = warp - 1;
  float4 acc = *reinterpret_cast<const float4*>(b1 + lane * 4);
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
  const float r[4] = {fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f)};
  store_split4(pre_hi, pre_lo, ((size_t)i * K + k) * TOK + lane * 4, r);
}

size_t padded_rows(int64_t c, int K) { return ((size_t)c * K + 127) / 128 * 128; }

}  // namespace

size_t gt_decoder_workspace_bytes(int64_t n, int precision) {
  const int64_t c = n < GT_CHUNK ? (n < 1 ? 1 : n) : GT_CHUNK;
  const size_t rows = padded_rows(c, 12);  // K <= 12
  const size_t tok = rows * TOK * (2 * sizeof(__half) + sizeof(float)) + (size_t)c * TOK * sizeof(float) + 1024;
  const size_t dec = precision == S3D_PREC_FP32 ? decoder_simt_workspace_bytes(c) : decoder_tc_workspace_bytes(c);
  return tok + ((dec + 255) / 256) * 256;
}

int gt_decoder_fwd(const s3d_model* m, const void* planes, int S, QueryCtx q, int64_t n, float out_scale, float* out,
                   int precision, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!m || m->kind != 1) {
    set_error("gt_decoder: the handle does not hold a Slices3DGTModel");
    return S3D_ERR_BAD_ARG;
  }
  if (n <= 0) return S3D_OK;
  const int K = m->K;
  const int64_t cmax = n < GT_CHUNK ? n : GT_CHUNK;
  const size_t rows = padded_rows(cmax, K);
  char* w = static_cast<char*>(ws);
  __half* pre_hi = reinterpret_cast<__half*>(w);            w += rows * TOK * sizeof(__half);
  __half* pre_lo = reinterpret_cast<__half*>(w);            w += rows * TOK * sizeof(__half);
  float* tok_slice = reinterpret_cast<float*>(w);           w += rows * TOK * sizeof(float);
  float* tok_query = reinterpret_cast<float*>(w);           w += (size_t)cmax * TOK * sizeof(float);
  w = static_cast<char*>(ws) + ((w - static_cast<char*>(ws)) + 255) / 256 * 256;
  const size_t dec_bytes = ws_bytes - (size_t)(w - static_cast<char*>(ws));
  const DecF32& d = m->dec32;
  for (int64_t i0 = 0; i0 < n; i0 += GT_CHUNK) {
    const int c = (int)((n - i0) < GT_CHUNK ? (n - i0) : GT_CHUNK);
    const size_t r = padded_rows(c, K);
    if (r > (size_t)c * K)  // rows of the last 128-row tile beyond the chunk: defined (zero) GEMM inputs
      S3D_CUDA(cudaMemsetAsync(pre_hi + (size_t)c * K * TOK, 0, (r - (size_t)c * K) * TOK * sizeof(__half), st));
    if (r > (size_t)c * K)
      S3D_CUDA(cudaMemsetAsync(pre_lo + (size_t)c * K * TOK, 0, (r - (size_t)c * K) * TOK * sizeof(__half), st));
    k_gt_tokens<<<c, 32 * 13, 0, st>>>(q, i0, c, static_cast<const float*>(planes), S, K, d.fcs_b, m->pts_w[0], m->pts_b[0],
                                      m->pts_w[1], m->pts_b[1], m->pts_w[2], m->pts_b[2], tok_query, pre_hi, pre_lo);
    S3D_LAUNCH_CHECK();
    // fc_local's second Linear + ReLU over the (c * K) rows: a 1x1 convolution on an image of r / 128 rows of 128 pixels
    S3D_TRY(conv_tc(m->tfcl2, pre_hi, pre_lo, 1, (int)(r / 128), 128, nullptr, 1, 1, tok_slice, TOK, nullptr, nullptr, 0, st));
    QueryCtx qc{};
    qc.tok_query = tok_query;
    qc.tok_slice = tok_slice;
    qc.T = q.T;  // (unused by the decoders in token mode; kept non-null for their argument checks)
    if (precision == S3D_PREC_FP32)
      S3D_TRY(decoder_simt(m, static_cast<const float*>(planes), S, qc, c, out_scale, out + i0, nullptr, w, dec_bytes, st));
    else
      S3D_TRY(decoder_tc(m, static_cast<const float*>(planes), S, qc, c, out_scale, out + i0, precision, w, dec_bytes, st));
  }
  return S3D_OK;
}

}  // namespace s3d
