// Plane encoder: the reference's slice generator (reg_slices/src/unet_custom.py:40-69,
// reg_slices/src/unet_parts.py) in eval mode, NHWC fp32, every convolution an implicit GEMM
// (gemm_simt.cuh), BatchNorm folded into the GEMM epilogue, plus the hoisted fc_s projection
// (reg_slices/src/models.py:80) that turns the five feature planes into five 128-channel
// channels-last planes the decoder samples directly.
//
// Work the reference repeats 12x (the 1x1 skip adapters trans_up1..4 and the trans_c
// contribution of x5 run on the slice-tiled batch, unet_custom.py:57-66) is done once per
// input view here; the results are identical because those operands do not depend on the slice.
#include <cstdlib>
#include <cstring>

#include "gemm_simt.cuh"

namespace s3d {

namespace {

__global__ void k_nchw3_to_nhwc4(const float* __restrict__ in, float* __restrict__ out, int B, int HW) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * HW) return;
  int b = (int)(i / HW), p = (int)(i % HW);
  const float* s = in + (size_t)b * 3 * HW + p;
  *reinterpret_cast<float4*>(out + i * 4) = make_float4(s[0], s[HW], s[2 * HW], 0.f);
}

// y = maxpool2x2(relu(x*scale + shift)) on NHWC; C % 4 == 0.
__global__ void k_bn_relu_pool(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ scale,
                               const float* __restrict__ shift, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * Ho * Wo * C4;
  if (i >= total) return;
  int c = (int)(i % C4) * 4;
  long long t = i / C4;
  int xo = (int)(t % Wo);
  t /= Wo;
  int yo = (int)(t % Ho);
  int b = (int)(t / Ho);
  float4 sc = *reinterpret_cast<const float4*>(scale + c);
  float4 sh = *reinterpret_cast<const float4*>(shift + c);
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f);  // relu output >= 0
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      float4 v = *reinterpret_cast<const float4*>(in + (((size_t)b * H + 2 * yo + dy) * W + 2 * xo + dx) * C + c);
      m.x = fmaxf(m.x, fmaf(v.x, sc.x, sh.x));
      m.y = fmaxf(m.y, fmaf(v.y, sc.y, sh.y));
      m.z = fmaxf(m.z, fmaf(v.z, sc.z, sh.z));
      m.w = fmaxf(m.w, fmaf(v.w, sc.w, sh.w));
    }
  *reinterpret_cast<float4*>(out + (((size_t)b * Ho + yo) * Wo + xo) * C + c) = m;
}

// Same, written in the split-fp16 format the tensor-core convolutions read.
__global__ void k_bn_relu_pool_split(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo,
                                     const float* __restrict__ scale, const float* __restrict__ shift, int B, int H, int W,
                                     int C) {
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * Ho * Wo * C4;
  if (i >= total) return;
  int c = (int)(i % C4) * 4;
  long long t = i / C4;
  int xo = (int)(t % Wo);
  t /= Wo;
  int yo = (int)(t % Ho);
  int b = (int)(t / Ho);
  float4 sc = *reinterpret_cast<const float4*>(scale + c);
  float4 sh = *reinterpret_cast<const float4*>(shift + c);
  float m[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      float4 v = *reinterpret_cast<const float4*>(in + (((size_t)b * H + 2 * yo + dy) * W + 2 * xo + dx) * C + c);
      m[0] = fmaxf(m[0], fmaf(v.x, sc.x, sh.x));
      m[1] = fmaxf(m[1], fmaf(v.y, sc.y, sh.y));
      m[2] = fmaxf(m[2], fmaf(v.z, sc.z, sh.z));
      m[3] = fmaxf(m[3], fmaf(v.w, sc.w, sh.w));
    }
  store_split4(hi, lo, (((size_t)b * Ho + yo) * Wo + xo) * C + c, m);
}

// latent[b,k,p,:] = base[b,p,:] + e[k,:]   (trans_c split into its x5 part and its slice-embedding part)
__global__ void k_add_slice_bias(const float* __restrict__ base, const float* __restrict__ e, float* __restrict__ out,
                                 int B, int K, int HW, int C, __half* __restrict__ hi = nullptr,
                                 __half* __restrict__ lo = nullptr) {
  const int C4 = C / 4;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * K * HW * C4;
  if (i >= total) return;
  int c = (int)(i % C4) * 4;
  long long t = i / C4;
  int p = (int)(t % HW);
  t /= HW;
  int k = (int)(t % K);
  int b = (int)(t / K);
  float4 a = *reinterpret_cast<const float4*>(base + ((size_t)b * HW + p) * C + c);
  float4 v = *reinterpret_cast<const float4*>(e + (size_t)k * C + c);
  const float r[4] = {a.x + v.x, a.y + v.y, a.z + v.z, a.w + v.w};
  *reinterpret_cast<float4*>(out + i * 4) = make_float4(r[0], r[1], r[2], r[3]);
  if (hi) store_split4(hi, lo, (size_t)i * 4, r);
}

// slices_rec = tanh(conv1x1 32->3) written NCHW (unet_parts.py:78-84).
__global__ void k_outc_tanh(const float* __restrict__ f, const float* __restrict__ w, const float* __restrict__ b,
                            float* __restrict__ out, long long NI, int HW) {
  __shared__ float sw[3 * 32 + 3];
  if (threadIdx.x < 99) sw[threadIdx.x] = threadIdx.x < 96 ? w[threadIdx.x] : b[threadIdx.x - 96];
  __syncthreads();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NI * HW) return;
  long long img = i / HW;
  int p = (int)(i % HW);
  const float4* src = reinterpret_cast<const float4*>(f + i * 32);
  float a0 = sw[96], a1 = sw[97], a2 = sw[98];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 v = __ldg(src + q);
    const float* w0 = sw + q * 4;
    a0 += v.x * w0[0] + v.y * w0[1] + v.z * w0[2] + v.w * w0[3];
    a1 += v.x * w0[32] + v.y * w0[33] + v.z * w0[34] + v.w * w0[35];
    a2 += v.x * w0[64] + v.y * w0[65] + v.z * w0[66] + v.w * w0[67];
  }
  float* o = out + (size_t)img * 3 * HW + p;
  o[0] = tanhf(a0);
  o[HW] = tanhf(a1);
  o[2 * HW] = tanhf(a2);
}

// NHWC -> NCHW export of a feature plane (parity tests / callers that want the raw planes).
__global__ void k_nhwc_to_nchw(const float* __restrict__ in, float* __restrict__ out, int HW, int C) {
  __shared__ float tile[32][33];
  int img = blockIdx.z;
  int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* src = in + (size_t)img * HW * C;
  float* dst = out + (size_t)img * HW * C;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int p = p0 + j, c = c0 + threadIdx.x;
    if (p < HW && c < C) tile[j][threadIdx.x] = src[(size_t)p * C + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, p = p0 + threadIdx.x;
    if (p < HW && c < C) dst[(size_t)c * HW + p] = tile[threadIdx.x][j];
  }
}

struct Bump {
  char* base;
  size_t off = 0, cap;
  float* take(size_t floats) {
    size_t o = off;
    off += ((floats * sizeof(float) + 255) / 256) * 256;
    return base ? reinterpret_cast<float*>(base + o) : nullptr;
  }
};

struct EncBufs {
  float *x0, *ta, *tb, *x[6], *base5, *skip[5], *pskip[5], *feat[5], *u, *d;
  float *xs[6], *fs[5];  // split copies of the taps / feature planes (inputs of tensor-core GEMMs)
  SplitK sk;             // split-K scratch of the small-image convolutions (conv_tc.cu)
};

// A split-fp16 activation tensor carved from `elems` floats of workspace: hi then lo, `elems` bf16 each.
struct Split {
  __half *hi, *lo;
};
inline Split split_of(float* base, size_t elems) {
  __half* h = reinterpret_cast<__half*>(base);
  return Split{h, h + elems};
}

// The single place that lays out the encoder workspace; called with base == nullptr to size it.
void carve(Bump& bp, EncBufs& e, int B, int K, int S) {
  const size_t S2 = (size_t)S * S;
  e.x0 = bp.take(B * S2 * 4);
  e.ta = bp.take(B * S2 * 64);
  e.tb = bp.take(B * S2 * 64);
  const int xc[6] = {0, 64, 128, 256, 512, 512};
  for (int i = 1; i <= 5; ++i) e.x[i] = bp.take(B * (S2 >> (2 * (i - 1))) * xc[i]);
  const int R0 = S / 16;
  e.base5 = bp.take((size_t)B * R0 * R0 * 512);
  for (int i = 1; i <= 5; ++i) e.xs[i] = bp.take(B * (S2 >> (2 * (i - 1))) * xc[i]);
  for (int s = 0; s < 5; ++s)
    e.fs[s] = bp.take((size_t)B * K * plane_res(S, s) * plane_res(S, s) * (kPlaneC[s] < 64 ? 64 : kPlaneC[s]));
  for (int n = 1; n <= 4; ++n)
    e.skip[n] = bp.take((size_t)B * plane_res(S, n) * plane_res(S, n) * (kPlaneC[n] < 64 ? 64 : kPlaneC[n]));
  for (int n = 1; n <= 4; ++n) e.pskip[n] = bp.take((size_t)B * plane_res(S, n) * plane_res(S, n) * kPlaneC[n]);
  for (int s = 0; s < 5; ++s) e.feat[s] = bp.take((size_t)B * K * plane_res(S, s) * plane_res(S, s) * kPlaneC[s]);
  e.u = bp.take((size_t)B * K * S2 * 64);  // (split tensors with the channel pitch padded to 64 at the last stage)
  e.d = bp.take((size_t)B * K * S2 * 64);
  e.sk.part = bp.take(SPLITK_PART_BYTES / sizeof(float));
  e.sk.part_bytes = SPLITK_PART_BYTES;
  e.sk.cnt = reinterpret_cast<unsigned*>(bp.take(SPLITK_COUNTERS));
  e.sk.n_cnt = SPLITK_COUNTERS;
}

int conv(const ConvW& w, const float* src0, int c0, int bcast0, const float* src1, int c1, int NI, int H, int W,
         float* out, int relu, cudaStream_t st) {
  LoadConv L{src0, src1, NI * H * W, w.k, H, W, c0, c1, bcast0, w.ks};
  EpiAffine E{out, w.scale, w.shift, w.ncols, relu};
  return launch_gemm(L, w.w, w.ncols, w.kpad, E, st);
}

int dense(const ConvW& w, const float* a, long long M, float* out, int relu, cudaStream_t st) {
  LoadPlain L{a, (int)M, w.k, w.k};
  EpiAffine E{out, w.scale, w.shift, w.ncols, relu};
  return launch_gemm(L, w.w, w.ncols, w.kpad, E, st);
}

inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }


}  // namespace

// VGG16-BN trunk on B images on the tensor cores (unet_custom.py:42-48 / vgg16bn_feats.py:44-52): taps x[1..5] = the
// pre-BatchNorm outputs of the last convolution of each block, fp32 NHWC, and their split-fp16 copies xs[1..5].
int trunk_tc(const s3d_model* m, const float* img, int B, int S, float* x0, float* ta, float* tb, float* const* x,
             float* const* xs, cudaStream_t st, const SplitK* sk) {
  const size_t S2 = (size_t)S * S;
  k_nchw3_to_nhwc4<<<blocks_for((long long)B * S * S, 256), 256, 0, st>>>(img, x0, B, S * S);
  S3D_LAUNCH_CHECK();
  int H = S;
  const int xc[6] = {0, 64, 128, 256, 512, 512};
  auto xs_of = [&](int i) { return split_of(xs[i], B * (S2 >> (2 * (i - 1))) * xc[i]); };
  Split sa = split_of(ta, B * S2 * 64);
  {  // down1: conv0 on the fp32 path (K = 27), written split; conv1 -> tap x1 (pre-BN)
    const ConvW& w = m->vgg[0];
    LoadConv L{x0, nullptr, B * H * H, w.k, H, H, 4, 0, 1, w.ks};
    EpiAffineSplit E{sa.hi, sa.lo, w.scale, w.shift, w.ncols, 1};
    S3D_TRY(launch_gemm(L, w.w, w.ncols, w.kpad, E, st));
    Split x1 = xs_of(1);
    S3D_TRY(conv_tc(m->tvgg[1], sa.hi, sa.lo, B, H, H, nullptr, 1, 0, x[1], 64, x1.hi, x1.lo, 64, st));
  }
  const int first_conv[4] = {2, 4, 7, 10};
  const int n_conv[4] = {2, 3, 3, 3};
  const int cprev[4] = {64, 128, 256, 512};
  for (int b = 0; b < 4; ++b) {
    long long tot = (long long)B * (H / 2) * (H / 2) * (cprev[b] / 4);
    H /= 2;
    Split cur = split_of(ta, (size_t)B * H * H * cprev[b]);
    k_bn_relu_pool_split<<<blocks_for(tot, 256), 256, 0, st>>>(x[b + 1], cur.hi, cur.lo, m->bn_scale[b], m->bn_shift[b], B,
                                                               2 * H, 2 * H, cprev[b]);
    S3D_LAUNCH_CHECK();
    float* nxt_base = tb;
    float* cur_base = ta;
    for (int j = 0; j < n_conv[b]; ++j) {
      const ConvTC& w = m->tvgg[first_conv[b] + j];
      const bool last = (j == n_conv[b] - 1);
      if (last) {
        Split xo = xs_of(b + 2);
        S3D_TRY(conv_tc(w, cur.hi, cur.lo, B, H, H, nullptr, 1, 0, x[b + 2], w.cout, xo.hi, xo.lo, w.cout, st, 0, sk));
      } else {
        Split nxt = split_of(nxt_base, (size_t)B * H * H * w.cout);
        S3D_TRY(conv_tc(w, cur.hi, cur.lo, B, H, H, nullptr, 1, 1, nullptr, 0, nxt.hi, nxt.lo, w.cout, st, 0, sk));
        cur = nxt;
        float* t = cur_base;
        cur_base = nxt_base;
        nxt_base = t;
      }
    }
  }
  return S3D_OK;
}

namespace {

// The encoder on the tensor cores (conv_tc.cu): every convolution except the first (3 input channels), the
// transposed convolutions, the 1x1 adapters and the hoisted fc_s projection.  Activations that feed a tensor-core
// GEMM are written in the split-fp16 format by their producer; taps and feature planes are also kept in fp32 for
// their fp32 consumers (BatchNorm+pool, tanh head, export).  DoubleConv's first convolution over cat([skip, up])
// (unet_parts.py:73) is split by input-channel half: the skip half does not depend on the slice, so it is evaluated
// once per view (B images) and added in the epilogue of the per-slice half (B*K images) -- half of that
// convolution's work, 12x less of it, same result.
int trunk_and_up_tc(const s3d_model* m, const float* img, int B, int S, EncBufs& e, cudaStream_t st) {
  const int K = m->K;
  const size_t S2 = (size_t)S * S;
  const int xc[6] = {0, 64, 128, 256, 512, 512};
  auto xs_of = [&](int i) { return split_of(e.xs[i], B * (S2 >> (2 * (i - 1))) * xc[i]); };
  S3D_TRY(trunk_tc(m, img, B, S, e.x0, e.ta, e.tb, e.x, e.xs, st, &e.sk));
  // latent = trans_c(cat[x5 tiled, slice embedding]) (unet_custom.py:52-57) -> feat[0]
  const int R0 = S / 16;
  auto fs_of = [&](int s) {
    const int R = plane_res(S, s), CP = kPlaneC[s] < 64 ? 64 : kPlaneC[s];
    return split_of(e.fs[s], (size_t)B * K * R * R * CP);
  };
  {
    Split x5 = xs_of(5);
    S3D_TRY(conv_tc(m->ttrans_c, x5.hi, x5.lo, B, R0, R0, nullptr, 1, 0, e.base5, 512, nullptr, nullptr, 0, st, 0, &e.sk));
    long long tot = (long long)B * K * R0 * R0 * (512 / 4);
    Split f0 = fs_of(0);
    k_add_slice_bias<<<blocks_for(tot, 256), 256, 0, st>>>(e.base5, m->trans_c_e, e.feat[0], B, K, R0 * R0, 512, f0.hi, f0.lo);
    S3D_LAUNCH_CHECK();
  }
  for (int n = 1; n <= 4; ++n) {
    const int Rp = plane_res(S, n - 1), R = plane_res(S, n), C = kPlaneC[n];
    const int CP = C < 64 ? 64 : C;  // channel pitch of the split tensors
    const size_t elems = (size_t)B * K * R * R * CP;
    Split su = split_of(e.u, elems), sd = split_of(e.d, elems);
    Split sk = split_of(e.skip[n], (size_t)B * R * R * CP);
    Split xin = xs_of(5 - n), fin = fs_of(n - 1), fout = fs_of(n);
    // skip adapter on the un-tiled tap, then the skip half of DoubleConv's first convolution (raw sums)
    S3D_TRY(conv_tc(m->ttrans_up[n - 1], xin.hi, xin.lo, B, R, R, nullptr, 1, 0, nullptr, 0, sk.hi, sk.lo, CP, st, 0, &e.sk));
    S3D_TRY(conv_tc(m->tdc1s[n - 1], sk.hi, sk.lo, B, R, R, nullptr, 1, 0, e.pskip[n], C, nullptr, nullptr, 0, st, 0, &e.sk));
    // ConvTranspose2d 2x2 s2: 1x1 GEMM with N = 4*C + pixel shuffle
    if (CP != C) S3D_CUDA(cudaMemsetAsync(e.u, 0, elems * 4, st));  // zero channel padding of `up`
    S3D_TRY(conv_tc(m->tup_t[n - 1], fin.hi, fin.lo, B * K, Rp, Rp, nullptr, 1, 0, nullptr, 0, su.hi, su.lo, CP, st, C));
    // DoubleConv: per-slice half of the first convolution (+ skip half, BN, ReLU), then the second convolution
    S3D_TRY(conv_tc(m->tdc1[n - 1], su.hi, su.lo, B * K, R, R, e.pskip[n], K, 1, nullptr, 0, sd.hi, sd.lo, CP, st, 0, &e.sk));
    S3D_TRY(conv_tc(m->tdc2[n - 1], sd.hi, sd.lo, B * K, R, R, nullptr, 1, 1, e.feat[n], C, fout.hi, fout.lo, CP, st, 0, &e.sk));
  }
  return S3D_OK;
}

// Hoisted fc_s (models.py:80): plane_s = fc_s[:, scale-s columns] applied to feature plane s, a 1x1 GEMM per scale.
int project_planes_tc(const s3d_model* m, int B, int S, EncBufs& e, float* pl, cudaStream_t st) {
  const int K = m->K;
  const size_t per_img = s3d_planes_bytes(1, K, S) / sizeof(float);
  for (int b = 0; b < B; ++b)
    for (int s = 0; s < 5; ++s) {
      const int R = plane_res(S, s), CP = kPlaneC[s] < 64 ? 64 : kPlaneC[s];
      Split f = split_of(e.fs[s], (size_t)B * K * R * R * CP);
      const size_t o = (size_t)b * K * R * R * CP;
      S3D_TRY(conv_tc(m->tfcs[s], f.hi + o, f.lo + o, K, R, R, nullptr, 1, 0, pl + b * per_img + plane_offset_floats(K, S, s), 128,
                      nullptr, nullptr, 0, st));
    }
  return S3D_OK;
}

}  // namespace

// Weight images of the tensor-core convolutions + the skip half of each DoubleConv (see trunk_and_up_tc).
int enctc_pack(s3d_model* m, cudaStream_t st) {
  auto pack_pvgg = [&]() -> int {
    for (int i = 1; i < 14; ++i) {
      S3D_TRY(convtc_pack(m, m->pvgg[i], m->pvgg[i].cin, 0, m->pvgg[i].cin, m->tpvgg[i], st));
      S3D_TRY(convtc_pack(m, m->pvgg_d[i], m->pvgg_d[i].cin, 0, m->pvgg_d[i].cin, m->tpvgg_d[i], st));
    }
    return S3D_OK;
  };
  if (m->kind == 2) return pack_pvgg();
  for (int i = 1; i < 13; ++i) S3D_TRY(convtc_pack(m, m->vgg[i], m->vgg[i].cin, 0, m->vgg[i].cin, m->tvgg[i], st));
  if (m->kind == 1) {  // Slices3DGTModel: trunk + hoisted fc_local.0 per tap + fc_local.2
    const int sc[5] = {512, 512, 256, 128, 64};  // channels of the tap behind plane scale s (tap 4 - s)
    for (int s = 0; s < 5; ++s) S3D_TRY(convtc_pack(m, m->fcs[s], sc[s], 0, sc[s], m->tfcs[s], st));
    S3D_TRY(convtc_pack(m, m->fcl2, 128, 0, 128, m->tfcl2, st));
    return S3D_OK;
  }
  for (int n = 0; n < 4; ++n) {
    const int C = kPlaneC[n + 1];
    const ConvW& d1 = m->dc1[n];
    S3D_TRY(convtc_pack(m, d1, 2 * C, C, C, m->tdc1[n], st));  // input channels [C, 2C) = the up-sampled half
    S3D_TRY(convtc_pack(m, m->dc2[n], C, 0, C, m->tdc2[n], st));
    // skip half: fp32 [9*C][C], raw sums (scale/shift are applied by the per-slice half's epilogue)
    std::vector<float> h((size_t)d1.kpad * C), t((size_t)9 * C * C);
    S3D_CUDA(cudaMemcpyAsync(h.data(), d1.w, h.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    S3D_CUDA(cudaStreamSynchronize(st));
    for (int tap = 0; tap < 9; ++tap)
      for (int ci = 0; ci < C; ++ci)
        std::memcpy(&t[((size_t)tap * C + ci) * C], &h[((size_t)tap * 2 * C + ci) * C], C * sizeof(float));
    void* d = nullptr;
    S3D_CUDA(cudaMalloc(&d, t.size() * sizeof(float)));
    m->allocs.push_back(d);
    S3D_CUDA(cudaMemcpyAsync(d, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    S3D_CUDA(cudaStreamSynchronize(st));
    ConvW& s = m->dc1s[n];
    s.w = static_cast<float*>(d);
    s.cin = C; s.ncols = C; s.ks = 3; s.k = 9 * C; s.kpad = 9 * C;
    S3D_TRY(convtc_pack(m, s, C, 0, C, m->tdc1s[n], st));
    S3D_TRY(convtc_pack(m, m->trans_up[n], 2 * C, 0, 2 * C, m->ttrans_up[n], st));
    S3D_TRY(convtc_pack(m, m->up_t[n], m->up_t[n].cin, 0, m->up_t[n].cin, m->tup_t[n], st));
  }
  if (m->has_pvgg) S3D_TRY(pack_pvgg());
  S3D_TRY(convtc_pack(m, m->trans_c, 512, 0, 512, m->ttrans_c, st));
  for (int s = 0; s < 5; ++s) S3D_TRY(convtc_pack(m, m->fcs[s], kPlaneC[s], 0, kPlaneC[s], m->tfcs[s], st));
  return S3D_OK;
}

namespace {
}  // namespace

size_t encoder_workspace_bytes(int B, int K, int S) {
  Bump bp{nullptr, 0, 0};
  EncBufs e;
  carve(bp, e, B, K, S);
  return bp.off;
}

int encoder_fwd(const s3d_model* m, const float* img, int B, int S, void* planes, float* const* feats_nchw,
                float* slices_rec, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int K = m->K;
  if (m->kind != 0) {
    set_error("encoder: this handle holds a Slices3DGTModel (use s3d_gt_encoder_fwd)");
    return S3D_ERR_BAD_ARG;
  }
  if (B <= 0 || S < 32 || (S % 16) != 0) {
    set_error("encoder: S must be a multiple of 16 (>= 32) and B positive");
    return S3D_ERR_BAD_ARG;
  }
  if ((long long)B * K * S * S >= (1ll << 31) / 4) {
    set_error("encoder: B*K*S*S too large for 32-bit row indexing");
    return S3D_ERR_UNSUPPORTED;
  }
  if (ws == nullptr || ws_bytes < encoder_workspace_bytes(B, K, S)) {
    set_error("encoder: workspace too small");
    return S3D_ERR_WORKSPACE;
  }
  Bump bp{static_cast<char*>(ws), 0, ws_bytes};
  EncBufs e;
  carve(bp, e, B, K, S);

  if (!m->enc_simt) {
    S3D_CUDA(cudaMemsetAsync(e.sk.cnt, 0, SPLITK_COUNTERS * sizeof(unsigned), st));
    S3D_TRY(trunk_and_up_tc(m, img, B, S, e, st));
  } else {
    // ---- VGG16-BN trunk on the input view (unet_custom.py:42-48); taps x1..x5 are the
    //      pre-BN outputs of the last convolution of each block (unet_custom.py:15-19).
    k_nchw3_to_nhwc4<<<blocks_for((long long)B * S * S, 256), 256, 0, st>>>(img, e.x0, B, S * S);
    S3D_LAUNCH_CHECK();
    int H = S;
    // down1
    S3D_TRY(conv(m->vgg[0], e.x0, 4, 1, nullptr, 0, B, H, H, e.ta, 1, st));
    S3D_TRY(conv(m->vgg[1], e.ta, 64, 1, nullptr, 0, B, H, H, e.x[1], 0, st));
    // down2..down5: BN+ReLU+pool on the previous tap, then 2 or 3 convolutions
    const int first_conv[4] = {2, 4, 7, 10};
    const int n_conv[4] = {2, 3, 3, 3};
    const int cprev[4] = {64, 128, 256, 512};
    for (int b = 0; b < 4; ++b) {
      long long tot = (long long)B * (H / 2) * (H / 2) * (cprev[b] / 4);
      k_bn_relu_pool<<<blocks_for(tot, 256), 256, 0, st>>>(e.x[b + 1], e.ta, m->bn_scale[b], m->bn_shift[b], B, H, H,
                                                           cprev[b]);
      S3D_LAUNCH_CHECK();
      H /= 2;
      float* cur = e.ta;
      float* nxt = e.tb;
      int cin = cprev[b];
      for (int j = 0; j < n_conv[b]; ++j) {
        const ConvW& w = m->vgg[first_conv[b] + j];
        bool last = (j == n_conv[b] - 1);
        float* dst = last ? e.x[b + 2] : nxt;
        S3D_TRY(conv(w, cur, cin, 1, nullptr, 0, B, H, H, dst, last ? 0 : 1, st));
        cin = w.ncols;
        if (!last) {
          float* t = cur;
          cur = nxt;
          nxt = t;
        }
      }
    }
    // ---- latent = trans_c(cat[x5 tiled, slice embedding]) (unet_custom.py:52-57) -> feat[0]
    const int R0 = S / 16;
    S3D_TRY(dense(m->trans_c, e.x[5], (long long)B * R0 * R0, e.base5, 0, st));
    {
      long long tot = (long long)B * K * R0 * R0 * (512 / 4);
      k_add_slice_bias<<<blocks_for(tot, 256), 256, 0, st>>>(e.base5, m->trans_c_e, e.feat[0], B, K, R0 * R0, 512);
      S3D_LAUNCH_CHECK();
    }
    // ---- four Up stages (unet_custom.py:60-66, unet_parts.py:57-75)
    for (int n = 1; n <= 4; ++n) {
      const int Rp = plane_res(S, n - 1), R = plane_res(S, n), C = kPlaneC[n];
      // skip adapter on the un-tiled tap x_{5-n}
      S3D_TRY(dense(m->trans_up[n - 1], e.x[5 - n], (long long)B * R * R, e.skip[n], 0, st));
      // ConvTranspose2d 2x2 s2: GEMM with N = 4*C + pixel shuffle
      {
        const ConvW& w = m->up_t[n - 1];
        LoadPlain L{e.feat[n - 1], B * K * Rp * Rp, w.k, w.k};
        EpiShuffle2x E{e.u, w.shift, Rp, Rp, C};
        S3D_TRY(launch_gemm(L, w.w, w.ncols, w.kpad, E, st));
      }
      // DoubleConv on cat([skip, up]) (skip first: unet_parts.py:73)
      S3D_TRY(conv(m->dc1[n - 1], e.skip[n], C, K, e.u, C, B * K, R, R, e.d, 1, st));
      S3D_TRY(conv(m->dc2[n - 1], e.d, C, 1, nullptr, 0, B * K, R, R, e.feat[n], 1, st));
    }
  }
  // ---- outputs
  if (slices_rec) {
    long long tot = (long long)B * K * S * S;
    k_outc_tanh<<<blocks_for(tot, 256), 256, 0, st>>>(e.feat[4], m->outc_w, m->outc_b, slices_rec, (long long)B * K,
                                                      S * S);
    S3D_LAUNCH_CHECK();
  }
  if (planes && !m->enc_simt) {
    S3D_TRY(project_planes_tc(m, B, S, e, static_cast<float*>(planes), st));
  } else if (planes) {
    float* pl = static_cast<float*>(planes);
    const size_t per_img = s3d_planes_bytes(1, K, S) / sizeof(float);
    for (int b = 0; b < B; ++b)
      for (int s = 0; s < 5; ++s) {
        const int R = plane_res(S, s);
        const float* src = e.feat[s] + (size_t)b * K * R * R * kPlaneC[s];
        S3D_TRY(dense(m->fcs[s], src, (long long)K * R * R, pl + b * per_img + plane_offset_floats(K, S, s), 0, st));
      }
  }
  if (feats_nchw) {
    for (int s = 0; s < 5; ++s) {
      if (!feats_nchw[s]) continue;
      const int R = plane_res(S, s), C = kPlaneC[s];
      dim3 grid((R * R + 31) / 32, (C + 31) / 32, B * K), blk(32, 8);
      k_nhwc_to_nchw<<<grid, blk, 0, st>>>(e.feat[s], feats_nchw[s], R * R, C);
      S3D_LAUNCH_CHECK();
    }
  }
  return S3D_OK;
}

// ---- Slices3DGTModel (reg_slices/src/model_gt.py:81-96, src/vgg16bn_feats.py:44-58) --------------------------------
// The trunk runs on the B*K GIVEN slice images; tap i (conv1_2 .. conv5_3: 64 @ S ... 512 @ S/16) is projected by its
// column block of fc_local's first Linear (hoisted: bilinear sampling is linear) into plane scale 4 - i of the same
// (K, R_s, R_s, 128) fp32 channels-last blob the regression model's decoder samples.
size_t gt_encoder_workspace_bytes(int N, int S) { return encoder_workspace_bytes(N, 1, S); }

int gt_encoder_fwd(const s3d_model* m, const float* img_slices, int B, int S, void* planes, float* const* taps_nchw, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  if (!m || m->kind != 1) {
    set_error("gt_encoder: the handle does not hold a Slices3DGTModel");
    return S3D_ERR_BAD_ARG;
  }
  const int K = m->K, N = B * K;
  if (!img_slices || !planes || B <= 0 || S < 32 || (S % 16) != 0) {
    set_error("gt_encoder: bad argument (S must be a multiple of 16, >= 32)");
    return S3D_ERR_BAD_ARG;
  }
  if (ws == nullptr || ws_bytes < gt_encoder_workspace_bytes(N, S)) {
    set_error("gt_encoder: workspace too small");
    return S3D_ERR_WORKSPACE;
  }
  Bump bp{static_cast<char*>(ws), 0, ws_bytes};
  EncBufs e;
  carve(bp, e, N, 1, S);
  S3D_CUDA(cudaMemsetAsync(e.sk.cnt, 0, SPLITK_COUNTERS * sizeof(unsigned), st));
  S3D_TRY(trunk_tc(m, img_slices, N, S, e.x0, e.ta, e.tb, e.x, e.xs, st, &e.sk));
  const int xc[6] = {0, 64, 128, 256, 512, 512};
  const size_t per_img = s3d_planes_bytes(1, K, S) / sizeof(float);
  float* pl = static_cast<float*>(planes);
  for (int i = 1; i <= 5; ++i) {
    const int s = 5 - i, R = plane_res(S, s), C = xc[i];
    Split x = split_of(e.xs[i], (size_t)N * R * R * C);
    for (int b = 0; b < B; ++b) {
      const size_t o = (size_t)b * K * R * R * C;
      S3D_TRY(conv_tc(m->tfcs[s], x.hi + o, x.lo + o, K, R, R, nullptr, 1, 0, pl + b * per_img + plane_offset_floats(K, S, s), 128,
                      nullptr, nullptr, 0, st));
    }
    if (taps_nchw && taps_nchw[i - 1]) {
      dim3 grid((R * R + 31) / 32, (C + 31) / 32, N), blk(32, 8);
      k_nhwc_to_nchw<<<grid, blk, 0, st>>>(e.x[i], taps_nchw[i - 1], R * R, C);
      S3D_LAUNCH_CHECK();
    }
  }
  return S3D_OK;
}

}  // namespace s3d
