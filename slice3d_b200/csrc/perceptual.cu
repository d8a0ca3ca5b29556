// Frozen VGG19 perceptual loss on the tensor-core convolution (conv_tc.cu).
//
// reference: reg_slices/src/vgg_perceptual_loss.py:6-71 -- de-normalise to [0,1], ImageNet-normalise, the 14 3x3
// convolutions of torchvision's VGG19 up to conv5_2, taps after conv1_2, conv2_2, conv3_2, conv4_2 (POST-ReLU: the
// next slice's in-place ReLU overwrites the stored tensor) and conv5_2 (pre-ReLU), loss = sum_t w_t * mean|x_t - y_t|.
// The reference evaluates it on every forward, also at test time (src/models.py:90-92).
//
// Both image sets (the predicted slices and the targets) run as ONE batch of 2N images; the first convolution
// (3 input channels) is the fp32 CUDA-core GEMM, the other 13 are conv_tc (fp16 hi/lo split activations).  Taps are
// written as fp32 NHWC; the L1 terms are reduced in two deterministic stages (per-block double partials, then one block).
#include <cmath>

#include "gemm_simt.cuh"

namespace s3d {

namespace {

constexpr int L1_BLOCKS = 1024;
constexpr int L1_THREADS = 256;

// x in [-1,1] NCHW -> ((x+1)/2 - mean)/std as NHWC4; images [0,N) from a, [N,2N) from b
__global__ void k_vgg_input(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mean,
                            const float* __restrict__ stdv, float* __restrict__ out, int N, int HW) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2ll * N * HW) return;
  const int img = (int)(i / HW), p = (int)(i % HW);
  const float* s = (img < N ? a + (size_t)img * 3 * HW : b + (size_t)(img - N) * 3 * HW) + p;
  float v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = ((s[(size_t)c * HW] + 1.f) / 2.0f - mean[c]) / stdv[c];
  *reinterpret_cast<float4*>(out + i * 4) = make_float4(v[0], v[1], v[2], 0.f);
}

// 2x2 max pool of an fp32 NHWC tensor (already rectified), written in the split-fp16 format
__global__ void k_pool_split(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo, int NI, int H, int W,
                             int C) {
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)NI * Ho * Wo * C4) return;
  const int c = (int)(i % C4) * 4;
  long long t = i / C4;
  const int xo = (int)(t % Wo);
  t /= Wo;
  const int yo = (int)(t % Ho), n = (int)(t / Ho);
  float m[4];
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const float4 v = *reinterpret_cast<const float4*>(in + (((size_t)n * H + 2 * yo + dy) * W + 2 * xo + dx) * C + c);
      if (dy == 0 && dx == 0) {
        m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
      } else {
        m[0] = fmaxf(m[0], v.x); m[1] = fmaxf(m[1], v.y); m[2] = fmaxf(m[2], v.z); m[3] = fmaxf(m[3], v.w);
      }
    }
  store_split4(hi, lo, (((size_t)n * Ho + yo) * Wo + xo) * C + c, m);
}

// partial[tap][block] = sum over this block's share of |x - y| (x = first half of the tap, y = second half)
__global__ void __launch_bounds__(L1_THREADS) k_l1_partial(const float* __restrict__ tap, long long half_elems,
                                                           double* __restrict__ partial) {
  __shared__ double red[L1_THREADS];
  const float4* x = reinterpret_cast<const float4*>(tap);
  const float4* y = reinterpret_cast<const float4*>(tap + half_elems);
  const long long n4 = half_elems / 4;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 u = __ldg(x + i), v = __ldg(y + i);
    acc += (double)(fabsf(u.x - v.x) + fabsf(u.y - v.y)) + (double)(fabsf(u.z - v.z) + fabsf(u.w - v.w));
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = L1_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

struct L1Final {
  double inv_count[5];
  double weight[5];
};
// loss = sum_t weight_t * (sum of tap t's partials) / count_t, summed in a fixed order
__global__ void __launch_bounds__(L1_THREADS) k_l1_final(const double* __restrict__ partial, L1Final f, float* __restrict__ out) {
  __shared__ double red[L1_THREADS];
  double total = 0.0;
  for (int t = 0; t < 5; ++t) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < L1_BLOCKS; i += L1_THREADS) acc += partial[t * L1_BLOCKS + i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = L1_THREADS / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) total += f.weight[t] * red[0] * f.inv_count[t];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)total;
}

struct PBump {
  char* base;
  size_t off = 0;
  float* take(size_t floats) {
    const size_t o = off;
    off += ((floats * sizeof(float) + 255) / 256) * 256;
    return base ? reinterpret_cast<float*>(base + o) : nullptr;
  }
};

struct PBufs {
  float *x0, *sa, *sb, *tap[5], *tmpf;
  double* partial;
};

void pcarve(PBump& bp, PBufs& b, int N2, int S) {
  const size_t S2 = (size_t)S * S;
  b.x0 = bp.take((size_t)N2 * S2 * 4);
  b.sa = bp.take((size_t)N2 * S2 * 64);  // split ping-pong buffers (hi + lo = 4 bytes per element)
  b.sb = bp.take((size_t)N2 * S2 * 64);
  const int tc[5] = {64, 128, 256, 512, 512};
  for (int t = 0; t < 5; ++t) b.tap[t] = bp.take((size_t)N2 * (S2 >> (2 * t)) * tc[t]);
  b.tmpf = bp.take((size_t)N2 * (S2 >> 4) * 256);  // conv3_4 / conv4_4 outputs before their pools
  b.partial = reinterpret_cast<double*>(bp.take(5 * L1_BLOCKS * 2));
}

inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

struct Sp {
  __half *hi, *lo;
};
inline Sp sp(float* base, size_t elems) {
  __half* h = reinterpret_cast<__half*>(base);
  return Sp{h, h + elems};
}

}  // namespace

size_t vgg_loss_workspace_bytes(int N, int S) {
  PBump bp{nullptr, 0};
  PBufs b;
  pcarve(bp, b, 2 * N, S);
  return bp.off;
}

int vgg_loss_fwd(const s3d_model* m, const float* a, const float* b, int N, int S, float* loss, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
  if (!m->has_pvgg) {
    set_error("vgg_loss: the model was created without the vggptlossfunc.* tensors");
    return S3D_ERR_MISSING_TENSOR;
  }
  if (N <= 0 || S < 16 || (S % 16) != 0 || !a || !b || !loss) {
    set_error("vgg_loss: bad argument (S must be a multiple of 16)");
    return S3D_ERR_BAD_ARG;
  }
  if (ws == nullptr || ws_bytes < vgg_loss_workspace_bytes(N, S)) {
    set_error("vgg_loss: workspace too small");
    return S3D_ERR_WORKSPACE;
  }
  const int N2 = 2 * N;
  if ((long long)N2 * S * S >= (1ll << 31) / 4) {
    set_error("vgg_loss: N*S*S too large for 32-bit row indexing");
    return S3D_ERR_UNSUPPORTED;
  }
  PBump bp{static_cast<char*>(ws), 0};
  PBufs B;
  pcarve(bp, B, N2, S);
  k_vgg_input<<<nblk((long long)N2 * S * S, 256), 256, 0, st>>>(a, b, m->pvgg_mean, m->pvgg_std, B.x0, N, S * S);
  S3D_LAUNCH_CHECK();
  int H = S;
  auto elems = [&](int c) { return (size_t)N2 * H * H * c; };
  // conv index -> (cin, cout); 0: conv1_1 ... 13: conv5_2
  auto tc = [&](int i, Sp in, int relu, float* out_f32, Sp out, int cout) -> int {
    return conv_tc(m->tpvgg[i], in.hi, in.lo, N2, H, H, nullptr, 1, relu, out_f32, cout, out.hi, out.lo, cout, st);
  };
  const Sp none{nullptr, nullptr};
  auto pool = [&](const float* in, Sp out, int c) -> int {
    k_pool_split<<<nblk((long long)N2 * (H / 2) * (H / 2) * (c / 4), 256), 256, 0, st>>>(in, out.hi, out.lo, N2, H, H, c);
    S3D_LAUNCH_CHECK();
    H /= 2;
    return S3D_OK;
  };
  {  // conv1_1 on the fp32 path (K = 27), ReLU, written split
    const ConvW& w = m->pvgg[0];
    Sp o = sp(B.sa, elems(64));
    LoadConv L{B.x0, nullptr, N2 * H * H, w.k, H, H, 4, 0, 1, w.ks};
    EpiAffineSplit E{o.hi, o.lo, nullptr, w.shift, w.ncols, 1};
    S3D_TRY(launch_gemm(L, w.w, w.ncols, w.kpad, E, st));
  }
  S3D_TRY(tc(1, sp(B.sa, elems(64)), 1, B.tap[0], none, 64));                  // conv1_2 -> tap 1
  S3D_TRY(pool(B.tap[0], sp(B.sa, (size_t)N2 * (H / 2) * (H / 2) * 64), 64));
  S3D_TRY(tc(2, sp(B.sa, elems(64)), 1, nullptr, sp(B.sb, elems(128)), 128));  // conv2_1
  S3D_TRY(tc(3, sp(B.sb, elems(128)), 1, B.tap[1], none, 128));                // conv2_2 -> tap 2
  S3D_TRY(pool(B.tap[1], sp(B.sa, (size_t)N2 * (H / 2) * (H / 2) * 128), 128));
  S3D_TRY(tc(4, sp(B.sa, elems(128)), 1, nullptr, sp(B.sb, elems(256)), 256));           // conv3_1
  S3D_TRY(tc(5, sp(B.sb, elems(256)), 1, B.tap[2], sp(B.sa, elems(256)), 256));          // conv3_2 -> tap 3 (+ split for 3_3)
  S3D_TRY(tc(6, sp(B.sa, elems(256)), 1, nullptr, sp(B.sb, elems(256)), 256));           // conv3_3
  S3D_TRY(tc(7, sp(B.sb, elems(256)), 1, B.tmpf, none, 256));                            // conv3_4
  S3D_TRY(pool(B.tmpf, sp(B.sa, (size_t)N2 * (H / 2) * (H / 2) * 256), 256));
  S3D_TRY(tc(8, sp(B.sa, elems(256)), 1, nullptr, sp(B.sb, elems(512)), 512));           // conv4_1
  S3D_TRY(tc(9, sp(B.sb, elems(512)), 1, B.tap[3], sp(B.sa, elems(512)), 512));          // conv4_2 -> tap 4 (+ split)
  S3D_TRY(tc(10, sp(B.sa, elems(512)), 1, nullptr, sp(B.sb, elems(512)), 512));          // conv4_3
  S3D_TRY(tc(11, sp(B.sb, elems(512)), 1, B.tmpf, none, 512));                           // conv4_4
  S3D_TRY(pool(B.tmpf, sp(B.sa, (size_t)N2 * (H / 2) * (H / 2) * 512), 512));
  S3D_TRY(tc(12, sp(B.sa, elems(512)), 1, nullptr, sp(B.sb, elems(512)), 512));          // conv5_1
  S3D_TRY(tc(13, sp(B.sb, elems(512)), 0, B.tap[4], none, 512));                         // conv5_2 -> tap 5 (pre-ReLU)
  // L1 terms
  const int tcn[5] = {64, 128, 256, 512, 512};
  const double wts[5] = {1.0 / 2.6, 1.0 / 4.8, 1.0 / 3.7, 1.0 / 5.6, 10.0 / 1.5};
  L1Final f;
  for (int t = 0; t < 5; ++t) {
    const long long half = (long long)N * (S >> t) * (S >> t) * tcn[t];
    k_l1_partial<<<L1_BLOCKS, L1_THREADS, 0, st>>>(B.tap[t], half, B.partial + t * L1_BLOCKS);
    S3D_LAUNCH_CHECK();
    f.inv_count[t] = 1.0 / (double)half;
    f.weight[t] = wts[t];
  }
  k_l1_final<<<1, L1_THREADS, 0, st>>>(B.partial, f, loss);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

// ================================================================================================================
// Training: the same loss with its gradient with respect to the first image batch (reg_slices/train.py:41-53:
// loss.backward() through ret['vgg_loss'], models.py:90-92).  The network is frozen, so the backward pass is data
// gradients only, and the data gradient of a 3x3 convolution is a 3x3 convolution of the output gradient with the
// rotated, transposed weights: the SAME tcgen05 kernel runs it (weights packed once at creation, api.cu:build_pvgg).
//
//   forward   as vgg_loss_fwd, every activation kept in fp32 NHWC (post-ReLU; conv5_2 pre-ReLU) for both halves
//   backward  tap gradient  G * w_t / count_t * sign(x_t - y_t)  injected at the five taps; per layer (top down):
//             acc = conv_tc(dgrad weights, g)  ->  one elementwise kernel: route through the 2x2 max pool (first maximum in
//             scan order, as ATen), add the tap term, mask with the ReLU (activation > 0), write g of the layer below in the
//             split-fp16 GEMM format.  conv1_1's 3-channel data gradient runs on the fp32 CUDA-core GEMM.
// G is a power of two that lifts the gradients (1e-9 .. 1e-5 in real units) into fp16's normal range; it is divided out
// at the end together with the input normalisation d/dx = 0.5 / std.
namespace {

struct VInfo {
  int c[14], lvl[14];  // channels and resolution level (H = S >> lvl) of layer i's output
  int tap_of[14];      // tap index of layer i (or -1)
};
VInfo vinfo() {
  VInfo v;
  const int c[14] = {64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512, 512};
  const int l[14] = {0, 0, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4};
  const int t[14] = {-1, 0, -1, 1, -1, 2, -1, -1, -1, 3, -1, -1, -1, 4};
  for (int i = 0; i < 14; ++i) {
    v.c[i] = c[i];
    v.lvl[i] = l[i];
    v.tap_of[i] = t[i];
  }
  return v;
}

struct TBufs {
  float *x0, *sa, *sb, *A[14], *acc, *acc2;
  double* partial;
};
void tcarve(PBump& bp, TBufs& b, int N2, int S) {
  const size_t S2 = (size_t)S * S;
  const VInfo v = vinfo();
  b.x0 = bp.take((size_t)N2 * S2 * 4);
  b.sa = bp.take((size_t)N2 * S2 * 64);
  b.sb = bp.take((size_t)N2 * S2 * 64);
  for (int i = 0; i < 14; ++i) b.A[i] = bp.take((size_t)N2 * (S2 >> (2 * v.lvl[i])) * v.c[i]);
  b.acc = bp.take((size_t)(N2 / 2) * S2 * 64);   // data gradient of a layer (fp32), largest: 64 channels at S x S
  b.acc2 = bp.take((size_t)(N2 / 2) * S2 * 64);
  b.partial = reinterpret_cast<double*>(bp.take(5 * L1_BLOCKS * 2));
}

// fp32 NHWC -> split fp16
__global__ void k_to_split(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = *reinterpret_cast<const float4*>(in + i * 4);
  const float r[4] = {v.x, v.y, v.z, v.w};
  store_split4(hi, lo, (size_t)i * 4, r);
}

// g[n][y][x][c] = mask * (routed acc + tap term), 4 channels per thread.
//   acc      data gradient of the layer above w.r.t. its input: [N][H >> pooled][W >> pooled][C] (null at the top)
//   A        this layer's activation, [2N][H][W][C]: images [0,N) = first batch, [N,2N) = second
//   tw       G * w_t / count_t if this layer is a tap, else 0;  relu: mask by A > 0
__global__ void __launch_bounds__(256) k_vgg_bwd_elem(const float* __restrict__ acc, int pooled, const float* __restrict__ A,
                                                      float tw, int relu, int N, int H, int W, int C, __half* __restrict__ ghi,
                                                      __half* __restrict__ glo, float* __restrict__ gf32) {
  const int C4 = C / 4;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * H * W * C4) return;
  const int c = (int)(i % C4) * 4;
  long long t = i / C4;
  const int x = (int)(t % W);
  t /= W;
  const int y = (int)(t % H), n = (int)(t / H);
  const size_t o = (((size_t)n * H + y) * W + x) * C + c;
  const float4 a4 = *reinterpret_cast<const float4*>(A + o);
  const float a[4] = {a4.x, a4.y, a4.z, a4.w};
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  if (acc) {
    if (!pooled) {
      const float4 v = *reinterpret_cast<const float4*>(acc + o);
      g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
    } else {
      const int y0 = y & ~1, x0 = x & ~1, me = (y & 1) * 2 + (x & 1);
      float w4[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(A + (((size_t)n * H + y0 + (q >> 1)) * W + x0 + (q & 1)) * C + c);
        w4[q][0] = v.x; w4[q][1] = v.y; w4[q][2] = v.z; w4[q][3] = v.w;
      }
      const float4 up = *reinterpret_cast<const float4*>(acc + (((size_t)n * (H / 2) + (y >> 1)) * (W / 2) + (x >> 1)) * C + c);
      const float u[4] = {up.x, up.y, up.z, up.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int best = 0;  // first maximum in scan order (0,0), (0,1), (1,0), (1,1): ATen's max_pool2d backward
        float bv = w4[0][j];
#pragma unroll
        for (int q = 1; q < 4; ++q)
          if (w4[q][j] > bv) {
            bv = w4[q][j];
            best = q;
          }
        g[j] = best == me ? u[j] : 0.f;
      }
    }
  }
  if (tw != 0.f) {
    const float4 b4 = *reinterpret_cast<const float4*>(A + o + (size_t)N * H * W * C);
    const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] += a[j] > b[j] ? tw : (a[j] < b[j] ? -tw : 0.f);
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] = a[j] > 0.f ? g[j] : 0.f;
  }
  store_split4(ghi, glo, o, g);
  if (gf32) *reinterpret_cast<float4*>(gf32 + o) = make_float4(g[0], g[1], g[2], g[3]);
}

// grad_a[n][c][p] = dx[n][p][c] * 0.5 / std[c] * gout / G      (dx: [N*HW][4] from the conv1_1 data-gradient GEMM)
__global__ void k_vgg_bwd_out(const float* __restrict__ dx, const float* __restrict__ stdv, const float* __restrict__ gout,
                              float inv_g, float* __restrict__ out, int N, int HW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * HW) return;
  const int n = (int)(i / HW), p = (int)(i % HW);
  const float4 v = *reinterpret_cast<const float4*>(dx + i * 4);
  const float s = gout[0] * inv_g * 0.5f;
  out[((size_t)n * 3 + 0) * HW + p] = v.x * s / stdv[0];
  out[((size_t)n * 3 + 1) * HW + p] = v.y * s / stdv[1];
  out[((size_t)n * 3 + 2) * HW + p] = v.z * s / stdv[2];
}

const double kTapW[5] = {1.0 / 2.6, 1.0 / 4.8, 1.0 / 3.7, 1.0 / 5.6, 10.0 / 1.5};

}  // namespace

size_t vgg_loss_train_bytes(int N, int S) {
  PBump bp{nullptr, 0};
  TBufs b;
  tcarve(bp, b, 2 * N, S);
  return bp.off;
}

int vgg_loss_train_fwd(const s3d_model* m, const float* a, const float* b, int N, int S, float* loss, void* saved,
                       size_t saved_bytes, cudaStream_t st) {
  if (!m->has_pvgg) {
    set_error("vgg_loss_train: the model was created without the vggptlossfunc.* tensors");
    return S3D_ERR_MISSING_TENSOR;
  }
  if (N <= 0 || S < 16 || (S % 16) != 0 || !a || !b || !loss || !saved || saved_bytes < vgg_loss_train_bytes(N, S)) {
    set_error("vgg_loss_train: bad argument (S must be a multiple of 16) or saved buffer too small");
    return S3D_ERR_BAD_ARG;
  }
  const int N2 = 2 * N;
  if ((long long)N2 * S * S >= (1ll << 31) / 4) {
    set_error("vgg_loss_train: N*S*S too large for 32-bit row indexing");
    return S3D_ERR_UNSUPPORTED;
  }
  PBump bp{static_cast<char*>(saved), 0};
  TBufs B;
  tcarve(bp, B, N2, S);
  const VInfo v = vinfo();
  k_vgg_input<<<nblk((long long)N2 * S * S, 256), 256, 0, st>>>(a, b, m->pvgg_mean, m->pvgg_std, B.x0, N, S * S);
  S3D_LAUNCH_CHECK();
  auto HH = [&](int i) { return S >> v.lvl[i]; };
  auto elems = [&](int i) { return (size_t)N2 * HH(i) * HH(i) * v.c[i]; };
  {  // conv1_1 on the fp32 path: activation in fp32, then its split copy
    const ConvW& w = m->pvgg[0];
    LoadConv L{B.x0, nullptr, N2 * S * S, w.k, S, S, 4, 0, 1, w.ks};
    EpiAffine E{B.A[0], nullptr, w.shift, w.ncols, 1};
    S3D_TRY(launch_gemm(L, w.w, w.ncols, w.kpad, E, st));
    Sp o = sp(B.sa, elems(0));
    k_to_split<<<nblk((long long)elems(0) / 4, 256), 256, 0, st>>>(B.A[0], o.hi, o.lo, (long long)elems(0) / 4);
    S3D_LAUNCH_CHECK();
  }
  float* cur = B.sa;  // split input of the next convolution (ping-pong sa / sb)
  float* oth = B.sb;
  size_t cur_elems = elems(0);
  for (int i = 1; i < 14; ++i) {
    const int H = HH(i);
    if (v.lvl[i] != v.lvl[i - 1]) {  // 2x2 max pool of the previous (rectified) activation, written split
      Sp o = sp(oth, (size_t)N2 * H * H * v.c[i - 1]);
      k_pool_split<<<nblk((long long)N2 * H * H * (v.c[i - 1] / 4), 256), 256, 0, st>>>(B.A[i - 1], o.hi, o.lo, N2, 2 * H, 2 * H,
                                                                                     v.c[i - 1]);
      S3D_LAUNCH_CHECK();
      float* t = cur; cur = oth; oth = t;
      cur_elems = (size_t)N2 * H * H * v.c[i - 1];
    }
    Sp in = sp(cur, cur_elems);
    const bool next_pooled = i < 13 && v.lvl[i + 1] != v.lvl[i];
    const bool need_split = i < 13 && !next_pooled;
    Sp out = need_split ? sp(oth, elems(i)) : Sp{nullptr, nullptr};
    S3D_TRY(conv_tc(m->tpvgg[i], in.hi, in.lo, N2, H, H, nullptr, 1, i < 13 ? 1 : 0, B.A[i], v.c[i], out.hi, out.lo, v.c[i], st));
    if (need_split) {
      float* t = cur; cur = oth; oth = t;
      cur_elems = elems(i);
    }
  }
  L1Final f;
  for (int t = 0, i = 0; i < 14; ++i) {
    if (v.tap_of[i] < 0) continue;
    const long long half = (long long)N * HH(i) * HH(i) * v.c[i];
    k_l1_partial<<<L1_BLOCKS, L1_THREADS, 0, st>>>(B.A[i], half, B.partial + t * L1_BLOCKS);
    S3D_LAUNCH_CHECK();
    f.inv_count[t] = 1.0 / (double)half;
    f.weight[t] = kTapW[t];
    ++t;
  }
  k_l1_final<<<1, L1_THREADS, 0, st>>>(B.partial, f, loss);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

int vgg_loss_train_bwd(const s3d_model* m, int N, int S, const float* gout, void* saved, size_t saved_bytes, float* grad_a,
                       cudaStream_t st) {
  if (!m->has_pvgg || N <= 0 || !gout || !saved || !grad_a || saved_bytes < vgg_loss_train_bytes(N, S)) {
    set_error("vgg_loss_train_bwd: bad argument");
    return S3D_ERR_BAD_ARG;
  }
  PBump bp{static_cast<char*>(saved), 0};
  TBufs B;
  tcarve(bp, B, 2 * N, S);
  const VInfo v = vinfo();
  auto HH = [&](int i) { return S >> v.lvl[i]; };
  // gradient scale: the smallest tap gradient (tap 1: w / count) lands at 2^-4
  double gmin = 1e30;
  double tw[5];
  for (int t = 0, i = 0; i < 14; ++i)
    if (v.tap_of[i] >= 0) {
      tw[t] = kTapW[t] / ((double)N * HH(i) * HH(i) * v.c[i]);
      gmin = tw[t] < gmin ? tw[t] : gmin;
      ++t;
    }
  int ex;
  std::frexp(0.0625 / gmin, &ex);
  const float G = std::ldexp(1.f, ex - 1);
  // layer 13 (tap 5, pre-ReLU): g = tap term only
  float* gcur = B.sa;  // split gradient buffers reuse the forward's ping-pong areas
  float* goth = B.sb;
  auto gsp = [&](float* base, int i) { return sp(base, (size_t)N * HH(i) * HH(i) * v.c[i]); };
  auto elem = [&](const float* acc, int pooled, int i, float* gbase, float* gf32) -> int {
    const int H = HH(i), t = v.tap_of[i];
    Sp g = gsp(gbase, i);
    const long long n4 = (long long)N * H * H * (v.c[i] / 4);
    k_vgg_bwd_elem<<<nblk(n4, 256), 256, 0, st>>>(acc, pooled, B.A[i], t >= 0 ? (float)(tw[t] * (double)G) : 0.f, i < 13 ? 1 : 0, N,
                                                 H, H, v.c[i], g.hi, g.lo, gf32);
    S3D_LAUNCH_CHECK();
    return S3D_OK;
  };
  S3D_TRY(elem(nullptr, 0, 13, gcur, nullptr));
  for (int i = 13; i >= 1; --i) {
    const int H = HH(i);  // layer i's input has resolution H (after the pool, if any) and c[i-1] channels
    Sp g = gsp(gcur, i);
    S3D_TRY(conv_tc(m->tpvgg_d[i], g.hi, g.lo, N, H, H, nullptr, 1, 0, B.acc, v.c[i - 1], nullptr, nullptr, 0, st));
    const int pooled = v.lvl[i] != v.lvl[i - 1];
    S3D_TRY(elem(B.acc, pooled, i - 1, goth, i == 1 ? B.acc2 : nullptr));
    float* t = gcur; gcur = goth; goth = t;
  }
  {  // conv1_1: 64 -> 3 (+1 pad) channels on the fp32 CUDA-core GEMM, from the fp32 copy of g0
    const ConvW& w = m->pvgg_d[0];
    LoadConv L{B.acc2, nullptr, N * S * S, w.k, S, S, 64, 0, 1, w.ks};
    EpiAffine E{B.acc, nullptr, nullptr, w.ncols, 0};
    S3D_TRY(launch_gemm(L, w.w, w.ncols, w.kpad, E, st));
    k_vgg_bwd_out<<<nblk((long long)N * S * S, 256), 256, 0, st>>>(B.acc, m->pvgg_std, gout, 1.f / G, grad_a, N, S * S);
    S3D_LAUNCH_CHECK();
  }
  return S3D_OK;
}

}  // namespace s3d
