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

}  // namespace s3d
