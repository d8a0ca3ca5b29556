// Input pipeline on the device: what Slice3DDataset.__getitem__ does to every PNG after
// decoding it (reference: reg_slices/src/datasets.py:75-88 png_2_whitebg / png_2_rgb, :37 preprocess =
// T.Resize((S, S)) -> T.ToTensor() -> T.Normalize(0.5, 0.5), :98-118 the 1 + 12 images of a sample), for a whole
// batch of decoded RGBA images at once:
//
//   composite   white background: rgb where alpha != 0 else 255;  black: trunc(rgb * (alpha / 255.0)) in float64
//   resize      Pillow's antialiased bilinear resample of 8-bit images, bit for bit: two separable passes (horizontal,
//               then vertical) with 22-bit fixed-point coefficients, an 8-bit intermediate image and round-half-up
//               (ImagingResampleHorizontal_8bpc / Vertical_8bpc); the coefficient tables are computed on the host in
//               float64 exactly as precompute_coeffs / normalize_coeffs_8bpc do (slice3d_b200/inputs.py)
//   to tensor   u8 / 255 (fp32), then (x - 0.5) / 0.5, written NCHW
//
// Two kernels (the 8-bit intermediate between the passes is part of the reference's arithmetic).  HBM-bound byte work:
// 4 B/pixel in, 12 B/pixel out; one thread per output element, channel-interleaved reads coalesce over x.
#include "common.cuh"

namespace s3d {

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ unsigned char clip8(int v) {
  v >>= PRECISION_BITS;
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__device__ __forceinline__ unsigned char composite(const unsigned char* px, int c, int white_bg) {
  const unsigned char a = px[3], v = px[c];
  if (white_bg) return a == 0 ? (unsigned char)255 : v;
  return (unsigned char)((double)v * ((double)a / 255.0));
}

// tmp[n][y][xx][c] = horizontal resample of the composited image; bounds (xmin, count) and ksize coefficients per xx
__global__ void __launch_bounds__(256) k_prep_h(const unsigned char* __restrict__ rgba, int N, int H, int W, int S,
                                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                int white_bg, unsigned char* __restrict__ tmp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * H * S * 3) return;
  const int c = (int)(i % 3);
  long long t = i / 3;
  const int xx = (int)(t % S);
  t /= S;
  const int y = (int)(t % H), n = (int)(t / H);
  const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
  const int* k = kk + (size_t)xx * ksize;
  const unsigned char* row = rgba + ((size_t)n * H + y) * W * 4;
  int ss = 1 << (PRECISION_BITS - 1);
  for (int x = 0; x < cnt; ++x) ss += (int)composite(row + (size_t)(xmin + x) * 4, c, white_bg) * k[x];
  tmp[i] = clip8(ss);
}

// out[n][c][yy][xx] = ((vertical resample) / 255 - 0.5) / 0.5
__global__ void __launch_bounds__(256) k_prep_v(const unsigned char* __restrict__ tmp, int N, int H, int S,
                                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * 3 * S * S) return;
  const int xx = (int)(i % S);
  long long t = i / S;
  const int yy = (int)(t % S);
  t /= S;
  const int c = (int)(t % 3), n = (int)(t / 3);
  const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
  const int* k = kk + (size_t)yy * ksize;
  int ss = 1 << (PRECISION_BITS - 1);
  for (int y = 0; y < cnt; ++y) ss += (int)tmp[(((size_t)n * H + ymin + y) * S + xx) * 3 + c] * k[y];
  const float v = __fdiv_rn((float)clip8(ss), 255.f);
  out[i] = __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
}

}  // namespace

size_t preprocess_workspace_bytes(int N, int H, int S) { return (size_t)N * H * S * 3 + 256; }

int preprocess_rgba(const unsigned char* rgba, int N, int H, int W, int S, int white_bg, const int* bounds_h, const int* kk_h,
                    int ksize_h, const int* bounds_v, const int* kk_v, int ksize_v, float* out, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
  if (!rgba || !out || !bounds_h || !kk_h || !bounds_v || !kk_v || N < 1 || H < 1 || W < 1 || S < 1 || ksize_h < 1 ||
      ksize_v < 1) {
    set_error("preprocess: bad argument");
    return S3D_ERR_BAD_ARG;
  }
  if (!ws || ws_bytes < preprocess_workspace_bytes(N, H, S)) {
    set_error("preprocess: workspace too small");
    return S3D_ERR_WORKSPACE;
  }
  unsigned char* tmp = static_cast<unsigned char*>(ws);
  const long long n1 = (long long)N * H * S * 3, n2 = (long long)N * 3 * S * S;
  k_prep_h<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(rgba, N, H, W, S, bounds_h, kk_h, ksize_h, white_bg, tmp);
  S3D_LAUNCH_CHECK();
  k_prep_v<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(tmp, N, H, S, bounds_v, kk_v, ksize_v, out);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

}  // namespace s3d
