// Shared host/device declarations of the slice3d_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <atomic>

#include "../../include/slice3d_b200.h"

namespace s3d {

// ---- error plumbing -------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<long long> g_launches;

#define S3D_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      s3d::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));               \
      return S3D_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

#define S3D_LAUNCH_CHECK()                                                              \
  do {                                                                                  \
    s3d::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      s3d::set_error(std::string("kernel launch: ") + cudaGetErrorString(_e));          \
      return S3D_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

// ---- model ------------------------------------------------------------------
// A convolution lowered to an implicit GEMM: w is [ks*ks*cin][ncols] fp32 (ncols
// contiguous); epilogue v = acc*scale[col] + shift[col] (scale may be null = 1).
struct ConvW {
  float* w = nullptr;
  float* scale = nullptr;
  float* shift = nullptr;
  int cin = 0, ncols = 0, ks = 1;
};

struct DecLayerF32 {
  const float *in_wt, *in_b;    // [128][384], [384]
  const float *out_wt, *out_b;  // [128][128], [128]
  const float *l1_wt, *l1_b;    // [128][2048], [2048]
  const float *l2_wt, *l2_b;    // [2048][128], [128]
  const float *n1_w, *n1_b, *n2_w, *n2_b;
};

struct DecF32 {
  const float* fcp_wt;  // [3][128]
  const float* fcp_b;   // [128]
  const float* fcs_b;   // [128]
  DecLayerF32 L[3];
  const float* fco_w;   // [128]
  const float* fco_b;   // [1]
};

// tcgen05 decoder weights: bf16 hi/lo "shared-memory images" (128B-swizzled K-major
// tiles, see decoder_tc.cu) plus fp32 bias / LayerNorm vectors.
struct DecTC {
  const __nv_bfloat16* wimg = nullptr;  // all tiles, hi then lo per tile
  const float* vec = nullptr;           // packed fp32 vectors
};

}  // namespace s3d

struct s3d_model {
  int device = 0;
  int K = 12;
  // encoder
  s3d::ConvW vgg[13];
  float* bn_scale[4] = {nullptr, nullptr, nullptr, nullptr};  // block-leading BNs (idx 4, 11, 21, 31)
  float* bn_shift[4] = {nullptr, nullptr, nullptr, nullptr};
  s3d::ConvW trans_c;        // 512 -> 512 part acting on x5
  float* trans_c_e = nullptr;  // [K][512] = W[:,512:] . emb_k + bias
  s3d::ConvW up_t[4];        // ConvTranspose2d as [cin][4*cout]
  s3d::ConvW dc1[4], dc2[4];
  s3d::ConvW trans_up[4];
  float* outc_w = nullptr;   // [3][32]
  float* outc_b = nullptr;
  s3d::ConvW fcs[5];         // fc_s hoisted per scale: [C_s][128]
  // decoder
  s3d::DecF32 dec32;
  s3d::DecTC dectc;
  std::vector<void*> allocs;
};

namespace s3d {

static inline int plane_res(int S, int s) { return (S / 16) << s; }
static const int kPlaneC[5] = {512, 256, 128, 64, 32};

// encoder.cu
int encoder_fwd(const s3d_model* m, const float* img, int B, int S, void* planes, float* const* feats_nchw,
                float* slices_rec, void* ws, size_t ws_bytes, cudaStream_t st);
size_t encoder_workspace_bytes(int B, int K, int S);

// decoder_simt.cu
int decoder_simt(const s3d_model* m, const float* planes, int S, float* qry, const s3d_grid* grid, int64_t first,
                 int64_t n, const float* T, const float* rot, int flip_in_place, float out_scale, float* out,
                 cudaStream_t st);

// decoder_tc.cu
int dectc_pack(s3d_model* m, const DecF32& src, cudaStream_t st);
size_t dectc_workspace_bytes(int64_t n);
int decoder_tc(const s3d_model* m, const float* planes, int S, float* qry, const s3d_grid* grid, int64_t first,
               int64_t n, const float* T, const float* rot, int flip_in_place, float out_scale, float* out,
               int precision, void* ws, size_t ws_bytes, cudaStream_t st);

// ---- device helpers shared by the decoders ---------------------------------------
struct QueryCtx {
  const float* qry;     // explicit points (n,3) or null
  float* qry_rw;        // same pointer when flip_in_place, else null
  int nx, ny, nz;       // grid (when qry == null)
  const float *px, *py, *pz;
  long long first;
  const float* T;       // (4,3)
  const float* rot;     // (3,3) or null
};

#ifdef __CUDACC__
// Query i -> model-space point (after the test-mode y,z flip or the train-mode rotation)
// and the clamped grid_sample coordinates (models.py:53-60, 28-36).
__device__ __forceinline__ void load_query(const QueryCtx& c, long long i, float& x, float& y, float& z, float& gu,
                                           float& gv) {
  if (c.qry) {
    x = c.qry[3 * i + 0];
    y = c.qry[3 * i + 1];
    z = c.qry[3 * i + 2];
  } else {
    long long g = c.first + i;
    int iz = (int)(g % c.nz);
    long long t = g / c.nz;
    int iy = (int)(t % c.ny);
    int ix = (int)(t / c.ny);
    x = c.px[ix];
    y = c.py[iy];
    z = c.pz[iz];
  }
  if (c.rot) {
    const float* R = c.rot;
    float rx = x * R[0] + y * R[3] + z * R[6];
    float ry = x * R[1] + y * R[4] + z * R[7];
    float rz = x * R[2] + y * R[5] + z * R[8];
    x = rx; y = ry; z = rz;
  } else {
    y = -y;
    z = -z;
  }
  const float* T = c.T;
  float pu = x * T[0] + y * T[3] + z * T[6] + T[9];
  float pv = x * T[1] + y * T[4] + z * T[7] + T[10];
  float pw = x * T[2] + y * T[5] + z * T[8] + T[11];
  gu = fminf(fmaxf(2.f * (pu / pw - 0.5f), -1.f), 1.f);
  gv = fminf(fmaxf(2.f * (pv / pw - 0.5f), -1.f), 1.f);
}

// grid_sample(bilinear, zeros, align_corners=True) tap set for one plane resolution R.
struct Taps {
  int o00, o01, o10, o11;  // pixel offsets (y*R + x) of the four taps (clamped in-bounds)
  float w00, w01, w10, w11;
};
__device__ __forceinline__ Taps make_taps(float gu, float gv, int R) {
  float ix = ((gu + 1.f) / 2.f) * (float)(R - 1);
  float iy = ((gv + 1.f) / 2.f) * (float)(R - 1);
  float fx = floorf(ix), fy = floorf(iy);
  int x0 = (int)fx, y0 = (int)fy;
  float ax = ix - fx, ay = iy - fy;  // weight of the +1 neighbours
  float bx = (fx + 1.f) - ix, by = (fy + 1.f) - iy;
  Taps t;
  bool x0ok = (x0 >= 0) && (x0 < R), x1ok = (x0 + 1 >= 0) && (x0 + 1 < R);
  bool y0ok = (y0 >= 0) && (y0 < R), y1ok = (y0 + 1 >= 0) && (y0 + 1 < R);
  int xc0 = min(max(x0, 0), R - 1), xc1 = min(max(x0 + 1, 0), R - 1);
  int yc0 = min(max(y0, 0), R - 1), yc1 = min(max(y0 + 1, 0), R - 1);
  t.o00 = yc0 * R + xc0; t.o01 = yc0 * R + xc1; t.o10 = yc1 * R + xc0; t.o11 = yc1 * R + xc1;
  t.w00 = (x0ok && y0ok) ? bx * by : 0.f;
  t.w01 = (x1ok && y0ok) ? ax * by : 0.f;
  t.w10 = (x0ok && y1ok) ? bx * ay : 0.f;
  t.w11 = (x1ok && y1ok) ? ax * ay : 0.f;
  return t;
}
#endif

}  // namespace s3d
