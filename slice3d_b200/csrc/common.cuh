// Shared host/device declarations of the slice3d_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>
#include <vector>

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

#define S3D_TRY(expr)                                                                   \
  do {                                                                                  \
    int _r = (expr);                                                                    \
    if (_r != S3D_OK) return _r;                                                        \
  } while (0)

// ---- model ------------------------------------------------------------------
// A convolution / linear layer lowered to an implicit GEMM: w is [kpad][ncols] fp32
// (ncols contiguous, rows >= k zero); epilogue v = acc*scale[col] + shift[col]
// (scale null = 1, shift null = 0).
struct ConvW {
  float* w = nullptr;
  float* scale = nullptr;
  float* shift = nullptr;
  int cin = 0;    // input channels as stored in the NHWC activation (after padding to 4)
  int ncols = 0;  // GEMM N
  int ks = 1;     // 1 or 3
  int k = 0, kpad = 0;
};

// A convolution lowered to the tcgen05 implicit GEMM (conv_tc.cu): fp16 hi/lo weight tiles (of w * 2^e, so that
// the lo parts are normal fp16 numbers; acc_scale = 2^-e undoes it) in the UMMA shared-memory layout;
// scale/shift alias the fp32 epilogue vectors of the ConvW it was packed from.
struct ConvTC {
  uint8_t* wimg = nullptr;
  int cinp = 0;     // input channels padded to a multiple of 64 (pitch of the split NHWC input)
  int cout = 0;     // real output channels
  int bn = 128;     // output-channel tile (64 or 128)
  int n_tiles = 0;
  int taps = 9;
  float acc_scale = 1.f;
  const float* scale = nullptr;
  const float* shift = nullptr;
};

struct DecLayerF32 {
  ConvW in_proj;   // 128 -> 384, shift = in_proj_bias
  ConvW out_proj;  // 128 -> 128
  ConvW lin1;      // 128 -> 2048
  ConvW lin2;      // 2048 -> 128
  float *n1_w = nullptr, *n1_b = nullptr, *n2_w = nullptr, *n2_b = nullptr;
};

struct DecF32 {
  float* fcp_wt = nullptr;  // [3][128]
  float* fcp_b = nullptr;   // [128]
  float* fcs_b = nullptr;   // [128]
  DecLayerF32 L[3];
  float* fco_w = nullptr;   // [128]
  float* fco_b = nullptr;   // [1]
};

// tcgen05 decoder weights (decoder_tc.cu): bf16 hi/lo operand images in the canonical
// K-major 128B-swizzled shared-memory layout, plus packed fp32 vectors.
struct DecTC {
  __nv_bfloat16* wimg = nullptr;  // bf16 hi/lo pairs
  __half* wimg_h = nullptr;        // fp16 hi/lo pairs (S3D_PREC_FP16X3)
  uint8_t* wimg_f8 = nullptr;      // fp16 pairs (attention) + fp16 / e4m3 / e4m3 (FFN) (S3D_PREC_FP16F8)
  float* vec = nullptr;
  size_t wimg_elems = 0;
};

}  // namespace s3d

struct s3d_model {
  int device = 0;
  int K = 12;
  int kind = 0;                // 0 = Slices3DRegModel, 1 = Slices3DGTModel (model_gt.py)
  // GT model only: second Linear of fc_local (128 -> 128, shift = bias) and the query MLP (PyTorch (out, in) layouts)
  s3d::ConvW fcl2;
  s3d::ConvTC tfcl2;
  float* pts_w[3] = {};
  float* pts_b[3] = {};
  // encoder (all BatchNorms are eval-mode and folded)
  s3d::ConvW vgg[13];          // conv idx 0,3 | 7,10 | 14,17,20 | 24,27,30 | 34,37,40
  float* bn_scale[4] = {};     // block-leading BNs (features idx 4, 11, 21, 31) applied to the raw taps
  float* bn_shift[4] = {};
  s3d::ConvW trans_c;          // 512 -> 512 part acting on x5 (no bias)
  float* trans_c_e = nullptr;  // [K][512] = W[:,512:] . emb_k + bias
  s3d::ConvW up_t[4];          // ConvTranspose2d as [cin][4*cout]; shift = bias[cout]
  s3d::ConvW dc1[4], dc2[4];   // DoubleConv (BN folded, ReLU)
  s3d::ConvW trans_up[4];      // 1x1 skip adapters
  float* outc_w = nullptr;     // [3][32]
  float* outc_b = nullptr;     // [3]
  s3d::ConvW fcs[5];           // fc_s hoisted per scale: [C_s][128], no bias
  // tensor-core encoder (conv_tc.cu): the 3x3 convolutions; dc1 is split into its slice-independent skip half
  // (dc1s, run once per view on the fp32 path) and its per-slice half (tdc1)
  s3d::ConvTC tvgg[13];        // [0] unused (3 input channels: fp32 path)
  s3d::ConvTC tdc1[4], tdc2[4], tdc1s[4];
  s3d::ConvTC ttrans_c, ttrans_up[4], tup_t[4], tfcs[5];  // the 1x1 / transposed convolutions and the fc_s projection
  s3d::ConvW dc1s[4];
  // VGG19 perceptual loss (perceptual.cu); optional: present when the vggptlossfunc.* tensors were given
  int has_pvgg = 0;
  s3d::ConvW pvgg[14];         // conv1_1 .. conv5_2 (shift = bias)
  s3d::ConvTC tpvgg[14];       // [0] unused
  s3d::ConvW pvgg_d[14];       // data-gradient convolutions of the backward pass (rotated, transposed weights)
  s3d::ConvTC tpvgg_d[14];     // [0] unused (3 output channels: fp32 path)
  float* pvgg_mean = nullptr;
  float* pvgg_std = nullptr;
  int enc_simt = 0;            // s3d_debug_set_encoder(m, 1): whole encoder on the fp32 CUDA-core path (debugging)
  // decoder
  s3d::DecF32 dec32;
  s3d::DecTC dectc;
  std::vector<void*> allocs;
};

namespace s3d {

__host__ __device__ static inline int plane_res(int S, int s) { return (S / 16) << s; }
static const int kPlaneC[5] = {512, 256, 128, 64, 32};

// Offset (in floats) of scale s inside one image's projected-plane blob: per scale
// (K, R_s, R_s, 128) fp32 channels-last.
static inline size_t plane_offset_floats(int K, int S, int s) {
  size_t o = 0;
  for (int i = 0; i < s; ++i) o += (size_t)K * plane_res(S, i) * plane_res(S, i) * 128;
  return o;
}

// encoder.cu
int encoder_fwd(const s3d_model* m, const float* img, int B, int S, void* planes, float* const* feats_nchw,
                float* slices_rec, void* ws, size_t ws_bytes, cudaStream_t st);
size_t encoder_workspace_bytes(int B, int K, int S);

// conv_tc.cu
int convtc_pack(s3d_model* m, const ConvW& cw, int src_cin, int ci0, int cin, ConvTC& out, cudaStream_t st);
// Split-K scratch of the tensor-core convolutions: partial tiles + per-tile arrival counters.  The owner zeroes the
// counters before the first launch that uses them; the kernel leaves them zero.
struct SplitK {
  float* part = nullptr;
  size_t part_bytes = 0;
  unsigned* cnt = nullptr;
  int n_cnt = 0;
};
constexpr size_t SPLITK_PART_BYTES = 16u << 20;
constexpr int SPLITK_COUNTERS = 1024;
int conv_tc(const ConvTC& w, const __half* in_hi, const __half* in_lo, int NI, int H, int W, const float* add,
            int add_div, int relu, float* out_f32, int ldf, __half* out_hi, __half* out_lo, int lds, cudaStream_t st,
            int shuffle_c = 0, const SplitK* sk = nullptr);
int enctc_pack(s3d_model* m, cudaStream_t st);

// mcubes.cu
int mc_count(const double* vol, int nx, int ny, int nz, double iso, const int* tri_count, int* vcount, int* tcount,
             unsigned char* owned, cudaStream_t st);
int mc_emit(const double* vol, int nx, int ny, int nz, double iso, const signed char* table, const long long* vbase,
            const long long* tbase, const int* tcount, const unsigned char* owned, double* verts, long long* tris,
            cudaStream_t st);
size_t scan_scratch_bytes(long long n);
int exclusive_scan_i32_i64(const int* in, long long n, long long* out, long long* total, void* scratch, cudaStream_t st);

// mise.cu
size_t mise_scratch_ints(int res0, int depth);
int mise_subdivide(int res0, int depth, double thr, const double* value, const unsigned char* known, signed char* cell_level,
                   unsigned char* exists, int* flags_zeroed, cudaStream_t st);

int mise_query_device(int R, double box, const unsigned char* exists, const unsigned char* known, int* blk, int* count, int cap,
                      int* round_count, int* overflow, int* pt_idx, float* pts, cudaStream_t st);
size_t mise_query_blocks(int R);
int mise_apply_device(const int* count, const int* pt_idx, const float* vals, double* value, unsigned char* known,
                      cudaStream_t st);

// inputs.cu
size_t preprocess_workspace_bytes(int N, int H, int S);
int preprocess_rgba(const unsigned char* rgba, int N, int H, int W, int S, int white_bg, const int* bounds_h, const int* kk_h,
                    int ksize_h, const int* bounds_v, const int* kk_v, int ksize_v, float* out, void* ws, size_t ws_bytes,
                    cudaStream_t st);

// perceptual.cu
size_t vgg_loss_workspace_bytes(int N, int S);
int vgg_loss_fwd(const s3d_model* m, const float* a, const float* b, int N, int S, float* loss, void* ws, size_t ws_bytes,
                 cudaStream_t st);
size_t vgg_loss_train_bytes(int N, int S);
int vgg_loss_train_fwd(const s3d_model* m, const float* a, const float* b, int N, int S, float* loss, void* saved,
                       size_t saved_bytes, cudaStream_t st);
int vgg_loss_train_bwd(const s3d_model* m, int N, int S, const float* gout, void* saved, size_t saved_bytes, float* grad_a,
                       cudaStream_t st);

// train_decoder.cu
size_t train_decoder_saved_bytes(const s3d_train_cfg* c);
size_t train_decoder_bwd_workspace_bytes(const s3d_train_cfg* c);
int train_decoder_fwd(const s3d_train_cfg* c, const float* const* feats, const float* qry, const float* T,
                      const float* const* params, float* sdf, void* saved, size_t saved_bytes, cudaStream_t st);
int train_decoder_bwd(const s3d_train_cfg* c, const float* qry, const float* T, const float* const* params, const float* dsdf,
                      void* saved, size_t saved_bytes, float* const* dfeats, float* const* dparams, void* ws, size_t ws_bytes,
                      cudaStream_t st);

// ---- queries ----------------------------------------------------------------------
struct QueryCtx {
  const float* qry;     // explicit points (n,3) or null
  int preflipped;       // 1: explicit points already carry the test-mode y,z flip (api.cu flips the
                        // caller's tensor in place first, reproducing models.py:55's side effect)
  int nx, ny, nz;       // grid (when qry == null)
  const float *px, *py, *pz;
  long long first;
  const float* T;       // (4,3)
  const float* rot;     // (3,3) or null
  // Locality order of a grid range made of whole x-planes [x0, x0 + nxs) (blk = 1): the i-th query of the launch is
  // not flat index first + i but the i-th point of a walk over GRID_BX x GRID_BY blocks of (x, y) columns, z fastest
  // inside a column.  The columns of a block project to neighbouring rays of every plane, so the texels a wave of CTAs
  // gathers stay in L2 until the neighbouring columns need them (the flat order re-read ~260 MB per x-plane).
  int blk, x0, nxs;
  // Ready tokens (Slices3DGTModel: built by gt.cu before the decoder launch): query token [n][128] and slice tokens
  // [n][K][128]; the decoders then skip their own token build.
  const float* tok_query;
  const float* tok_slice;
  // Batched explicit points (per_img > 0): query i belongs to image i / per_img, whose camera is T + 12 b, rotation
  // rot + 9 b and projected planes planes + b * plane_stride floats (the encoder's batch layout).
  long long per_img;
  size_t plane_stride;
};
#ifndef S3D_GRID_B
#define S3D_GRID_B 16  // (overridable for locality experiments: profiles/r2_summary.md)
#endif
constexpr int GRID_BX = S3D_GRID_B, GRID_BY = S3D_GRID_B;

// gt.cu (Slices3DGTModel)
size_t gt_encoder_workspace_bytes(int N, int S);
int gt_encoder_fwd(const s3d_model* m, const float* img_slices, int B, int S, void* planes, float* const* taps_nchw, void* ws,
                   size_t ws_bytes, cudaStream_t st);
size_t gt_decoder_workspace_bytes(int64_t n, int precision);
int gt_decoder_fwd(const s3d_model* m, const void* planes, int S, QueryCtx q, int64_t n, float out_scale, float* out,
                   int precision, void* ws, size_t ws_bytes, cudaStream_t st);
int trunk_tc(const s3d_model* m, const float* img, int B, int S, float* x0, float* ta, float* tb, float* const* x,
             float* const* xs, cudaStream_t st, const SplitK* sk = nullptr);

// decoder_simt.cu
size_t decoder_simt_workspace_bytes(int64_t n);
int decoder_simt(const s3d_model* m, const float* planes, int S, const QueryCtx& q, int64_t n, float out_scale,
                 float* out, float* debug_tokens, void* ws, size_t ws_bytes, cudaStream_t st);

// decoder_tc.cu
int dectc_pack(s3d_model* m, cudaStream_t st);
bool decoder_tc_supported(const s3d_model* m);
void decoder_tc_set_debug(int flags);
size_t decoder_tc_workspace_bytes(int64_t n);
int decoder_tc(const s3d_model* m, const float* planes, int S, const QueryCtx& q, int64_t n, float out_scale,
               float* out, int precision, void* ws, size_t ws_bytes, cudaStream_t st, const int* n_dev = nullptr);

int debug_profile(long long* out32, int reset);
int umma_selftest(int mode, int passes, const float* a_dev, const float* w_dev, float* d_dev, cudaStream_t st);

#ifdef __CUDACC__
// i-th query of a grid launch -> (ix, iy, iz) and its position in the launch's output range.
__device__ __forceinline__ long long grid_point(const QueryCtx& c, long long i, int& ix, int& iy, int& iz) {
  if (c.blk) {
    const long long colq = c.nz, rowq = (long long)GRID_BX * c.ny * colq;  // queries per column / per row of blocks
    const int bx = (int)(i / rowq);
    long long r = i - bx * rowq;
    const int wx = min(GRID_BX, c.nxs - GRID_BX * bx);
    const long long blkq = (long long)wx * GRID_BY * colq;
    const int by = (int)(r / blkq);
    r -= by * blkq;
    const int wy = min(GRID_BY, c.ny - GRID_BY * by);
    const int rem = (int)r;  // < 16 * 16 * nz
    const int lx = rem / (wy * c.nz);
    const int r3 = rem - lx * (wy * c.nz);
    const int ly = r3 / c.nz;
    iz = r3 - ly * c.nz;
    ix = c.x0 + GRID_BX * bx + lx;
    iy = GRID_BY * by + ly;
    return ((long long)(ix - c.x0) * c.ny + iy) * c.nz + iz;
  }
  const long long g = c.first + i;
  iz = (int)(g % c.nz);
  const long long t = g / c.nz;
  iy = (int)(t % c.ny);
  ix = (int)(t / c.ny);
  return i;
}
// where the value of the i-th query of the launch goes in the caller's output
__device__ __forceinline__ long long out_index(const QueryCtx& c, long long i) {
  if (c.qry || !c.blk) return i;
  int ix, iy, iz;
  return grid_point(c, i, ix, iy, iz);
}

// Query i -> model-space point (after the test-mode y,z flip or the train-mode rotation)
// and the clamped grid_sample coordinates (reference models.py:53-60, 28-36).
__device__ __forceinline__ int load_query(const QueryCtx& c, long long i, float& x, float& y, float& z, float& gu,
                                          float& gv) {
  const int b = c.per_img > 0 ? (int)(i / c.per_img) : 0;  // image of the batch this query belongs to
  if (c.qry) {
    x = c.qry[3 * i + 0];
    y = c.qry[3 * i + 1];
    z = c.qry[3 * i + 2];
  } else {
    int ix, iy, iz;
    grid_point(c, i, ix, iy, iz);
    x = c.px[ix];
    y = c.py[iy];
    z = c.pz[iz];
  }
  if (c.rot) {
    const float* R = c.rot + 9 * b;  // qry_rot = qry (1x3) . R (3x3)
    float rx = __fmaf_rn(z, R[6], __fmaf_rn(y, R[3], __fmul_rn(x, R[0])));
    float ry = __fmaf_rn(z, R[7], __fmaf_rn(y, R[4], __fmul_rn(x, R[1])));
    float rz = __fmaf_rn(z, R[8], __fmaf_rn(y, R[5], __fmul_rn(x, R[2])));
    x = rx; y = ry; z = rz;
  } else if (!c.preflipped) {
    y = -y;
    z = -z;
  }
  const float* T = c.T + 12 * b;
  float pu = x * T[0] + y * T[3] + z * T[6] + T[9];
  float pv = x * T[1] + y * T[4] + z * T[7] + T[10];
  float pw = x * T[2] + y * T[5] + z * T[8] + T[11];
  gu = fminf(fmaxf(2.f * (pu / pw - 0.5f), -1.f), 1.f);
  gv = fminf(fmaxf(2.f * (pv / pw - 0.5f), -1.f), 1.f);
  return b;
}

// grid_sample(bilinear, zeros, align_corners=True) tap set for one plane resolution R
// (reference models.py:45).  Coordinates are clamped to [-1,1] so every tap with a
// non-zero weight is in bounds; out-of-range neighbours get weight 0 and a clamped offset.
struct Taps {
  int o00, o01, o10, o11;  // pixel offsets (y*R + x)
  float w00, w01, w10, w11;
};
__device__ __forceinline__ Taps make_taps(float gu, float gv, int R) {
  float ix = ((gu + 1.f) * 0.5f) * (float)(R - 1);
  float iy = ((gv + 1.f) * 0.5f) * (float)(R - 1);
  float fx = floorf(ix), fy = floorf(iy);
  int x0 = (int)fx, y0 = (int)fy;
  float ax = ix - fx, ay = iy - fy;  // weight of the +1 neighbours
  float bx = (fx + 1.f) - ix, by = (fy + 1.f) - iy;  // ATen: (ix_se - ix), (iy_se - iy)
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
