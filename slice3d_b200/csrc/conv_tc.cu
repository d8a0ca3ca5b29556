// 3x3 / 1x1 convolution of the plane encoder as a TMA-staged tcgen05 implicit GEMM (sm_100a).
//
//   out[pixel][co] = epilogue( sum_{tap, ci} in[pixel + tap offset][ci] * W[co][tap][ci] )
//
// reference: every Conv2d(3x3, padding=1) of reg_slices/src/unet_custom.py:15-19 (VGG16-BN trunk) and of
// reg_slices/src/unet_parts.py:8-25 (DoubleConv of the four Up stages), eval mode, BatchNorm folded.
//
// Activations live in HBM as "split" NHWC tensors: two fp16 arrays hi = fp16(x), lo = fp16(x - hi)
// (x = hi + lo + O(2^-22 |x|); bf16 pairs only give 2^-17, measured as 2.5e-4 max-abs on the planes after the ~20
// convolutions between the image and the last plane, against the 1e-4 bar), written by the producing kernel's epilogue.  One CTA computes a tile of 128 output
// pixels (a BW x BH patch of one image) x BN output channels:
//
//   producer warp   per k-block (one filter tap x 64 input channels): two 4-D tiled TMA loads (hi, lo) of the
//                   patch shifted by the tap offset -- out-of-bounds rows/columns are zero-filled by the TMA unit,
//                   which IS the convolution's zero padding -- landing as [128 pixels][64 ch] K-major tiles in the
//                   128-byte-swizzle layout UMMA reads; plus one bulk copy of the pre-swizzled weight tile
//                   ([BN][64] hi | lo).  NST-stage ring, mbarrier complete_tx.
//   MMA warp        per k-block 4 k-steps x 3 passes (lo.hi + hi.lo + hi.hi) into the next of the 512 / BN TMEM accumulators;
//                   tcgen05.commit frees the stage and hands the accumulator to the epilogue warps.
//   8 epilogue warps  drain every k-block's accumulator into fp32 registers (round-to-nearest running sum, see the
//                   note at the kernel), then (+ slice-independent addend) * scale + shift, ReLU -> fp32 NHWC
//                   and / or split fp16 NHWC for the next convolution.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace s3d {

namespace {

using namespace ptx;

constexpr int CT_EPI_WARPS = 8;
constexpr int CT_THREADS = 64 + 32 * CT_EPI_WARPS;
constexpr uint32_t CT_A_PART = 16384;  // [128 pixels][64 ch] fp16
constexpr int CT_MAX_KSPLIT = 4;       // split-K factor: the last item of a tile reads every partial tile through ONE SM

struct ConvTcParams {
  int H, W, NI;
  int bw_log2, BW, BH, tiles_x, tiles_y;
  int ncb;   // input channel blocks of 64
  int taps;  // 9 (3x3, pad 1) or 1
  int nkb;   // taps * ncb
  int n_tiles;  // output-channel tiles of BN
  const uint8_t* wimg;
  int Cout;  // real output channels (multiple of 32)
  const float* scale;
  const float* shift;
  const float* add;  // optional fp32 [NI / add_div][H][W][Cout], added before scale/shift
  int add_div;
  int relu;
  int shuffle;      // ConvTranspose2d(k=2, s=2) as a 1x1 "convolution" with N = 4*C (n = (dy*2+dx)*C + co): column n of
                    // pixel (y, x) goes to channel co of pixel (2y+dy, 2x+dx) of a 2H x 2W image; shift is indexed by co
  int Cq;           // shuffle: C
  float acc_scale;  // 2^-e: the weights are packed as w * 2^e
  float* out_f32;  // fp32 NHWC, row pitch ldf (or null)
  int ldf;
  __half* out_hi;  // split NHWC, row pitch lds >= Cout; channels [Cout, lds) are written as zero (or null)
  __half* out_lo;
  int lds;
  // split-K (deep contractions over few pixels: the 16^2 / 32^2 trunk layers fill 8-32 of the 148 SMs otherwise): work
  // item = (tile, split) with the split fastest; every item writes its fp32 partial tile to `part`, the item that arrives
  // LAST at the tile's counter sums the partials in split order (so the result does not depend on who was last) and runs
  // the epilogue.  The counters are zero before the launch and are left zero.
  int ksplit;
  float* part;    // [tiles][ksplit][128 rows][BN]
  unsigned* cnt;  // [tiles]
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

template <int BN, int NST>
struct CtSmem {
  static constexpr uint32_t B_PART = BN * 128;                    // [BN][64] fp16
  static constexpr uint32_t STAGE = 2 * CT_A_PART + 2 * B_PART;   // A hi | A lo | B hi | B lo
  static constexpr uint32_t OFF_BAR = NST * STAGE;
  static constexpr int NACC = 512 / BN;                           // accumulator buffers: all of tensor memory
  static constexpr uint32_t OFF_TMEMPTR = OFF_BAR + 8 * (2 * NST + 2 * NACC);
  static constexpr uint32_t BYTES = OFF_TMEMPTR + 16 + 1024;      // + alignment slack
};

// Accumulation: the tensor core adds every MMA's products into the fp32 TMEM accumulator with truncation, a
// bias of ~2^-24 of the accumulator per instruction; over the 864 instructions of a K = 4608 contraction that
// was measured as ~1e-4 relative error after the 13-convolution trunk.  So a TMEM accumulator only ever holds ONE
// k-block (12 instructions); the epilogue warps drain it and keep the running sum in registers with round-to-nearest
// fp32 adds.  (Measured in round 2: chaining 2 / 3 k-blocks per drain costs 1.0e-4 / 1.6e-4 on the feature planes instead
// of 4.7e-5, for 4-5 % of the encoder's time.)  The accumulators rotate through ALL of tensor memory (512 / BN buffers):
// with two, the hand-over "commit -> 8 epilogue warps drain -> 8 arrives -> MMA warp" (~1 kcycle round trip, see
// tools/sync_cost.cu) sat between every second k-block and the next; with 4-8 in flight it is hidden (encoder 2.20 -> 1.97 ms).
// What is left is the MMA warp's own issue time: a Cout = 64 layer issues N = 64 MMAs at ~55 cycles each (their floor is
// 32) plus ~350 cycles of waits / commits per k-block -- the two 64 -> 64 convolutions of the last Up stage run 10.3 kcycles
// per 128-pixel tile against 3.5 of tensor-pipe time.  L2 -> shared-memory traffic is NOT the limit: a variant that kept
// the nine weight tiles resident and loaded three 130-pixel halo rows per tile (100 KB instead of 432 KB per tile; tap
// operands = descriptors whose start address is shifted by dx x 128 bytes, which works with the descriptor's base-offset
// field left at ZERO -- the swizzle follows the absolute shared-memory address; setting the field as the PTX text
// suggests gives wrong results) was bit-identical and exactly as fast (2.11 vs 2.12 ms), so it was not kept.
template <int BN, int NST, bool SPLIT>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const ConvTcParams p) {
  using L = CtSmem<BN, NST>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto full = [&](int s) { return sbase + L::OFF_BAR + 8u * s; };
  auto empty = [&](int s) { return sbase + L::OFF_BAR + 8u * (NST + s); };
  constexpr uint32_t NACC = L::NACC;
  auto acc_full = [&](int b) { return sbase + L::OFF_BAR + 8u * (2 * NST + b); };
  auto acc_empty = [&](int b) { return sbase + L::OFF_BAR + 8u * (2 * NST + NACC + b); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < (int)NACC; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), CT_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(sbase + L::OFF_TMEMPTR, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + L::OFF_TMEMPTR);

  // Persistent CTAs: tile t = (pixel tile, n tile) with the n tile fastest (CTAs that run together share A tiles in
  // L2).  The stage ring and the two accumulators keep rolling across tiles: while the epilogue warps finish a
  // tile (affine, stores) the producer and the MMA warp are already in the next one.
  const int tiles_img = p.tiles_x * p.tiles_y;
  const int n_tiles = p.n_tiles;
  const int ksplit = SPLIT ? p.ksplit : 1;  // (compile-time 1 in the unsplit instantiation: its code is the round-1 kernel's)
  const int total_tiles = tiles_img * p.NI * n_tiles * ksplit;  // work items
  auto kb_range = [&](int item, int& kb0, int& kb1) {
    const int sp = item % ksplit;
    kb0 = (int)((long long)sp * p.nkb / ksplit);
    kb1 = (int)((long long)(sp + 1) * p.nkb / ksplit);
  };
  auto decode = [&](int item, int& img, int& x0, int& y0, int& n_tile) {
    const int t = item / ksplit;
    n_tile = t % n_tiles;
    const int mt = t / n_tiles;
    img = mt / tiles_img;
    const int trem = mt - img * tiles_img;
    const int tyi = trem / p.tiles_x, txi = trem - tyi * p.tiles_x;
    x0 = txi * p.BW;
    y0 = tyi * p.BH;
  };

  if (warp == 0) {
    // ===================================================================== producer
    uint32_t g = 0;  // k-blocks issued so far (ring position)
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int img, x0, y0, n_tile;
      decode(t, img, x0, y0, n_tile);
      const uint8_t* wsrc = p.wimg + (size_t)n_tile * p.nkb * (2 * L::B_PART);
      int kb0, kb1;
      kb_range(t, kb0, kb1);
#pragma unroll 1
      for (int kb = kb0; kb < kb1; ++kb, ++g) {
        const uint32_t s = g % NST, it = g / NST;
        mbar_wait(empty(s), (it & 1) ^ 1u);  // "empty"-type: the first pass over the ring does not block
        if (lane == 0) {
          const int tap = kb / p.ncb, cb = kb - tap * p.ncb;
          int dy = 0, dx = 0;
          if (p.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap - (tap / 3) * 3 - 1;
          }
          const uint32_t st = sbase + s * L::STAGE;
          mbar_arrive_expect_tx(full(s), L::STAGE);
          tma_load_4d(st, &tm_hi, cb * 64, x0 + dx, y0 + dy, img, full(s));
          tma_load_4d(st + CT_A_PART, &tm_lo, cb * 64, x0 + dx, y0 + dy, img, full(s));
          bulk_g2s(st + 2 * CT_A_PART, wsrc + (size_t)kb * (2 * L::B_PART), 2 * L::B_PART, full(s));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // (Round 2 experiment, not kept: TWO issuer warps taking the k-blocks alternately -- every k-block has its own
    // accumulator, so they need no order -- made the big layers 4-7 % SLOWER: 243 vs 230 us for the 64 -> 64 layers.)
    constexpr uint32_t IDESC = make_idesc_f16(BN, 128);
    uint32_t g = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int kb0, kb1;
      kb_range(t, kb0, kb1);
#pragma unroll 1
      for (int kb = kb0; kb < kb1; ++kb, ++g) {
        const uint32_t s = g % NST, it = g / NST;
        const uint32_t buf = g % NACC;
        mbar_wait(acc_empty(buf), ((g / NACC) & 1) ^ 1u);  // the epilogue has drained this accumulator
        mbar_wait(full(s), it & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = sbase + s * L::STAGE;
          const uint32_t d = tmem + buf * BN;
          const uint64_t a_hi = make_desc_sw128(st), a_lo = make_desc_sw128(st + CT_A_PART);
          const uint64_t b_hi = make_desc_sw128(st + 2 * CT_A_PART), b_lo = make_desc_sw128(st + 2 * CT_A_PART + L::B_PART);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t o = (uint64_t)((ks * 32) >> 4);
            // correction terms first: they are ~2^-11 of the leading term
            umma_f16kind(d, a_lo + o, b_hi + o, IDESC, ks == 0 ? 0u : 1u);
            umma_f16kind(d, a_hi + o, b_lo + o, IDESC, 1u);
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t o = (uint64_t)((ks * 32) >> 4);
            umma_f16kind(d, a_hi + o, b_hi + o, IDESC, 1u);
          }
          umma_commit(empty(s));
          umma_commit(acc_full(buf));
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    constexpr int NC = BN / 2;  // columns per thread: the two warps of a TMEM lane quarter take half of the tile each
    const int q = warp & 3;     // TMEM lane quarter this warp may access
    const int ch = (warp - 2) >> 2;
    const int row = 32 * q + lane;
    const int bx = row & (p.BW - 1), by = row >> p.bw_log2;
    const uint32_t tsrc = tmem + (static_cast<uint32_t>(32 * q) << 16) + ch * NC;
    uint32_t g = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    int img, x0, y0, n_tile;
    decode(t, img, x0, y0, n_tile);
    const int x = x0 + bx, y = y0 + by;
    const bool valid = (x < p.W) && (y < p.H);
    const size_t pix = ((size_t)img * p.H + (valid ? y : 0)) * p.W + (valid ? x : 0);
    const float* addp = p.add ? p.add + (((size_t)(img / p.add_div) * p.H + (valid ? y : 0)) * p.W + (valid ? x : 0)) * p.Cout : nullptr;
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.f;
    int kb0, kb1;
    kb_range(t, kb0, kb1);
#pragma unroll 1
    for (int kb = kb0; kb < kb1; ++kb, ++g) {
      const uint32_t buf = g % NACC;
      mbar_wait(acc_full(buf), (g / NACC) & 1);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < NC / 32; ++j) {
        float v[32];
        tmem_ld32(tsrc + buf * BN + 32 * j, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[32 * j + c] += v[c];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
    }
    bool finish = true;
    if (SPLIT) {
      const int tile = t / ksplit, sp = t - tile * ksplit;
      // partial tile of item (tile, sp): [column half ch][16-byte chunk c][row] -- a warp's 32 rows are contiguous
      auto part_ptr = [&](int sp_) {
        return reinterpret_cast<float4*>(p.part) + (((size_t)tile * ksplit + sp_) * 2 + ch) * (size_t)(NC / 4) * 128 + row;
      };
      float4* mine = part_ptr(sp);
#pragma unroll
      for (int c = 0; c < NC / 4; ++c) __stcg(mine + c * 128, make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]));
      __threadfence();
      named_bar_sync(1, 32 * CT_EPI_WARPS);  // all partial rows of this item are written
      uint32_t* flag = reinterpret_cast<uint32_t*>(sgen + L::OFF_TMEMPTR + 8);
      if (warp == 2 && lane == 0) {
        const unsigned old = atomicAdd(p.cnt + tile, 1u);
        const bool last = old == (unsigned)(ksplit - 1);
        if (last) p.cnt[tile] = 0;  // every item of the tile has arrived: leave the counter ready for the next launch
        __threadfence();
        *flag = last ? 1u : 0u;
      }
      named_bar_sync(1, 32 * CT_EPI_WARPS);
      finish = *reinterpret_cast<volatile uint32_t*>(flag) != 0u;
      named_bar_sync(1, 32 * CT_EPI_WARPS);  // (the flag is rewritten by the next item)
      if (finish) {
        // sum in split order whoever arrived last (ksplit <= CT_MAX_KSPLIT; absent splits add an exact zero); four
        // chunks x four splits of loads in flight per thread
#pragma unroll
        for (int c0 = 0; c0 < NC / 4; c0 += 4) {
          float4 v4[CT_MAX_KSPLIT][4];
#pragma unroll
          for (int s2 = 0; s2 < CT_MAX_KSPLIT; ++s2) {
            const float4* src = part_ptr(s2 < ksplit ? s2 : 0);
#pragma unroll
            for (int j = 0; j < 4; ++j) v4[s2][j] = s2 < ksplit ? __ldcg(src + (c0 + j) * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 a = v4[0][j];
#pragma unroll
            for (int s2 = 1; s2 < CT_MAX_KSPLIT; ++s2) {
              a.x += v4[s2][j].x; a.y += v4[s2][j].y; a.z += v4[s2][j].z; a.w += v4[s2][j].w;
            }
            acc[4 * (c0 + j)] = a.x; acc[4 * (c0 + j) + 1] = a.y; acc[4 * (c0 + j) + 2] = a.z; acc[4 * (c0 + j) + 3] = a.w;
          }
        }
      }
    }
    if (valid && finish) {
#pragma unroll
      for (int j = 0; j < NC / 32; ++j) {
        float* v = acc + 32 * j;
        const int n0 = n_tile * BN + ch * NC + 32 * j;
        int nb = n0;         // index into scale / shift / add and output channel
        size_t opix = pix;   // output pixel
        if (p.shuffle) {
          const int qd = n0 / p.Cq;
          nb = n0 - qd * p.Cq;
          opix = ((size_t)img * (2 * p.H) + (2 * y + (qd >> 1))) * (2 * p.W) + (2 * x + (qd & 1));
        }
        if (n0 < p.Cout) {
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            float4 t = make_float4(v[c] * p.acc_scale, v[c + 1] * p.acc_scale, v[c + 2] * p.acc_scale, v[c + 3] * p.acc_scale);
            if (addp) {
              const float4 a4 = __ldg(reinterpret_cast<const float4*>(addp + nb + c));
              t.x += a4.x; t.y += a4.y; t.z += a4.z; t.w += a4.w;
            }
            if (p.scale) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + nb + c));
              t.x *= s4.x; t.y *= s4.y; t.z *= s4.z; t.w *= s4.w;
            }
            if (p.shift) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.shift + nb + c));
              t.x += s4.x; t.y += s4.y; t.z += s4.z; t.w += s4.w;
            }
            if (p.relu) {
              t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f);
            }
            v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
          }
          if (p.out_f32) {
            float4* dst = reinterpret_cast<float4*>(p.out_f32 + opix * p.ldf + nb);
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          }
          if (p.out_hi) {
            uint4* dh = reinterpret_cast<uint4*>(p.out_hi + opix * p.lds + nb);
            uint4* dl = reinterpret_cast<uint4*>(p.out_lo + opix * p.lds + nb);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t h[4], l[4];
              split8_h(v + 8 * c, h, l);
              dh[c] = make_uint4(h[0], h[1], h[2], h[3]);
              dl[c] = make_uint4(l[0], l[1], l[2], l[3]);
            }
          }
        } else if (p.out_hi && !p.shuffle && n0 < p.lds) {  // zero the channel padding the next convolution's TMA will read
          uint4* dh = reinterpret_cast<uint4*>(p.out_hi + pix * p.lds + n0);
          uint4* dl = reinterpret_cast<uint4*>(p.out_lo + pix * p.lds + n0);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            dh[c] = make_uint4(0, 0, 0, 0);
            dl[c] = make_uint4(0, 0, 0, 0);
          }
        }
      }
    }
    }  // tile loop
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// fp16 NHWC [NI][H][W][C] -> 4-D tensor map with a (64 ch, BW, BH, 1) box, 128-byte swizzle, zero OOB fill.
int make_tmap(CUtensorMap* tm, const __half* base, int NI, int H, int W, int C, int BW, int BH) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("conv_tc: cuTensorMapEncodeTiled is not available from this driver");
    return S3D_ERR_CUDA;
  }
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NI};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)BW, (cuuint32_t)BH, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_tc: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return S3D_ERR_CUDA;
  }
  return S3D_OK;
}

// Split-K factor of one launch.  Cost model in units of one k-block (~0.75 us): waves x k-blocks per item, + 3 per wave
// for the pipeline fill / drain of an item, + the last item's reduction (2 + ks: it pulls ks x 64 KB through one SM --
// a first version that split 18 ways made the 16^2 layers SLOWER, 60 vs 54 us).  Split only when the model promises
// < 0.7 x the unsplit cost and the scratch holds the partial tiles.
int choose_ksplit(long long tiles, int nkb, int sms, size_t tile_bytes, const SplitK* sk) {
#ifdef CT_NO_SPLITK  // (A / B builds)
  return 1;
#endif
  if (!sk || !sk->part || !sk->cnt || tiles > sk->n_cnt || tiles >= sms || nkb < 8) return 1;
  auto cost = [&](int ks) {
    const long long waves = (tiles * ks + sms - 1) / sms;
    const int per = (nkb + ks - 1) / ks;
    return (double)waves * (per + 3) + (ks > 1 ? 2.0 + ks : 0.0);
  };
  int best = 1;
  double best_cost = cost(1);
  const double base = best_cost;
  for (int ks = 2; ks <= nkb / 2 && ks <= CT_MAX_KSPLIT; ++ks) {
    if ((size_t)tiles * ks * tile_bytes > sk->part_bytes) break;
    const double c = cost(ks);
    if (c < best_cost) {
      best_cost = c;
      best = ks;
    }
  }
  return best_cost < 0.7 * base ? best : 1;
}

inline uint16_t f16_bits_h(float x) { return __half_as_ushort(__float2half_rn(x)); }
inline float f16_val_h(uint16_t b) { return __half2float(__ushort_as_half(b)); }

}  // namespace

// Weight image of one convolution: [n_tile][k-block = tap x 64-channel block][hi BN x 64 | lo BN x 64] bf16, every
// [BN][64] tile in the K-major 128-byte-swizzle layout (so a plain bulk copy lands a ready UMMA B operand).
// `src` is the fp32 GEMM matrix [(tap*src_cin + ci)][src_ld] of the SIMT path (device); input channels
// [ci0, ci0 + cin) of it are used and padded with zeros to a multiple of 64; output channels padded to BN.
int convtc_pack(s3d_model* m, const ConvW& cw, int src_cin, int ci0, int cin, ConvTC& out, cudaStream_t st) {
  const int taps = cw.ks * cw.ks, cout = cw.ncols;
  const int cinp = (cin + 63) / 64 * 64;
  const int bn = cout >= 128 ? 128 : 64;
  const int n_tiles = (cout + bn - 1) / bn;
  const int ncb = cinp / 64, nkb = taps * ncb;
  std::vector<float> h((size_t)cw.kpad * cout);
  S3D_CUDA(cudaMemcpyAsync(h.data(), cw.w, h.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
  S3D_CUDA(cudaStreamSynchronize(st));
  // scale the weights by a power of two so that max |w| lands in [128, 256): hi and lo are then normal fp16 numbers
  float wmax = 0.f;
  for (int tap = 0; tap < taps; ++tap)
    for (int ci = 0; ci < cin; ++ci)
      for (int co = 0; co < cout; ++co) wmax = std::max(wmax, std::fabs(h[((size_t)tap * src_cin + ci0 + ci) * cout + co]));
  int e = 0;
  if (wmax > 0.f && std::isfinite(wmax)) {
    int ex;
    std::frexp(wmax, &ex);  // wmax = f * 2^ex, f in [0.5, 1)
    e = 8 - ex;
  }
  const float wmul = std::ldexp(1.f, e);
  out.acc_scale = std::ldexp(1.f, -e);
  const size_t tile_bytes = (size_t)bn * 256;
  std::vector<uint8_t> img((size_t)n_tiles * nkb * tile_bytes, 0);
  for (int nt = 0; nt < n_tiles; ++nt)
    for (int kb = 0; kb < nkb; ++kb) {
      const int tap = kb / ncb, cb = kb % ncb;
      uint16_t* hi = reinterpret_cast<uint16_t*>(img.data() + ((size_t)nt * nkb + kb) * tile_bytes);
      uint16_t* lo = hi + (size_t)bn * 64;
      for (int n = 0; n < bn; ++n) {
        const int co = nt * bn + n;
        if (co >= cout) continue;
        for (int k = 0; k < 64; ++k) {
          const int ci = cb * 64 + k;
          if (ci >= cin) continue;
          const float w = h[((size_t)tap * src_cin + ci0 + ci) * cout + co] * wmul;
          const uint16_t hb = f16_bits_h(w);
          const uint16_t lb = f16_bits_h(w - f16_val_h(hb));
          const size_t off = (ptx::sw128_chunk_off(n, k >> 3) + (k & 7) * 2) / 2;
          hi[off] = hb;
          lo[off] = lb;
        }
      }
    }
  void* d = nullptr;
  S3D_CUDA(cudaMalloc(&d, img.size()));
  m->allocs.push_back(d);
  S3D_CUDA(cudaMemcpyAsync(d, img.data(), img.size(), cudaMemcpyHostToDevice, st));
  S3D_CUDA(cudaStreamSynchronize(st));
  out.wimg = static_cast<uint8_t*>(d);
  out.cinp = cinp;
  out.cout = cout;
  out.bn = bn;
  out.n_tiles = n_tiles;
  out.taps = taps;
  out.scale = cw.scale;
  out.shift = cw.shift;
  return S3D_OK;
}

// in: split NHWC [NI][H][W][w.cinp] (hi, lo).  Outputs: fp32 NHWC (pitch ldf) and / or split NHWC (pitch lds).
int conv_tc(const ConvTC& w, const __half* in_hi, const __half* in_lo, int NI, int H, int W, const float* add,
            int add_div, int relu, float* out_f32, int ldf, __half* out_hi, __half* out_lo, int lds,
            cudaStream_t st, int shuffle_c, const SplitK* sk) {
  if (NI <= 0) return S3D_OK;
  int BW = 1, lg = 0;
  while (BW * 2 <= W && BW < 128) {
    BW *= 2;
    ++lg;
  }
  const int BH = 128 / BW;
  ConvTcParams p{};
  p.H = H; p.W = W; p.NI = NI;
  p.BW = BW; p.bw_log2 = lg; p.BH = BH;
  p.tiles_x = (W + BW - 1) / BW;
  p.tiles_y = (H + BH - 1) / BH;
  p.ncb = w.cinp / 64;
  p.taps = w.taps;
  p.nkb = p.taps * p.ncb;
  p.n_tiles = w.n_tiles;
  p.wimg = w.wimg;
  p.Cout = w.cout;
  p.scale = w.scale; p.shift = w.shift;
  p.add = add; p.add_div = add_div > 0 ? add_div : 1;
  p.relu = relu;
  p.shuffle = shuffle_c > 0 ? 1 : 0;
  p.Cq = shuffle_c > 0 ? shuffle_c : 1;
  p.acc_scale = w.acc_scale;
  p.out_f32 = out_f32; p.ldf = ldf;
  p.out_hi = out_hi; p.out_lo = out_lo; p.lds = lds;
  CUtensorMap tm_hi, tm_lo;
  S3D_TRY(make_tmap(&tm_hi, in_hi, NI, H, W, w.cinp, BW, BH));
  S3D_TRY(make_tmap(&tm_lo, in_lo, NI, H, W, w.cinp, BW, BH));
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    S3D_CUDA(cudaGetDevice(&dev));
    S3D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long tiles = (long long)p.tiles_x * p.tiles_y * NI * w.n_tiles;
  p.ksplit = choose_ksplit(tiles, p.nkb, sms, (size_t)128 * w.bn * sizeof(float), sk);
  if (p.ksplit > 1) {
    p.part = sk->part;
    p.cnt = sk->cnt;
  }
  const long long total = tiles * p.ksplit;
  dim3 grid((unsigned)(total < sms ? total : sms));
  auto launch = [&](auto kern, size_t smem) -> int {
    S3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, CT_THREADS, smem, st>>>(tm_hi, tm_lo, p);
    return S3D_OK;
  };
  if (w.bn == 128) {
    if (p.ksplit > 1) S3D_TRY(launch(conv_tc_kernel<128, 3, true>, CtSmem<128, 3>::BYTES));
    else S3D_TRY(launch(conv_tc_kernel<128, 3, false>, CtSmem<128, 3>::BYTES));
  } else {
    if (p.ksplit > 1) S3D_TRY(launch(conv_tc_kernel<64, 4, true>, CtSmem<64, 4>::BYTES));
    else S3D_TRY(launch(conv_tc_kernel<64, 4, false>, CtSmem<64, 4>::BYTES));
  }
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

}  // namespace s3d
