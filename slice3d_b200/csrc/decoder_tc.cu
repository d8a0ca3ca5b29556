// Fused tensor-core decoder for sm_100a (S3D_PREC_FP16F8 / FP16X3 / BF16X3 / BF16).
//
// One persistent CTA per SM processes tiles of 9 queries = 117 token rows (+11 pad rows) = one
// 128-row UMMA tile, start to finish inside the SM:
//
//   token build   bilinear gather of the fc_s-projected planes (reference models.py:69-80)
//   3 x layer     QKV projection -> 13x13 attention per query and head -> out-proj + residual +
//                 LayerNorm -> FFN 128->2048 (ReLU) ->128 + residual + LayerNorm
//                 (nn.TransformerEncoderLayer, post-norm; models.py:18-19,82-83)
//   head          fc_out on token 0 (models.py:84), scaled by out_scale
//
// Every dense contraction runs on tcgen05.mma (M=128, bf16 operands, fp32 accumulators in TMEM).
// Activations are written by the compute warps straight into the canonical K-major/128B-swizzled
// shared-memory operand layout; weights are pre-swizzled "operand images" in global memory streamed
// as 16 KB parts through a 5-slot ring with bulk async copies (TMA engine) completing on mbarriers.
// The residual stream never leaves TMEM: the accumulator of out-proj / FFN2 is pre-loaded with
// x + bias, so the MMA result is already residual + bias + contraction and LayerNorm runs in place.
// The FFN's hidden chunks live in three 128-column TMEM buffers: linear1's accumulator, overwritten in place by
// the packed H operand of linear2 (see TM_D1 and the MMA issuer).
// Only token 0 of a query feeds fc_out, so in the last layer only those rows attend and the rest of
// that layer (out-proj, FFN, LayerNorms, head) runs as one "tail pass" per 14 tiles over 126 queries.
//
// Precision: BF16X3 splits both operands into bf16 hi + lo and issues hi*hi + lo*hi + hi*lo
// (3 passes, ~16 mantissa bits, max-abs error 3-4e-5 on sdf_pred); BF16 issues hi*hi only.
//
// Warp roles (640 threads): warps 0-15 = compute, thread = (tile row r = TMEM lane, column quarter g);
// warp 16 = weight producer; warp 17 = MMA issuer (warp-uniform control flow, one elected lane issues) / slot relay;
// warps 18-19 = gather warps (token build of the next tiles).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace s3d {

namespace {

using namespace ptx;

constexpr int TILE_Q = 9;    // queries per tile
constexpr int NTOK = 13;     // tokens per query (K = 12 slices + the query token)
constexpr int NSLOT = 5;       // weight ring: 5 slots of this CTA's half (16 KB) of one part
constexpr uint32_t SLOT_BYTES = 16384;
constexpr uint32_t RING_BYTES = NSLOT * SLOT_BYTES;
// Every weight unit is a [128 n][128 k] block (tcgen05.mma needs N >= 128 per instruction to run at the pipe's
// rate: N = 64 instructions were measured at 52-59 cycles against a 32-cycle floor, tools/mma_rate.cu).
constexpr int UNITS_PER_LAYER = 36;  // 3 (in_proj) + 1 (out_proj) + 16 (linear1 chunks) + 16 (linear2 chunks)
constexpr int UNIT_PART_BYTES = 32768;  // one precision part (hi or lo) of a unit: [2 k-blocks][128 n][64 k] bf16
constexpr int UNIT_STRIDE_BYTES = 2 * UNIT_PART_BYTES;  // hi then lo in global memory
constexpr int NCHUNK = 16;   // FFN hidden chunks of 128
constexpr int TAIL_SLOTS = 14;  // tiles whose token-0 rows are batched into one tail pass (14 x 9 = 126 rows)
constexpr int TAIL_UNIT0 = 2 * UNITS_PER_LAYER + 3;  // first unit of the tail pass: layer 2 out_proj
constexpr int NCW = 16;      // compute warps: warp w owns TMEM lanes 32*(w&3).. and column quarter w>>2
constexpr int NCT = NCW * 32;
#ifndef S3D_NGW
#define S3D_NGW 2
#endif
constexpr int NGW = S3D_NGW;       // gather warps (token build for the next tiles, fully asynchronous): warps 18-19
// 20 warps x 96 registers.  (24 warps with setmaxnreg re-budgeting -- 96 for the compute warps, 40/56 for the
// service warpgroups, 4 gather warps -- was measured: the MMA issuer spills at 40 registers and starves the pipe in
// bf16x3 mode; 8 % slower there, equal in bf16 mode.)
constexpr int NTHREADS = NCT + 64 + NGW * 32;  // + producer warp + MMA warp + gather warps

// per-layer fp32 vector block staged in shared memory (floats)
constexpr int V_BIN = 0, V_BONEXT = 384, V_LN1W = 512, V_LN1B = 640, V_B2 = 768, V_LN2W = 896, V_LN2B = 1024,
              V_B1 = 1152, VEC_FLOATS = 3200, V_SMEM_FLOATS = VEC_FLOATS;
constexpr int V_PART_B = 384;  // floats [0, 384) = b_in (attention phase); the rest is used from LayerNorm 1 on
// K/V staging for attention: fp32 [117 rows][ST_PITCH] in the H region (+ the last layer's [9][4][13] scores)
constexpr int ST_PITCH = 132;

// shared memory map (bytes from the 1024-aligned base)
constexpr uint32_t OFF_AX_HI = 0;                 // [2 k-blocks][128 rows][64] bf16 = 32 KB
constexpr uint32_t OFF_AX_LO = 32768;             // 32 KB
constexpr uint32_t OFF_H = 65536;                 // attention staging: K, then V, of all heads as fp32 [117][ST_PITCH] (+ last-layer scores)
constexpr uint32_t H_BUF_BYTES = 32768;
constexpr uint32_t OFF_RING = OFF_H + 2 * H_BUF_BYTES;  // 131072
constexpr uint32_t OFF_SC = OFF_H + TILE_Q * NTOK * ST_PITCH * 4;
static_assert(OFF_SC + TILE_Q * 4 * NTOK * 4 <= OFF_RING, "attention staging");
constexpr uint32_t OFF_VEC = OFF_RING + RING_BYTES;  // 212992
constexpr uint32_t OFF_RED = OFF_VEC + V_SMEM_FLOATS * 4;          // [128][4] x (mean, M2) fp32 LayerNorm partials
constexpr uint32_t OFF_BAR = OFF_RED + 5120;  // (last layer: [9][128] fp32 scaled queries of the token-0 rows)
constexpr uint32_t SMEM_BYTES = OFF_BAR + 512 + 1024;              // + alignment slack

// barrier indices (8 bytes each at OFF_BAR)
enum { B_AREADY = 0, B_DDONE, B_KDONE, B_QDONE, B_D1READY0, B_D1READY1, B_D1READY2, B_HREADY0, B_HREADY1, B_HREADY2,
       B_TOKFULL0, B_TOKFULL1, B_TOKEMPTY0, B_TOKEMPTY1, B_FULL0, B_EMPTY0 = B_FULL0 + NSLOT,
       B_COUNT = B_EMPTY0 + NSLOT };
static_assert(8 * B_COUNT + 8 <= 512, "barrier area");
constexpr uint32_t OFF_TMEMPTR = OFF_BAR + 8 * B_COUNT;

// TMEM columns
constexpr uint32_t TM_R = 0;      // residual / out-proj / FFN2 accumulator, 128 columns
constexpr uint32_t TM_S = 128;    // QKV accumulators (384 columns), or during the FFN:
constexpr uint32_t TM_D1 = 128;   //   128..511  three FFN chunk buffers of 128 columns: the accumulator D1 of linear1 (128 hidden
                                  //             units, fp32), overwritten IN PLACE by the H operand of linear2: the thread that
                                  //             owns columns 32g .. 32g+31 of a row writes the packed H values of the same hidden
                                  //             units back into them -- [hi 16 | lo 16] columns, fp16f8: [fp16 16 | e4m3 8 | e4m3 8]
constexpr int NDBUF = 3;
constexpr uint32_t TM_HT = 384;   //   (self-test kernels: one chunk buffer in the same in-place layout)

// Phase-cycle counters (clock64 deltas summed over CTAs), read through s3d_debug_profile().
enum { PF_TOKEN = 0, PF_VEC, PF_WAIT_QKV, PF_ATTN, PF_WAIT_OUT, PF_LN1, PF_FFN_WAIT_D1, PF_FFN_MATH, PF_FFN_WAIT_HFREE,
       PF_FFN_STORE, PF_WAIT_FFN, PF_LN2, PF_MMA_WAIT_A, PF_MMA_WAIT_FULL, PF_MMA_WAIT_H, PF_MMA_WAIT_D1FREE, PF_MMA_TOTAL,
       PF_PROD_WAIT_EMPTY, PF_PROD_TOTAL, PF_TILES, PF_COUNT };
__device__ unsigned long long g_prof[32];
#ifdef S3D_TRACE  // experiment builds only: event timeline of CTA 0's MMA issuer (role 0) and compute warp 0 (role 1) over
// FFN chunks 4..11 of one tile, staged in the idle upper KB of the LayerNorm scratch and printed at the end of the launch
__device__ uint32_t g_trace[256];
#define TR(role, on, ev, c)                                                                                   \
  if ((on) && lane == 0 && (c) >= 4 && (c) < 12 && tr_n < 128) {                                              \
    reinterpret_cast<uint32_t*>(sgen + OFF_RED + 4096)[(role) * 128 + tr_n++] =                               \
        ((uint32_t)(ev) << 26) | (((uint32_t)(c) & 15u) << 22) | ((uint32_t)clock() & 0x3fffffu);              \
  }
#define TR_FLUSH(role, on)                                                                                    \
  if ((on) && lane == 0) {                                                                                    \
    for (int i_ = 0; i_ < 128; ++i_)                                                                          \
      g_trace[(role) * 128 + i_] = i_ < tr_n ? reinterpret_cast<uint32_t*>(sgen + OFF_RED + 4096)[(role) * 128 + i_] : 0u; \
  }
#else
#define TR(role, on, ev, c)
#define TR_FLUSH(role, on)
#endif

struct TcParams {
  const uint8_t* wimg;  // [3 layers][36 units][hi 32 KB | lo 32 KB]
  const float* vecs;    // [3 layers][VEC_FLOATS]
  float* scratch;         // [grid][TAIL_SLOTS*9 rows][256] fp32: attention output | x + b_o of token-0 rows
  const float* planes;
  int S;
  QueryCtx q;
  long long n;
  int K;             // slices of the model (<= 12): token rows K + 1 .. 12 of every query are dead (zero tokens, masked keys)
  const int* n_dev;  // when set: the query count lives in device memory (n is then only the capacity)
  float out_scale;
  float* out;
  const float *fcp_wt, *fcp_b, *fcs_b, *fco_w, *fco_b, *b_o0;
  long long num_tiles;
  int dbg;  // S3D_TC_DBG bit 0: the producer skips the weight copies (timing experiments only; results are garbage)
};

// ---- operand writes ------------------------------------------------------------------------
// 8 consecutive k values of row r -> one 16-byte chunk of the hi tile (and of the lo tile).
// x = hi + lo + O(2^-17 |x|): hi = bf16_rn(x), lo = bf16_rn(x - hi), converted two at a time.
template <bool LO, bool F16>
__device__ __forceinline__ void split8x(const float* v, uint32_t* h, uint32_t* l) {
  if (F16) split8_hn(v, h, l);  // fp16 pairs (only used with LO): x = hi + lo + O(2^-22 |x|) for |x| < 65504
  else split8<LO>(v, h, l);
}
template <int NPASS, bool F16>
__device__ __forceinline__ void store_chunk(uint8_t* tile_hi, uint8_t* tile_lo, int r, int kc, const float* v) {
  uint32_t h[4], l[4];
  split8x<NPASS == 3, F16>(v, h, l);
  const uint32_t off = sw128_chunk_off(r, kc);
  *reinterpret_cast<uint4*>(tile_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  if (NPASS == 3) *reinterpret_cast<uint4*>(tile_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---- MMA issue -----------------------------------------------------------------------------
// One weight part (B) against NA activation operands (a0 [, a1]): D += a0.B [+ a1.B], KS k-steps of 16 each,
// unrolled (a rolled loop was measured to starve the pipe: the issuer shares its scheduler with busy warps),
// with 32-bit descriptor words so that the unrolled sequence stays small in registers.
//   *_KB: byte stride between 64-wide k-blocks of the A / B tiles.
template <int CG, int NA, int KS, uint32_t A_KB, uint32_t B_KB, uint32_t IDESC>
__device__ __forceinline__ void issue_part(uint32_t d_tmem, uint32_t a0, uint32_t a1, uint32_t b, bool fresh) {
  if (CG == 2) {
    const uint32_t bl = make_desc_lo(b);
#pragma unroll
    for (int pass = 0; pass < NA; ++pass) {
      const uint32_t al = make_desc_lo(pass == 1 ? a1 : a0);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
        umma_pair_lo(d_tmem, al + ((kb * A_KB + kin) >> 4), bl + ((kb * B_KB + kin) >> 4), IDESC,
                     (pass == 0 && ks == 0 && fresh) ? 0u : 1u);
      }
    }
  } else {
    const uint64_t ad0 = make_desc_sw128(a0), ad1 = make_desc_sw128(a1), bd = make_desc_sw128(b);
#pragma unroll
    for (int pass = 0; pass < NA; ++pass) {
      const uint64_t ad = (pass == 1) ? ad1 : ad0;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
        umma_bf16(d_tmem, ad + ((kb * A_KB + kin) >> 4), bd + ((kb * B_KB + kin) >> 4), IDESC,
                  (pass == 0 && ks == 0 && fresh) ? 0u : 1u);
      }
    }
  }
}

// Same with the A operand(s) in tensor memory (8 columns per k-step).  INPLACE: the operand sits in the columns of the
// accumulator it was computed from (see TM_D1): k-step ks (16 values) at column 32 (ks / 2) + 8 (ks % 2).
template <int CG, int NA, int KS, uint32_t B_KB, uint32_t IDESC, bool INPLACE = false>
__device__ __forceinline__ void issue_part_ts(uint32_t d_tmem, uint32_t a0_tmem, uint32_t a1_tmem, uint32_t b, bool fresh) {
  if (CG == 2) {
    const uint32_t bl = make_desc_lo(b);
#pragma unroll
    for (int pass = 0; pass < NA; ++pass) {
      const uint32_t at = (pass == 1) ? a1_tmem : a0_tmem;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
        const uint32_t acol = INPLACE ? 32u * (ks >> 1) + 8u * (ks & 1) : 8u * ks;
        umma_ts_pair_lo(d_tmem, at + acol, bl + ((kb * B_KB + kin) >> 4), IDESC, (pass == 0 && ks == 0 && fresh) ? 0u : 1u);
      }
    }
  } else {
    const uint64_t bd = make_desc_sw128(b);
#pragma unroll
    for (int pass = 0; pass < NA; ++pass) {
      const uint32_t at = (pass == 1) ? a1_tmem : a0_tmem;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
        const uint32_t acol = INPLACE ? 32u * (ks >> 1) + 8u * (ks & 1) : 8u * ks;
        umma_bf16_ts(d_tmem, at + acol, bd + ((kb * B_KB + kin) >> 4), IDESC, (pass == 0 && ks == 0 && fresh) ? 0u : 1u);
      }
    }
  }
}

// warp-level arrive: every lane has fenced its own writes; one lane arrives for the warp.
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// CG = 1: every CTA is on its own.  CG = 2: CTA pairs (cluster of 2 = one TPC): the leader's MMA warp issues
// cta_group::2 MMAs (M = 256: 128 rows = one tile in each CTA) and every CTA streams only ITS HALF of each weight
// part (N/2 rows of B), which halves the L2 -> shared-memory weight traffic and the shared-memory operand reads per SM
// and doubles the MMA time one ring slot covers.  Everything outside the MMA / producer warps is per CTA.
template <int NPASS, int CG, bool F16, bool F8>
__global__ void __launch_bounds__(NTHREADS, 1) decoder_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;  // (the dynamic window starts at the same offset in both CTAs of a pair)
  uint8_t* sgen = smem_raw + (sbase - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  static_assert(CG == 2, "the decoder runs as CTA pairs");
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  // tiles of this CTA: base + rank for base = CG * group, CG * (group + #groups), ...
  const long long tile_first = blockIdx.x;  // = CG * (blockIdx.x / CG) + rank
  const long long base_first = (long long)blockIdx.x - rank;
  // arrive on barrier i of the leader CTA (the MMA issuer's barriers)
  auto arrive_lead = [&](int i) {
    __syncwarp();
    if (lane == 0) {
      if (CG == 2 && !leader) mbar_arrive_cluster(mapa_u32(bar(i), 0));
      else mbar_arrive(bar(i));
    }
  };
  auto wait_lead = [&](int i, uint32_t parity) {  // leader-side wait on a barrier the peer arrives on too
    if (CG == 2) mbar_wait_cluster(bar(i), parity);
    else mbar_wait(bar(i), parity);
  };

  if (threadIdx.x == 0) {
    mbar_init(bar(B_AREADY), NCW * CG);
    mbar_init(bar(B_DDONE), 1);
    mbar_init(bar(B_KDONE), 1);
    mbar_init(bar(B_QDONE), 1);
    for (int b = 0; b < NDBUF; ++b) {
      mbar_init(bar(B_D1READY0 + b), 1);
      mbar_init(bar(B_HREADY0 + b), NCW * CG);
    }
    mbar_init(bar(B_TOKFULL0), NGW);
    mbar_init(bar(B_TOKFULL1), NGW);
    mbar_init(bar(B_TOKEMPTY0), NCW);
    mbar_init(bar(B_TOKEMPTY1), NCW);
    for (int s = 0; s < NSLOT; ++s) {
      // leader: the producer's arrive (+ its bytes) and the peer's relay ("my half has landed too") complete ONE barrier
      mbar_init(bar(B_FULL0 + s), (CG == 2 && leader) ? 2 : 1);
      mbar_init(bar(B_EMPTY0 + s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    if (CG == 2) tmem_alloc_pair(sbase + OFF_TMEMPTR, 512);
    else tmem_alloc(sbase + OFF_TMEMPTR, 512);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + OFF_TMEMPTR);
  // number of queries: a launch parameter, or read from device memory (device-resident MISE rounds: the count is
  // produced by the compaction kernel that precedes this launch in the stream; the grid is then launched full)
  const long long n_q = p.n_dev ? (long long)__ldg(p.n_dev) : p.n;
  const long long n_tiles = (n_q + TILE_Q - 1) / TILE_Q;

  constexpr int NPART = (NPASS == 3) ? 2 : 1;  // ring parts per weight unit (hi [, lo])

  // Single-pass mode only (there the gather warps bind: 36 of 149 kcycles per tile were token waits): register
  // re-budgeting inside the CTA's launch allocation (640 threads x 96) -- the four compute warpgroups give up 8 registers
  // per thread, the service warpgroup (producer, MMA issuer, two gather warps) takes 128, enough for the gather warps
  // to keep all 20 tap loads of a step in flight (one L2 round trip per step instead of two): 149 -> 130 kcycles.
  // In the three-pass mode the compute warps need their 96 registers more (measured 189.7 vs 183.3 kcycles).
  constexpr bool REGSPLIT = (NPASS == 1);
  if (REGSPLIT) {
    if (warp >= NCW) reg_alloc<128>();
    else reg_dealloc<88>();
  }
  if (warp >= NCW + 2) {
    // ===================================================================== gather warps
    // They run up to two tiles ahead of the compute warps (double-buffered token scratch), so the L2 latency
    // of the bilinear taps is off the critical path.
    // Token gather, decoupled from row ownership.  The 108 (slice, query, 32-channel block) tasks of a tile are
    // dealt four to a warp-step (one per quarter-warp, 16 bytes per lane; the four quarter-warps of a step take
    // the same slice and channel block of four consecutive queries, whose texels coincide or neighbour each
    // other, so their requests coalesce), 56 steps for each of the two gather warps.  Each step issues the 12 tap loads of plane scales
    // 0-2 together, then the 8 of scales 3-4 (two L2 round trips per step instead of five: with ~226 KB of shared
    // memory in use there is no L1 to speak of and the gather is bound by L2 latency x loads in flight).  The
    // 16-byte token pieces go to this CTA's global token scratch [128 rows][128] (L2-resident); the row owners
    // pick them up after a CTA barrier.
    float qgu = 0.f, qgv = 0.f;  // lane l < 9: grid_sample coordinates of query l of the current tile
    int qimg = 0;                // ... and the image of the batch it belongs to
    auto gather_step = [&](long long gt, int idx, float* tokbuf) {  // idx 0..111
      const int qtr = lane >> 3, l8 = lane & 7;
      int q, k, cb;
      if (idx < 96) {  // steps 0..5: queries 0..7, four per step
        cb = idx & 3;
        const int grp = idx >> 2;  // 0..23
        k = grp >> 1;
        q = 4 * (grp & 1) + qtr;
      } else {  // step 6: the 48 tasks of query 8, four slices per step
        const int t = (idx - 96) * 4 + qtr;  // 0..63, 48 real
        cb = t & 3;
        k = t >> 2;
        q = 8;
      }
      const long long gq = gt * TILE_Q + q;
      const float gu = __shfl_sync(0xffffffffu, qgu, q), gv = __shfl_sync(0xffffffffu, qgv, q);
      const int img = __shfl_sync(0xffffffffu, qimg, q);
      if (k >= p.K || gq >= n_q) return;
      const int ch = 32 * cb + l8 * 4;
      float4 acc = __ldg(reinterpret_cast<const float4*>(p.fcs_b + ch));
      const int R0 = plane_res(p.S, 0);
      const float* P = p.planes + (size_t)img * p.q.plane_stride + (size_t)k * R0 * R0 * 128 + ch;
      auto fold = [&](const float4* v, const Taps& t) {
        acc.x += v[0].x * t.w00 + v[1].x * t.w01 + v[2].x * t.w10 + v[3].x * t.w11;
        acc.y += v[0].y * t.w00 + v[1].y * t.w01 + v[2].y * t.w10 + v[3].y * t.w11;
        acc.z += v[0].z * t.w00 + v[1].z * t.w01 + v[2].z * t.w10 + v[3].z * t.w11;
        acc.w += v[0].w * t.w00 + v[1].w * t.w01 + v[2].w * t.w10 + v[3].w * t.w11;
      };
      auto issue = [&](float4* v, const Taps& t, const float* base) {
        v[0] = __ldg(reinterpret_cast<const float4*>(base + (size_t)t.o00 * 128));
        v[1] = __ldg(reinterpret_cast<const float4*>(base + (size_t)t.o01 * 128));
        v[2] = __ldg(reinterpret_cast<const float4*>(base + (size_t)t.o10 * 128));
        v[3] = __ldg(reinterpret_cast<const float4*>(base + (size_t)t.o11 * 128));
      };
      // plane of scale s for slice k starts at P + (K * sum_{i<s} R_i^2 + k * (R_s^2 - R_0^2)) * 128
      const size_t r2 = (size_t)R0 * R0;
      if (REGSPLIT) {  // all 20 tap loads in flight: one L2 round trip per step
        float4 v0[4], v1[4], v2[4], v3[4], v4[4];
        const Taps t0 = make_taps(gu, gv, R0), t1 = make_taps(gu, gv, 2 * R0), t2 = make_taps(gu, gv, 4 * R0);
        const Taps t3 = make_taps(gu, gv, 8 * R0), t4 = make_taps(gu, gv, 16 * R0);
        issue(v0, t0, P);
        issue(v1, t1, P + ((size_t)p.K * r2 + (size_t)k * 3 * r2) * 128);
        issue(v2, t2, P + ((size_t)p.K * 5 * r2 + (size_t)k * 15 * r2) * 128);
        issue(v3, t3, P + ((size_t)p.K * 21 * r2 + (size_t)k * 63 * r2) * 128);
        issue(v4, t4, P + ((size_t)p.K * 85 * r2 + (size_t)k * 255 * r2) * 128);
        fold(v0, t0);
        fold(v1, t1);
        fold(v2, t2);
        fold(v3, t3);
        fold(v4, t4);
      } else {
      {  // 12 + 8 loads in flight per lane: two L2 round trips per step
        float4 v0[4], v1[4], v2[4];
        const Taps t0 = make_taps(gu, gv, R0), t1 = make_taps(gu, gv, 2 * R0), t2 = make_taps(gu, gv, 4 * R0);
        issue(v0, t0, P);
        issue(v1, t1, P + ((size_t)p.K * r2 + (size_t)k * 3 * r2) * 128);
        issue(v2, t2, P + ((size_t)p.K * 5 * r2 + (size_t)k * 15 * r2) * 128);
        fold(v0, t0);
        fold(v1, t1);
        fold(v2, t2);
      }
      {
        float4 v3[4], v4[4];
        const Taps t3 = make_taps(gu, gv, 8 * R0), t4 = make_taps(gu, gv, 16 * R0);
        issue(v3, t3, P + ((size_t)p.K * 21 * r2 + (size_t)k * 63 * r2) * 128);
        issue(v4, t4, P + ((size_t)p.K * 85 * r2 + (size_t)k * 255 * r2) * 128);
        fold(v3, t3);
        fold(v4, t4);
      }
      }
      __stcg(reinterpret_cast<float4*>(tokbuf + (size_t)(NTOK * q + 1 + k) * 128 + ch), acc);
    };

    if (p.q.tok_slice == nullptr) {  // (ready tokens -- Slices3DGTModel, gt.cu -- need no gather)
      float* const tokbase = p.scratch + (size_t)gridDim.x * (TAIL_SLOTS * TILE_Q) * 256 + (size_t)blockIdx.x * (2 * 128 * 128);
      uint32_t ph_te = 3u;  // "empty"-type: the first wait on each buffer passes
      int it = 0;
      for (long long tile = tile_first; tile - rank < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(bar(B_TOKEMPTY0 + buf), (ph_te >> buf) & 1u);
        ph_te ^= 1u << buf;
        float* const tokbuf = tokbase + (size_t)buf * (128 * 128);
        {  // the tile's 9 queries: one per lane (the dependent loads of load_query happen once per tile, not per step)
          const int ql = lane < TILE_Q ? lane : TILE_Q - 1;
          const long long gq = tile * TILE_Q + ql;
          float qx = 0.f, qy = 0.f, qz = 0.f;
          if (gq < n_q) qimg = load_query(p.q, gq, qx, qy, qz, qgu, qgv);
          // query tokens fc_p(q) (models.py:79): rows 13 q of the tile, 4 channels per lane
          const int gw = warp - NCW - 2;
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.fcp_b) + lane);
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.fcp_wt) + lane);
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.fcp_wt + 128) + lane);
          const float4 w2 = __ldg(reinterpret_cast<const float4*>(p.fcp_wt + 256) + lane);
#pragma unroll
          for (int q = gw; q < TILE_Q; q += NGW) {
            const float px = __shfl_sync(0xffffffffu, qx, q), py = __shfl_sync(0xffffffffu, qy, q),
                        pz = __shfl_sync(0xffffffffu, qz, q);
            if (tile * TILE_Q + q < n_q)
              __stcg(reinterpret_cast<float4*>(tokbuf + (size_t)(NTOK * q) * 128) + lane,
                     make_float4(b4.x + px * w0.x + py * w1.x + pz * w2.x, b4.y + px * w0.y + py * w1.y + pz * w2.y,
                                 b4.z + px * w0.z + py * w1.z + pz * w2.z, b4.w + px * w0.w + py * w1.w + pz * w2.w));
          }
        }
#pragma unroll 1
        for (int idx = warp - NCW - 2; idx < 112; idx += NGW) gather_step(tile, idx, tokbuf);
        __threadfence_block();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_TOKFULL0 + buf));
      }
    }
  } else if (warp >= NCW) {
  if (warp == NCW) {
    // ===================================================================== weight producer
    {
      uint32_t ph_empty = 0xffffffffu;  // one parity bit per slot ("empty"-type: first wait passes)
      int slot = 0;
      uint32_t w_e = 0;
      const uint32_t t_start = (uint32_t)clock();
      // parts [g0, g1) of the stream (the image is packed in the issuer's consumption order, see dectc_pack)
      auto stream = [&](int g0, int g1) {
#pragma unroll 1
        for (int g = g0; g < g1; ++g) {
          const uint32_t t0 = (uint32_t)clock();
          mbar_wait(bar(B_EMPTY0 + slot), (ph_empty >> slot) & 1u);
          ph_empty ^= 1u << slot;
          w_e += (uint32_t)clock() - t0;
          if (elect_one()) {
            // part g of the stream = image part g (bf16x3: hi, lo of every unit) or 2g (bf16: hi parts only);
            // this CTA's half = rows 64 rank .. +63 of each of the two k-blocks ([128 n][64 k], 16 KB each)
            const uint8_t* src = p.wimg + (size_t)(NPASS == 3 ? g : 2 * g) * UNIT_PART_BYTES + rank * 8192;
            const uint32_t dst = sbase + OFF_RING + slot * SLOT_BYTES;
            if (p.dbg & 1) {
              mbar_arrive(bar(B_FULL0 + slot));
            } else {
              mbar_arrive_expect_tx(bar(B_FULL0 + slot), SLOT_BYTES);
              bulk_g2s(dst, src, 8192, bar(B_FULL0 + slot));
              bulk_g2s(dst + 8192, src + 16384, 8192, bar(B_FULL0 + slot));
            }
          }
          __syncwarp();
          slot = (slot + 1 == NSLOT) ? 0 : slot + 1;
        }
      };
      int pending = 0;
      for (long long base = base_first; base < n_tiles; base += gridDim.x) {
        stream(0, TAIL_UNIT0 * NPART);  // layers 0,1 and the in_proj of layer 2
        ++pending;
        if (pending == TAIL_SLOTS || base + gridDim.x >= n_tiles) {
          stream(TAIL_UNIT0 * NPART, 3 * UNITS_PER_LAYER * NPART);  // tail pass: out_proj + FFN of layer 2
          pending = 0;
        }
      }
      if (lane == 0) {
        atomicAdd(&g_prof[PF_PROD_WAIT_EMPTY], (unsigned long long)w_e);
        atomicAdd(&g_prof[PF_PROD_TOTAL], (unsigned long long)((uint32_t)clock() - t_start));
      }
    }
  } else if (warp == NCW + 1 && !leader) {
    // ===================================================================== peer CTA of a pair: slot relay
    // tells the leader's MMA warp that this CTA's half of a ring slot has landed (the commit that frees the slot
    // arrives on both CTAs' "empty" barriers)
    uint32_t ph_full = 0;
    int slot = 0;
    auto relay = [&](int nparts) {
#pragma unroll 1
      for (int g = 0; g < nparts; ++g) {
        mbar_wait(bar(B_FULL0 + slot), (ph_full >> slot) & 1u);
        ph_full ^= 1u << slot;
        if (lane == 0) mbar_arrive_cluster(mapa_u32(bar(B_FULL0 + slot), 0));
        __syncwarp();
        slot = (slot + 1 == NSLOT) ? 0 : slot + 1;
      }
    };
    int pending = 0;
    for (long long base = base_first; base < n_tiles; base += gridDim.x) {
      relay(TAIL_UNIT0 * NPART);
      ++pending;
      if (pending == TAIL_SLOTS || base + gridDim.x >= n_tiles) {
        relay((3 * UNITS_PER_LAYER - TAIL_UNIT0) * NPART);
        pending = 0;
      }
    }
  } else if (warp == NCW + 1) {
    // ===================================================================== MMA issuer (leader CTA of a pair)
    // The whole warp runs the control flow (waits are warp-uniform); one elected lane issues -- the form that lets the
    // compiler keep descriptors in uniform registers (a `lane == 0` region costs ~18 instructions and a waterfall loop per
    // MMA).  As little as possible sits between two groups of MMAs.  Measured on the B200 (tools/sync_cost.cu,
    // profiles/r2_sync_cost.log): a tcgen05.mma occupies its issuing thread for ~48 cycles (the pipe needs 64), a
    // tcgen05.commit for 46, an mbarrier try_wait for 55 even when the phase is complete, elect + syncwarp for 19; a commit
    // reaches its barrier after ~210 cycles, an arrive wakes a waiter after ~200.  Whatever the issuer does beyond that is a
    // bubble of the tensor pipe: the first version of this loop (two barriers per ring slot, elect + syncwarp around every
    // step) spent ~270 cycles per 8-MMA part in such overhead -- hidden with 24 MMAs per weight unit (three-pass modes: at
    // the pipe's floor), not with 16 (fp16f8: 2.7 instead of 2.05 kcycles per FFN chunk; the FFN loop ran at the same 2.7
    // kcycles with its MMAs removed).
    {
      uint32_t ph_a = 0, ph_full = 0, ph_hr = 0;  // parity bits (one per barrier / slot)
      int slot = 0;
      uint32_t w_a = 0, w_h = 0;  // 32-bit cycle sums (wrap after ~2 s; only read in profiling runs)
#ifdef S3D_TRACE
      bool tr_on = false;
      int tr_n = 0;
#endif
      const uint32_t t_start = (uint32_t)clock();
      const uint32_t ax_hi = sbase + OFF_AX_HI, ax_lo = sbase + OFF_AX_LO;
      constexpr uint32_t ID128 = F16 ? make_idesc_f16(128, 128 * CG) : make_idesc_bf16(128, 128 * CG);
      constexpr uint32_t BKB = 8192u;  // k-block stride of this CTA's half of a part: [64 n][64 k]
      auto wait_full = [&]() -> uint32_t {
        wait_lead(B_FULL0 + slot, (ph_full >> slot) & 1u);  // (both halves of the slot: see the barrier's count)
        ph_full ^= 1u << slot;
        tc_fence_after();
        return sbase + OFF_RING + slot * SLOT_BYTES;
      };
      auto commit1 = [&](int b) {  // (elected lane)
        if (CG == 2) umma_commit_pair(bar(b));
        else umma_commit(bar(b));
      };
      auto commit = [&](int b) {
        if (elect_one()) commit1(b);
        __syncwarp();
      };
      // (the warp reconverges after every elected block: lanes running ahead into the next try_wait would suspend the
      // warp, the issuing lane with it)
      auto next_slot = [&]() {
        __syncwarp();
        slot = (slot + 1 == NSLOT) ? 0 : slot + 1;
      };
      // One weight unit = its hi part (passes A_hi.B_hi [, A_lo.B_hi]) then, for bf16x3, its lo part
      // (pass A_hi.B_lo); each part is one ring slot, released as soon as its MMAs are issued.
      // D[:, d_col .. +127] (+)= AX . W^T, A = the activation tile in shared memory (K = 128)
      auto unit_ss = [&](uint32_t d_col, bool fresh) {
        uint32_t w = wait_full();
        if (elect_one()) {  // the part's MMAs, then the commit that frees its ring slot
          issue_part<CG, (NPASS == 3 ? 2 : 1), 8, 16384u, BKB, ID128>(tmem + d_col, ax_hi, ax_lo, w, fresh);
          commit1(B_EMPTY0 + slot);
        }
        next_slot();
        if (NPASS == 3) {
          w = wait_full();
          if (elect_one()) {
            issue_part<CG, 1, 8, 16384u, BKB, ID128>(tmem + d_col, ax_hi, ax_hi, w, false);
            commit1(B_EMPTY0 + slot);
          }
          next_slot();
        }
      };
      auto wait_a = [&]() {
        const uint32_t t0 = (uint32_t)clock();
        wait_lead(B_AREADY, ph_a);
        w_a += (uint32_t)clock() - t0;
        ph_a ^= 1;
        tc_fence_after();
      };
      // QKV projection: S[:, 0:384] = X . Win^T, issued K, Q, V with a commit each: the compute warps stage K
      // while Q is in the pipe and compute the scores while V is
      auto mma_qkv = [&]() {
        wait_a();
        unit_ss(TM_S + 128, true);
        commit(B_KDONE);
        unit_ss(TM_S, true);
        commit(B_QDONE);
        unit_ss(TM_S + 256, true);
        commit(B_DDONE);
      };
      // out-proj: R += O . Wo^T (R pre-loaded with x + b_o), then the FFN over 16 hidden chunks of 128:
      // D1 = X' . W1_c^T (A = X' in shared memory) ; R += relu(D1 + b1) . W2_c^T (A = H chunk in tensor memory)
      auto mma_out_ffn = [&]() {
        wait_a();
        unit_ss(TM_R, false);
        commit(B_DDONE);
        wait_a();
        // Three chunk buffers in tensor memory; a buffer holds D1 of chunk c, then (in place) the H operand of chunk c.
        // Issue order: MMA1 of chunks 0, 1, 2; then per chunk c: wait h_ready(c), MMA2(c) [R += H_c . W2_c^T], MMA1(c + 3) into
        // the buffer MMA2(c) has just read (MMAs of one thread execute in issue order).  There is no "h_free" hand-off, and
        // between "D1 of chunk c complete" and "MMA2(c) due" the pipe has four units (~4 kcycles) of other work: the
        // hand-off latencies of this kernel (commit -> barrier -> 32 warps -> arrives -> issuer: ~2.4 kcycles per round trip,
        // traced with S3D_TRACE) are off the critical path.  The first version (two D1 buffers, one H buffer next to
        // them, h_free / h_ready per chunk) ran at 2.97 kcycles per chunk in fp16f8 mode against a 2.05-kcycle MMA floor.
        // fp16f8 mode (F8): an FFN unit = part A [fp16(w * 2^8)] for the leading term xh.wh on kind::f16, then part B
        // [e4m3(wh * 2^4) | e4m3(wl * 2^15)] for the two cross terms on kind::f8f6f4 (4 k-steps of 32 each): 16 MMAs per
        // unit instead of 24, the accumulator carries 2^15 (tc_ptx.cuh).
        auto issue1 = [&](int c, int buf) {
          const uint32_t d = tmem + TM_D1 + 128 * buf;
          if (F8) {
            uint32_t w = wait_full();
            TR(0, tr_on, 16, c - 3)
            if (elect_one()) {
              issue_part<CG, 1, 8, 16384u, BKB, ID128>(d, ax_hi, ax_hi, w, true);
              commit1(B_EMPTY0 + slot);
            }
            next_slot();
            TR(0, tr_on, 17, c - 3)
            w = wait_full();
            TR(0, tr_on, 19, c - 3)
            if (elect_one()) {
              const uint32_t al = make_desc_lo(ax_lo), ah = make_desc_lo(ax_lo + 16384u), bl = make_desc_lo(w), bh = make_desc_lo(w + 8192u);
#pragma unroll
              for (uint32_t ks = 0; ks < 4; ++ks) umma_f8_pair_lo(d, al + 2 * ks, bl + 2 * ks, ID128, 1u);
#pragma unroll
              for (uint32_t ks = 0; ks < 4; ++ks) umma_f8_pair_lo(d, ah + 2 * ks, bh + 2 * ks, ID128, 1u);
              commit1(B_EMPTY0 + slot);
              commit1(B_D1READY0 + buf);
            }
            next_slot();
            TR(0, tr_on, 20, c - 3)
          } else {
            unit_ss(TM_D1 + 128 * buf, true);
            commit(B_D1READY0 + buf);
          }
        };
        auto issue2 = [&](int c, int buf) {
          (void)c;
          const uint32_t hb = tmem + TM_D1 + 128 * buf;  // per 32-column block: [hi 16 | lo 16] or [fp16 16 | e4m3 lo 8 | e4m3 hi 8]
          uint32_t w = wait_full();
          TR(0, tr_on, 11, c)
          if (F8) {
            if (elect_one()) {
              issue_part_ts<CG, 1, 8, BKB, ID128, true>(tmem + TM_R, hb, hb, w, false);
              commit1(B_EMPTY0 + slot);
            }
            next_slot();
            TR(0, tr_on, 12, c)
            w = wait_full();
            TR(0, tr_on, 13, c)
            if (elect_one()) {
              const uint32_t bl = make_desc_lo(w), bh = make_desc_lo(w + 8192u);
#pragma unroll
              for (uint32_t ks = 0; ks < 4; ++ks) umma_f8_ts_pair_lo(tmem + TM_R, hb + 32 * ks + 16, bl + 2 * ks, ID128, 1u);
#pragma unroll
              for (uint32_t ks = 0; ks < 4; ++ks) umma_f8_ts_pair_lo(tmem + TM_R, hb + 32 * ks + 24, bh + 2 * ks, ID128, 1u);
              commit1(B_EMPTY0 + slot);
            }
            next_slot();
            TR(0, tr_on, 14, c)
            return;
          }
          if (elect_one()) {
            issue_part_ts<CG, (NPASS == 3 ? 2 : 1), 8, BKB, ID128, true>(tmem + TM_R, hb, hb + 16, w, false);
            commit1(B_EMPTY0 + slot);
          }
          next_slot();
          if (NPASS == 3) {
            w = wait_full();
            if (elect_one()) {
              issue_part_ts<CG, 1, 8, BKB, ID128, true>(tmem + TM_R, hb, hb, w, false);
              commit1(B_EMPTY0 + slot);
            }
            next_slot();
          }
        };
        int buf = 0;  // buffer of chunk c (c >= 0) = c % 3 = buffer of chunk c + 3
#pragma unroll 1
        for (int c = -NDBUF; c < NCHUNK; ++c) {
          if (c >= 0) {
            // h_ready(c): the compute warps have written the H operand of chunk c over its D1
            const uint32_t t0 = (uint32_t)clock();
            wait_lead(B_HREADY0 + buf, (ph_hr >> buf) & 1u);
            ph_hr ^= 1u << buf;
            w_h += (uint32_t)clock() - t0;
            TR(0, tr_on, 10, c)
            tc_fence_after();
            issue2(c, buf);
          }
          if (c + NDBUF < NCHUNK) issue1(c + NDBUF, buf);
          buf = (buf + 1 == NDBUF) ? 0 : buf + 1;
        }
        commit(B_DDONE);
        TR_FLUSH(0, tr_on)
      };
      int pending = 0;
      for (long long base = base_first; base < n_tiles; base += gridDim.x) {
#pragma unroll 1
        for (int layer = 0; layer < 3; ++layer) {  // (one call site for each of the two issue sequences)
#ifdef S3D_TRACE
          tr_on = blockIdx.x == 0 && base == base_first + 6 * (long long)gridDim.x && layer == 0;
#endif
          mma_qkv();
          bool ffn = true;
          if (layer == 2) {  // last layer: only token 0 of every query is consumed downstream -> batched tail pass
            ++pending;
            ffn = (pending == TAIL_SLOTS || base + gridDim.x >= n_tiles);
            if (ffn) pending = 0;
          }
          if (ffn) mma_out_ffn();
        }
      }
      if (lane == 0) {
        atomicAdd(&g_prof[PF_MMA_WAIT_A], (unsigned long long)w_a);
        atomicAdd(&g_prof[PF_MMA_WAIT_H], (unsigned long long)w_h);
        atomicAdd(&g_prof[PF_MMA_TOTAL], (unsigned long long)((uint32_t)clock() - t_start));
      }
    }
  }
  } else {
    // ===================================================================== compute warps
    // thread = (tile row r = TMEM lane, column quarter g): 32 of the 128 model channels per thread
    const int q4 = warp & 3, g = warp >> 2;
    const int r = q4 * 32 + lane;
    const int tid = threadIdx.x;
    const uint32_t trow = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
    uint8_t* ax_hi = sgen + OFF_AX_HI;
    uint8_t* ax_lo = sgen + OFF_AX_LO;
    const float* vec = reinterpret_cast<const float*>(sgen + OFF_VEC);
    float* red0 = reinterpret_cast<float*>(sgen + OFF_RED);
    uint32_t ph_d = 0, ph_d1r = 0, ph_kq = 0;
    const int qi = r / NTOK, tk = r - qi * NTOK;
    uint32_t pf[16];
#ifdef S3D_TRACE
    int tr_tile = 0, tr_n = 0;
#define TRC ((int)blockIdx.x == ((p.dbg >> 16) & 1) && warp == ((p.dbg >> 8) & 15) && tr_tile == 7 && layer == 0)
#endif
#pragma unroll
    for (int i = 0; i < 16; ++i) pf[i] = 0;
    uint32_t tprev = (uint32_t)clock();
#define lap(i)                            \
  {                                       \
    const uint32_t t_ = (uint32_t)clock(); \
    pf[i] += t_ - tprev;                  \
    tprev = t_;                           \
  }

    // LayerNorm over the 128 channels of row r, 32 of them in v[].  Each of the row's four threads reduces its own
    // 32 values to (mean, M2 = sum of squared deviations); the pairs are exchanged through smem with ONE barrier and
    // combined exactly (Chan's parallel variance): mean = avg(m_i), M2 = sum(M2_i) + 32 sum((m_i - mean)^2).
    auto layer_norm = [&](float* v, const float* w, const float* b) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) s += v[c];
      const float m = s * (1.f / 32.f);
      float d2 = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float d = v[c] - m;
        d2 = fmaf(d, d, d2);
      }
      reinterpret_cast<float2*>(red0)[r * 4 + g] = make_float2(m, d2);
      named_bar_sync(1, NCT);
      const float4 p0 = *reinterpret_cast<const float4*>(red0 + r * 8);
      const float4 p1 = *reinterpret_cast<const float4*>(red0 + r * 8 + 4);
      const float mean = (p0.x + p0.z + p1.x + p1.z) * 0.25f;
      const float e0 = p0.x - mean, e1 = p0.z - mean, e2 = p1.x - mean, e3 = p1.z - mean;
      const float M2 = (p0.y + p0.w + p1.y + p1.w) + 32.f * (e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3);
      const float rstd = rsqrtf(M2 * (1.f / 128.f) + 1e-5f);
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + 32 * g + c);
        const float4 b4 = *reinterpret_cast<const float4*>(b + 32 * g + c);
        v[c] = fmaf((v[c] - mean) * rstd, w4.x, b4.x);
        v[c + 1] = fmaf((v[c + 1] - mean) * rstd, w4.y, b4.y);
        v[c + 2] = fmaf((v[c + 2] - mean) * rstd, w4.z, b4.z);
        v[c + 3] = fmaf((v[c + 3] - mean) * rstd, w4.w, b4.w);
      }
    };
    // v[0..31] = channels 32g.. of row r -> operand A (k-block g/2, chunks 4*(g&1)..+3)
    auto store_ax = [&](const float* v) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
        store_chunk<NPASS, F16>(ax_hi + (g >> 1) * 16384, ax_lo + (g >> 1) * 16384, r, (g & 1) * 4 + cc, v + 8 * cc);
    };
    // the same as the A operand of linear1 in fp16f8 mode: fp16(x * 2^7) in the hi tile; e4m3((x - xh) * 2^11) and e4m3(xh)
    // as two [128 rows][128 B] tiles in the lo region
    auto store_ax_f8 = [&](const float* v) {
      uint32_t h[16], l8[8], h8[8];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) split8_f8(v + 8 * cc, h + 4 * cc, l8 + 2 * cc, h8 + 2 * cc);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
        *reinterpret_cast<uint4*>(ax_hi + (g >> 1) * 16384 + sw128_chunk_off(r, (g & 1) * 4 + cc)) =
            make_uint4(h[4 * cc], h[4 * cc + 1], h[4 * cc + 2], h[4 * cc + 3]);
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const uint32_t off = sw128_chunk_off(r, 2 * g + c2);
        *reinterpret_cast<uint4*>(ax_lo + off) = make_uint4(l8[4 * c2], l8[4 * c2 + 1], l8[4 * c2 + 2], l8[4 * c2 + 3]);
        *reinterpret_cast<uint4*>(ax_lo + 16384 + off) = make_uint4(h8[4 * c2], h8[4 * c2 + 1], h8[4 * c2 + 2], h8[4 * c2 + 3]);
      }
    };
    // R[r][32g..] = (v + bias) * rscale ; publish operand A + R to the MMA issuer
    auto publish = [&](float* v, const float* bias, float rscale = 1.f) {
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias + 32 * g + c);
        v[c] = (v[c] + b4.x) * rscale; v[c + 1] = (v[c + 1] + b4.y) * rscale;
        v[c + 2] = (v[c + 2] + b4.z) * rscale; v[c + 3] = (v[c + 3] + b4.w) * rscale;
      }
      tmem_st32(trow + TM_R + 32 * g, v);
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async_smem();
      arrive_lead(B_AREADY);
    };

    float* const tokbase = p.scratch + (size_t)gridDim.x * (TAIL_SLOTS * TILE_Q) * 256 + (size_t)blockIdx.x * (2 * 128 * 128);
    // Self-attention over the 13 tokens of each query, 4 heads (one per column group = per thread of a row).
    // K (then V) of all heads is staged as an fp32 matrix [117 rows][ST_PITCH] in the H region.  The pitch of
    // 132 floats puts the key rows of the (at most four) queries a warp touches at once in distinct banks, and
    // every address is "row base of my query + compile-time offset": no address arithmetic in the loops.
    // The dot products run as packed fp32 FMAs (FFMA2).  (Measured and rejected: two rows of a query per thread, 252
    // tasks on 8 warps with Q staged to shared memory as well -- halves the LDS traffic, but the extra staging pass, the
    // wait for the V projection before it and the lower occupancy of the loops cost more: 186.5 vs 183.3 kcycles.)
    float* const st = reinterpret_cast<float*>(sgen + OFF_H);
    const float* const qrows = st + (qi < TILE_Q ? qi : 0) * (NTOK * ST_PITCH) + 32 * g;  // key/value rows of my query, my head
    // S[:, col + 32g ..+31] + bias -> my row of the staging matrix
    auto stage_kv = [&](uint32_t col, const float* bias) {
      float kk[32];
      tmem_ld32(trow + col + 32 * g, kk);
      tmem_ld_wait();
      if (r < TILE_Q * NTOK) {
        float4* dst = reinterpret_cast<float4*>(st + r * ST_PITCH + 32 * g);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + 32 * g + 4 * c);
          dst[c] = make_float4(kk[4 * c] + b4.x, kk[4 * c + 1] + b4.y, kk[4 * c + 2] + b4.z, kk[4 * c + 3] + b4.w);
        }
      }
    };
    // b_in of the next layer replaces this layer's once every thread is past its last use (after the "V staged"
    // barrier); the FFN-phase vectors of this layer are fetched at the start of attention and stored after the
    // "K staged" barrier (every thread is past LayerNorm 2 of the previous layer by then).
    auto vecB_fetch = [&](int layer, float4* v2) {
      const float4* src = reinterpret_cast<const float4*>(p.vecs + (size_t)layer * VEC_FLOATS) + V_PART_B / 4;
      v2[0] = __ldg(src + tid);
      v2[1] = (tid + NCT < (VEC_FLOATS - V_PART_B) / 4) ? __ldg(src + tid + NCT) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto vecB_store = [&](const float4* v2) {
      float4* dst = reinterpret_cast<float4*>(sgen + OFF_VEC) + V_PART_B / 4;
      // fp16f8: b1 is staged as 128 b1 (the linear1 epilogue produces 128 h, see split8_f8_h)
      const float s0 = (F8 && tid >= (V_B1 - V_PART_B) / 4) ? F8_XS : 1.f, s1 = F8 ? F8_XS : 1.f;
      dst[tid] = make_float4(v2[0].x * s0, v2[0].y * s0, v2[0].z * s0, v2[0].w * s0);
      if (tid + NCT < (VEC_FLOATS - V_PART_B) / 4) dst[tid + NCT] = make_float4(v2[1].x * s1, v2[1].y * s1, v2[1].z * s1, v2[1].w * s1);
    };
    auto vecA_next = [&](int layer) {  // b_in of the layer after `layer`
      if (tid < V_PART_B / 4) {
        const int nl = layer == 2 ? 0 : layer + 1;
        reinterpret_cast<float4*>(sgen + OFF_VEC)[tid] =
            __ldg(reinterpret_cast<const float4*>(p.vecs + (size_t)nl * VEC_FLOATS) + tid);
      }
    };
    // ---- attention of the full layers on the warp-level tensor-core path (mma.sync m16n8k16, fp16 hi/lo pairs, 3 passes)
    // Q (scaled), K, V (+ biases) are staged as fp16 hi / lo matrices [120 rows][128 channels] with rows of 256 bytes
    // whose 16-byte chunks are XOR-swizzled with (row & 7): the staging stores (one row per lane) and the ldmatrix
    // reads (8 consecutive rows per 8x8 matrix) are both conflict free.  K / V live in the H region, Q in the activation
    // operand region (free once the V projection has completed).  The 36 (query, head) problems of a tile are dealt to
    // the 16 compute warps; a problem is one 16-row tile (13 token rows + 3 rows of the next query, discarded):
    //   S = Q K^T   m16 n16 k32: 2 n-tiles x 2 k-steps x 3 passes = 12 MMAs; masked softmax on the accumulator fragments
    //   O = P V     m16 n32 k16: 4 n-tiles x 3 passes = 12 MMAs (P is re-used from the S fragments: the accumulator
    //               layout of two n-tiles IS the A layout of one k16 step), V through ldmatrix.trans
    // and O goes straight into the out-proj A operand.  (The CUDA-core version spent 15 kcycles per layer in LDS-bound
    // FMAs: every key / value row was re-read by the 13 row owners of its query.)
    constexpr uint32_t ATT_ARR = 120u * 256u;  // one staged matrix (hi or lo)
    auto att_off = [](int row, int chunk) -> uint32_t { return (uint32_t)(row * 256 + ((chunk ^ (row & 7)) << 4)); };
    auto stage_att = [&](uint32_t col, const float* bias, float scale, uint8_t* dst) {
      const float live = r < TILE_Q * NTOK ? scale : 0.f;  // rows 117..119 are read as padding: exact zeros
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {  // two halves of 16 columns: small live state next to the P fragments
        float v[16];
        tmem_ld16(trow + col + 32 * g + 16 * hf, v);
        tmem_ld_wait();
        if (r < 120) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float4 b0 = *reinterpret_cast<const float4*>(bias + 32 * g + 16 * hf + 8 * c);
            const float4 b1 = *reinterpret_cast<const float4*>(bias + 32 * g + 16 * hf + 8 * c + 4);
            const float w[8] = {(v[8 * c] + b0.x) * live,     (v[8 * c + 1] + b0.y) * live, (v[8 * c + 2] + b0.z) * live,
                                (v[8 * c + 3] + b0.w) * live, (v[8 * c + 4] + b1.x) * live, (v[8 * c + 5] + b1.y) * live,
                                (v[8 * c + 6] + b1.z) * live, (v[8 * c + 7] + b1.w) * live};
            uint32_t h[4], l[4];
            split8_hn(w, h, l);
            const uint32_t off = att_off(r, 4 * g + 2 * hf + c);
            *reinterpret_cast<uint4*>(dst + off) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(dst + ATT_ARR + off) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
      }
    };
    auto attention = [&](int layer, bool valid) {
      (void)valid;  // rows of queries beyond n carry zero tokens: finite everywhere, their outputs are never stored
      const float* b_in = vec + V_BIN;
      constexpr int NPROB = TILE_Q * 4, PPW = (NPROB + NCW - 1) / NCW;  // 36 problems, at most 3 per warp
      const uint32_t kv_s = sbase + OFF_H, q_s = sbase + OFF_AX_HI;
      const int l8 = lane & 7, mi = lane >> 3, tq = lane & 3, rq = lane >> 2;
      uint32_t phi[PPW][4], plo[PPW][4];
      {
        float4 vb[2];
        vecB_fetch(layer, vb);
        stage_att(TM_S + 128, b_in + 128, 1.f, sgen + OFF_H);  // K
        lap(12)
        mbar_wait(bar(B_QDONE), ph_kq);
        ph_kq ^= 1u;
        mbar_wait(bar(B_DDONE), ph_d);  // V projected: every MMA that reads the activation operand has completed
        ph_d ^= 1;
        tc_fence_after();
        stage_att(TM_S, b_in, 0.17677669529663687f, sgen + OFF_AX_HI);  // Q / sqrt(32)
        named_bar_sync(1, NCT);
        vecB_store(vb);
      }
#pragma unroll
      for (int i = 0; i < PPW; ++i) {
        const int pr = warp + NCW * i;
        if (pr < NPROB) {  // warp-uniform
          const int R0 = NTOK * (pr >> 2), hh = pr & 3;
          float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};  // n-tiles: keys 0-7, keys 8-15
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const int c0 = 4 * hh + 2 * ks;  // first 16-byte chunk of this k-step
            uint32_t qh[4], ql[4], kh[4], kl[4];
            const uint32_t qa = q_s + att_off(R0 + l8 + 8 * (mi & 1), c0 + (mi >> 1));
            ldsm_x4(qa, qh[0], qh[1], qh[2], qh[3]);
            ldsm_x4(qa + ATT_ARR, ql[0], ql[1], ql[2], ql[3]);
            const uint32_t ka = kv_s + att_off(R0 + l8 + 8 * (mi >> 1), c0 + (mi & 1));
            ldsm_x4(ka, kh[0], kh[1], kh[2], kh[3]);
            ldsm_x4(ka + ATT_ARR, kl[0], kl[1], kl[2], kl[3]);
            mma_f16_16816(s0, ql, kh[0], kh[1]);
            mma_f16_16816(s1, ql, kh[2], kh[3]);
            mma_f16_16816(s0, qh, kl[0], kl[1]);
            mma_f16_16816(s1, qh, kl[2], kl[3]);
            mma_f16_16816(s0, qh, kh[0], kh[1]);
            mma_f16_16816(s1, qh, kh[2], kh[3]);
          }
          // fragment: s0[0,1] / s1[0,1] = row rq, keys 2 tq + {0,1} / 8 + 2 tq + {0,1};  [2,3] = row rq + 8.  Keys >= 13 masked.
          // (keys >= K + 1: the rows of the next query, or the dead rows of a model with fewer than 12 slices)
          const int Lk = p.K + 1;
          const bool k0 = 2 * tq < Lk, k1 = 2 * tq + 1 < Lk, m0 = (8 + 2 * tq) < Lk, m1 = (9 + 2 * tq) < Lk;
#pragma unroll
          for (int hr = 0; hr < 2; ++hr) {
            float a0 = k0 ? s0[2 * hr] : -INFINITY, a1 = k1 ? s0[2 * hr + 1] : -INFINITY;
            float a2 = m0 ? s1[2 * hr] : -INFINITY, a3 = m1 ? s1[2 * hr + 1] : -INFINITY;
            float mx = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            a0 = k0 ? __expf(a0 - mx) : 0.f;
            a1 = k1 ? __expf(a1 - mx) : 0.f;
            a2 = m0 ? __expf(a2 - mx) : 0.f;
            a3 = m1 ? __expf(a3 - mx) : 0.f;
            float sum = (a0 + a1) + (a2 + a3);
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float inv = 1.f / sum;
            // A fragment of P: a0a1 (row, keys 0-7), a2a3 (row + 8, keys 0-7), a4a5 (row, keys 8-15), a6a7 (row + 8, keys 8-15)
            split2_h(a0 * inv, a1 * inv, phi[i][hr], plo[i][hr]);
            split2_h(a2 * inv, a3 * inv, phi[i][2 + hr], plo[i][2 + hr]);
          }
        }
      }
      lap(13)
      named_bar_sync(1, NCT);  // everyone has read K and Q
      stage_att(TM_S + 256, b_in + 256, 1.f, sgen + OFF_H);  // V over K
      if (r >= TILE_Q * NTOK) {  // padding rows of the out-proj operand (the region held Q until now): zeros
        const float z[32] = {};
        store_ax(z);
      }
      named_bar_sync(1, NCT);
      vecA_next(layer);
      lap(14)
#pragma unroll
      for (int i = 0; i < PPW; ++i) {
        const int pr = warp + NCW * i;
        if (pr < NPROB) {
          const int R0 = NTOK * (pr >> 2), hh = pr & 3;
          float o[4][4];
#pragma unroll
          for (int jn = 0; jn < 4; ++jn)
#pragma unroll
            for (int e = 0; e < 4; ++e) o[jn][e] = 0.f;
#pragma unroll
          for (int jp = 0; jp < 2; ++jp) {  // channel n-tiles 2 jp, 2 jp + 1 of head hh
            uint32_t vh[4], vl[4];
            const uint32_t va = kv_s + att_off(R0 + l8 + 8 * (mi & 1), 4 * hh + 2 * jp + (mi >> 1));
            ldsm_x4_t(va, vh[0], vh[1], vh[2], vh[3]);
            ldsm_x4_t(va + ATT_ARR, vl[0], vl[1], vl[2], vl[3]);
            mma_f16_16816(o[2 * jp], plo[i], vh[0], vh[1]);
            mma_f16_16816(o[2 * jp + 1], plo[i], vh[2], vh[3]);
            mma_f16_16816(o[2 * jp], phi[i], vl[0], vl[1]);
            mma_f16_16816(o[2 * jp + 1], phi[i], vl[2], vl[3]);
            mma_f16_16816(o[2 * jp], phi[i], vh[0], vh[1]);
            mma_f16_16816(o[2 * jp + 1], phi[i], vh[2], vh[3]);
          }
          // O[row][32 hh + 8 jn + 2 tq + {0,1}] -> out-proj A operand (k-block hh / 2, chunk 4 (hh & 1) + jn, element pair tq)
          uint8_t* const thi = ax_hi + (hh >> 1) * 16384;
          uint8_t* const tlo = ax_lo + (hh >> 1) * 16384;
#pragma unroll
          for (int hr = 0; hr < 2; ++hr) {
            const int row = rq + 8 * hr;
            if (row < NTOK) {
#pragma unroll
              for (int jn = 0; jn < 4; ++jn) {
                const uint32_t off = sw128_chunk_off(R0 + row, 4 * (hh & 1) + jn) + 4 * tq;
                uint32_t hv, lv;
                if (F16) {
                  split2_h(o[jn][2 * hr], o[jn][2 * hr + 1], hv, lv);
                } else {
                  const __nv_bfloat162 h2 = __floats2bfloat162_rn(o[jn][2 * hr], o[jn][2 * hr + 1]);
                  hv = *reinterpret_cast<const uint32_t*>(&h2);
                  const __nv_bfloat162 l2 = __floats2bfloat162_rn(o[jn][2 * hr] - __uint_as_float(hv << 16),
                                                                  o[jn][2 * hr + 1] - __uint_as_float(hv & 0xffff0000u));
                  lv = *reinterpret_cast<const uint32_t*>(&l2);
                }
                *reinterpret_cast<uint32_t*>(thi + off) = hv;
                if (NPASS == 3) *reinterpret_cast<uint32_t*>(tlo + off) = lv;
              }
            }
          }
        }
      }
      lap(15)
      fence_proxy_async_smem();
      arrive_lead(B_AREADY);
      lap(PF_ATTN)
    };
    // Last layer: only token 0 of a query is consumed downstream (models.py:83), so only those 9 rows attend.  The
    // work is re-dealt over the CTA instead of leaving it with the 9 row owners: 468 threads take one (query, head,
    // key) dot product each, 288 threads one float4 of the output.  The attention output and the residual (x + b_o)
    // of the token-0 rows are parked in the CTA's global scratch rows `slot*9 + q` for the tail pass.
    auto attention_tok0 = [&](long long tile, int slot) {
      const float* b_in = vec + V_BIN;
      float* const qs = reinterpret_cast<float*>(sgen + OFF_RED);  // [9][128] scaled queries of the token-0 rows
      float* const scs = reinterpret_cast<float*>(sgen + OFF_SC);  // [9][4][13] scores
      float* const tail0 = p.scratch + ((size_t)blockIdx.x * (TAIL_SLOTS * TILE_Q) + (size_t)slot * TILE_Q) * 256;
      {
        float4 vb[2];
        vecB_fetch(2, vb);
        stage_kv(TM_S + 128, b_in + 128);
        mbar_wait(bar(B_QDONE), ph_kq);
        ph_kq ^= 1u;
        tc_fence_after();
        {
          float qq[32];
          tmem_ld32(trow + TM_S + 32 * g, qq);
          tmem_ld_wait();
          if (tk == 0 && qi < TILE_Q) {
            float4* dst = reinterpret_cast<float4*>(qs + qi * 128 + 32 * g);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 b4 = *reinterpret_cast<const float4*>(b_in + 32 * g + 4 * c);
              dst[c] = make_float4((qq[4 * c] + b4.x) * 0.17677669529663687f, (qq[4 * c + 1] + b4.y) * 0.17677669529663687f,
                                   (qq[4 * c + 2] + b4.z) * 0.17677669529663687f, (qq[4 * c + 3] + b4.w) * 0.17677669529663687f);
            }
          }
        }
        named_bar_sync(1, NCT);
        vecB_store(vb);
      }
      lap(12)
      if (tid < TILE_Q * 4 * NTOK) {
        const int q = tid / (4 * NTOK), rem = tid - q * (4 * NTOK), hh = rem / NTOK, j = rem - hh * NTOK;
        const float4* qp = reinterpret_cast<const float4*>(qs + q * 128 + 32 * hh);
        const float4* kp = reinterpret_cast<const float4*>(st + (q * NTOK + j) * ST_PITCH + 32 * hh);
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 q4v = qp[c], k4 = kp[c];
          ffma2(a0, make_float2(q4v.x, q4v.y), make_float2(k4.x, k4.y));
          ffma2(a1, make_float2(q4v.z, q4v.w), make_float2(k4.z, k4.w));
        }
        scs[tid] = (a0.x + a0.y) + (a1.x + a1.y);
      }
      lap(13)
      mbar_wait(bar(B_DDONE), ph_d);  // V projected
      ph_d ^= 1;
      tc_fence_after();
      named_bar_sync(1, NCT);  // scores written, everyone has read K
      stage_kv(TM_S + 256, b_in + 256);
      {  // residual x + b_o of the token-0 rows (the pre-loaded accumulator of out-proj)
        float v[32];
        tmem_ld32(trow + TM_R + 32 * g, v);
        tmem_ld_wait();
        if (tk == 0 && qi < TILE_Q) {
          float4* dst = reinterpret_cast<float4*>(tail0 + (size_t)qi * 256 + 128 + 32 * g);
#pragma unroll
          for (int c = 0; c < 8; ++c) dst[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        }
      }
      named_bar_sync(1, NCT);
      vecA_next(2);
      lap(14)
      if (tid < TILE_Q * 32) {
        const int q = tid >> 5, c4 = tid & 31, hh = c4 >> 3;
        const float* sp = scs + q * (4 * NTOK) + hh * NTOK;
        float pj[NTOK];
        float mx = sp[0];
#pragma unroll
        for (int j = 0; j < NTOK; ++j) {
          pj[j] = j <= p.K ? sp[j] : -INFINITY;  // keys beyond the model's K + 1 tokens are dead rows
          mx = fmaxf(mx, pj[j]);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < NTOK; ++j) {
          pj[j] = j <= p.K ? __expf(pj[j] - mx) : 0.f;
          sum += pj[j];
        }
        const float inv = 1.f / sum;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* vp = st + (q * NTOK) * ST_PITCH + 4 * c4;
#pragma unroll
        for (int j = 0; j < NTOK; ++j) {
          const float4 v4 = *reinterpret_cast<const float4*>(vp + j * ST_PITCH);
          o.x = fmaf(pj[j], v4.x, o.x);
          o.y = fmaf(pj[j], v4.y, o.y);
          o.z = fmaf(pj[j], v4.z, o.z);
          o.w = fmaf(pj[j], v4.w, o.w);
        }
        if (tile * TILE_Q + q < n_q)
          *reinterpret_cast<float4*>(tail0 + (size_t)q * 256 + 4 * c4) = make_float4(o.x * inv, o.y * inv, o.z * inv, o.w * inv);
      }
      lap(15)
      lap(PF_ATTN)
    };
    // out-proj result + residual -> LayerNorm 1 -> FFN -> LayerNorm 2 -> next layer's operand / fc_out head
    auto post_attn = [&](int layer, bool head_valid, long long head_q) {
        // -------------------------------------------------------------- residual + LayerNorm 1 (in place in TMEM)
        mbar_wait(bar(B_DDONE), ph_d);
        ph_d ^= 1;
        tc_fence_after();
        lap(PF_WAIT_OUT)
        {
          float v[32];
          tmem_ld32(trow + TM_R + 32 * g, v);
          tmem_ld_wait();
          layer_norm(v, vec + V_LN1W, vec + V_LN1B);
          // X' -> A operand of linear1 (the activation tile is idle between out-proj and the next layer)
          if (F8) {
            store_ax_f8(v);
            publish(v, vec + V_B2, 32768.f);  // the FFN accumulates at scale 2^15 on top of the residual
          } else {
            store_ax(v);
            publish(v, vec + V_B2);
          }
        }
        lap(PF_LN1)
        // -------------------------------------------------------------- FFN hidden chunks (32 columns per thread)
        {
          int buf = 0;  // chunk buffer c % 3
#pragma unroll 1
          for (int c = 0; c < NCHUNK; ++c) {
            mbar_wait(bar(B_D1READY0 + buf), (ph_d1r >> buf) & 1u);
            ph_d1r ^= 1u << buf;
            tc_fence_after();
            TR(1, TRC, 30, c)
            lap(PF_FFN_WAIT_D1)
            const uint32_t tb = trow + TM_D1 + 128 * buf + 32 * g;  // my 32 columns of the chunk buffer
            float d[32];
            tmem_ld32(tb, d);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = *reinterpret_cast<const float4*>(vec + V_B1 + c * 128 + 32 * g + 4 * j);
              const float ds = F8 ? F8_ACC_INV * F8_XS : 1.f;  // (fp16f8: 128 h; the staged b1 carries the 2^7 too)
              d[4 * j] = fmaxf(fmaf(d[4 * j], ds, b4.x), 0.f);
              d[4 * j + 1] = fmaxf(fmaf(d[4 * j + 1], ds, b4.y), 0.f);
              d[4 * j + 2] = fmaxf(fmaf(d[4 * j + 2], ds, b4.z), 0.f);
              d[4 * j + 3] = fmaxf(fmaf(d[4 * j + 3], ds, b4.w), 0.f);
            }
            lap(PF_FFN_MATH)
            TR(1, TRC, 31, c)
            {  // H chunk -> A operand of linear2, written over the D1 values it was computed from (my own 32 columns:
              // nobody else reads or writes them until MMA2 of this chunk)
              uint32_t hh[16], hl[16];  // (fp16f8: hl[0..7] = e4m3 lo parts, hl[8..15] = e4m3 hi parts)
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (F8) split8_f8_h(d + 8 * j, hh + 4 * j, hl + 2 * j, hl + 8 + 2 * j);
                else split8x<NPASS == 3, F16>(d + 8 * j, hh + 4 * j, hl + 4 * j);
              }
              tmem_st16(tb, hh);
              if (F8) {
                tmem_st8(tb + 16, hl);
                tmem_st8(tb + 24, hl + 8);
              } else if (NPASS == 3) {
                tmem_st16(tb + 16, hl);
              }
              tmem_st_wait();
            }
            tc_fence_before();  // orders this thread's D1 load / H store before the MMAs that follow the arrive
            TR(1, TRC, 33, c)
            arrive_lead(B_HREADY0 + buf);
            TR(1, TRC, 34, c)
            lap(PF_FFN_STORE)
            buf = (buf + 1 == NDBUF) ? 0 : buf + 1;
          }
          TR_FLUSH(1, TRC)
        }
        // -------------------------------------------------------------- residual + LayerNorm 2
        mbar_wait(bar(B_DDONE), ph_d);
        ph_d ^= 1;
        tc_fence_after();
        lap(PF_WAIT_FFN)
        {
          float v[32];
          tmem_ld32(trow + TM_R + 32 * g, v);
          tmem_ld_wait();
          if (F8) {
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] *= F8_ACC_INV;
          }
          layer_norm(v, vec + V_LN2W, vec + V_LN2B);
          if (layer < 2) {
            store_ax(v);
            publish(v, vec + V_BONEXT);
          } else {  // fc_out on token 0 (models.py:83-84)
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) acc = fmaf(v[c], __ldg(p.fco_w + 32 * g + c), acc);
            named_bar_sync(1, NCT);  // everyone has read the LayerNorm partials
            red0[r * 4 + g] = acc;
            named_bar_sync(1, NCT);
            if (head_valid) {
              const float4 a4 = *reinterpret_cast<const float4*>(red0 + r * 4);
              p.out[out_index(p.q, head_q)] = p.out_scale * (a4.x + a4.y + a4.z + a4.w + __ldg(p.fco_b));
            }
          }
        }
        lap(PF_LN2)
    };
    if (tid < V_PART_B / 4)  // b_in of layer 0 (later layers / tiles: staged by vecA_next during the previous attention)
      reinterpret_cast<float4*>(sgen + OFF_VEC)[tid] = __ldg(reinterpret_cast<const float4*>(p.vecs) + tid);
    named_bar_sync(1, NCT);
    int pending = 0, tile_it = 0, tiles_done = 0;
    uint32_t ph_tf = 0;
    long long batch_tile0 = 0;
    for (long long tile = tile_first; tile - rank < n_tiles; tile += gridDim.x) {
      const long long q_idx = tile * TILE_Q + qi;
      const bool valid = (qi < TILE_Q) && (q_idx < n_q);
      // ------------------------------------------------------------------ token build
      const int tbuf = tile_it & 1;
      const float* tokrow;
      if (p.q.tok_slice) {  // ready tokens in global memory: query token [n][128], slice tokens [n][12][128]
        tokrow = tk == 0 ? p.q.tok_query + (size_t)(valid ? q_idx : 0) * 128
                         : p.q.tok_slice + ((size_t)(valid ? q_idx : 0) * p.K + (tk <= p.K ? tk - 1 : 0)) * 128;
      } else {
        tokrow = tokbase + (size_t)tbuf * (128 * 128) + (size_t)r * 128;
        mbar_wait(bar(B_TOKFULL0 + tbuf), (ph_tf >> tbuf) & 1u);  // slice tokens of this tile gathered
        ph_tf ^= 1u << tbuf;
      }
      {
        float v[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid && tk <= p.K) t4 = __ldcg(reinterpret_cast<const float4*>(tokrow + 32 * g) + c);  // dead rows: zero tokens
          v[4 * c] = t4.x; v[4 * c + 1] = t4.y; v[4 * c + 2] = t4.z; v[4 * c + 3] = t4.w;
        }
        warp_arrive(bar(B_TOKEMPTY0 + tbuf), lane);  // this warp has read its tokens: the buffer may be refilled
        store_ax(v);
        publish(v, p.b_o0);
      }
      ++tile_it;
#ifdef S3D_TRACE
      ++tr_tile;
#endif
      lap(PF_TOKEN)

#pragma unroll 1
      for (int layer = 0; layer < 3; ++layer) {
        // (this layer's vectors: b_in was staged during the previous layer's attention, the FFN-phase block is
        // staged inside attention -- no barrier here)
        // -------------------------------------------------------------- attention (13x13 per query and head)
        mbar_wait(bar(B_KDONE), ph_kq);  // K projected (Q and V follow, see mma_qkv)
        tc_fence_after();
        lap(PF_WAIT_QKV)
        // (post_attn has ONE call site: the unrolled FFN epilogue is large and the kernel's code already exceeds the
        // instruction cache)
        bool post = true, head_valid = false;
        long long head_q = 0;
        if (layer < 2) {
          attention(layer, valid);
        } else {
          attention_tok0(tile, pending);
          if (pending == 0) batch_tile0 = tile;
          ++pending;
          post = (pending == TAIL_SLOTS || tile - rank + gridDim.x >= n_tiles);
          if (post) {
            // -------------------------------------------------------------- tail pass: token 0 of up to 126 queries
            named_bar_sync(1, NCT);  // scratch rows written by other threads are visible after the CTA barrier
            const int k = r / TILE_Q, qq = r - k * TILE_Q;
            const long long tq = (batch_tile0 + (long long)k * gridDim.x) * TILE_Q + qq;
            const bool tvalid = (r < pending * TILE_Q) && (tq < n_q);
            const float* row = p.scratch + ((size_t)blockIdx.x * (TAIL_SLOTS * TILE_Q) + (r < TAIL_SLOTS * TILE_Q ? r : 0)) * 256;
            float v[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (tvalid) t4 = __ldcg(reinterpret_cast<const float4*>(row + 32 * g) + c);
              v[4 * c] = t4.x; v[4 * c + 1] = t4.y; v[4 * c + 2] = t4.z; v[4 * c + 3] = t4.w;
            }
            store_ax(v);  // attention output -> operand A of out-proj
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (tvalid) t4 = __ldcg(reinterpret_cast<const float4*>(row + 128 + 32 * g) + c);
              v[4 * c] = t4.x; v[4 * c + 1] = t4.y; v[4 * c + 2] = t4.z; v[4 * c + 3] = t4.w;
            }
            tmem_st32(trow + TM_R + 32 * g, v);  // x + b_o -> accumulator of out-proj
            tmem_st_wait();
            tc_fence_before();
            fence_proxy_async_smem();
            arrive_lead(B_AREADY);
            head_valid = (g == 0) && tvalid;
            head_q = tq;
            pending = 0;
          }
        }
        if (post) post_attn(layer, head_valid, head_q);
      }
      ++tiles_done;
    }
    if (tid == 0) {  // phase counters of this CTA's warp 0 (cycles; 32-bit sums are enough for ~10^4 tiles per CTA)
      atomicAdd(&g_prof[PF_TILES], (unsigned long long)tiles_done);
#pragma unroll
      for (int i = 0; i < 12; ++i) atomicAdd(&g_prof[i], (unsigned long long)pf[i]);
      for (int i = 12; i < 16; ++i) atomicAdd(&g_prof[i + 8], (unsigned long long)pf[i]);
    }
#undef lap
  }
#ifdef S3D_TRACE
  __syncthreads();
  if (blockIdx.x == 1) {  // (the peer's trace flush may still be in flight: crude delay)
    for (int i = 0; i < 2000; ++i) asm volatile("nanosleep.u32 100;");
  }
  if (blockIdx.x == 1 && threadIdx.x == 0) {
    for (int i = 0; i < 256; ++i)
      if (g_trace[i]) printf("TR %d %u %u %u\n", i >> 7, g_trace[i] >> 26, (g_trace[i] >> 22) & 15u, g_trace[i] & 0x3fffffu);
    for (int i = 0; i < 256; ++i) g_trace[i] = 0;
  }
#endif
  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // neither CTA frees tensor memory (or exits) while the pair's MMAs may touch it
  else __syncthreads();
  if (warp == 0) {
    if (CG == 2) tmem_dealloc_pair(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

// ---- self-test: one 128-row UMMA tile against one weight unit -------------------------------------
// D[128][128] = A[128][128] . W[128][128]^T, single CTA (cta_group::1).
// mode 0: A in shared memory (the QKV / out-proj / linear1 path); mode 1: A in tensor memory (the linear2 path).
template <int NPASS, bool F16>
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A, const uint8_t* wimg, int mode,
                                                               float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const int warp = threadIdx.x >> 5;
  const int r = threadIdx.x;
  const uint32_t full = sbase + OFF_BAR, done = sbase + OFF_BAR + 8;
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(sbase + OFF_TMEMPTR, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + OFF_TMEMPTR);
  const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  if (threadIdx.x == 0) {  // the whole unit (hi 32 KB | lo 32 KB) lands in the H + ring area
    mbar_arrive_expect_tx(full, UNIT_STRIDE_BYTES);
    for (uint32_t o = 0; o < UNIT_STRIDE_BYTES; o += 16384) bulk_g2s(sbase + OFF_H + o, wimg + o, 16384, full);
  }
  if (mode == 0) {
    for (int kc = 0; kc < 16; ++kc) {
      float v[8];
      for (int i = 0; i < 8; ++i) v[i] = A[r * 128 + kc * 8 + i];
      store_chunk<NPASS, F16>(sgen + OFF_AX_HI + (kc >> 3) * 16384, sgen + OFF_AX_LO + (kc >> 3) * 16384, r, kc & 7, v);
    }
    fence_proxy_async_smem();
  } else {  // the decoder's in-place H layout (TM_D1): per 32 k values one 32-column block [hi 16 | lo 16]
    for (int j = 0; j < 4; ++j) {
      float v[32];
      for (int i = 0; i < 32; ++i) v[i] = A[r * 128 + 32 * j + i];
      uint32_t h[16], l[16];
      for (int c = 0; c < 4; ++c) split8x<NPASS == 3, F16>(v + 8 * c, h + 4 * c, l + 4 * c);
      tmem_st16(trow + TM_HT + 32 * j, h);
      if (NPASS == 3) tmem_st16(trow + TM_HT + 32 * j + 16, l);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_wait(full, 0);
    tc_fence_after();
    const uint32_t w = sbase + OFF_H;
    constexpr uint32_t ID = F16 ? make_idesc_f16(128) : make_idesc_bf16(128);
    if (mode == 0) {
      issue_part<1, (NPASS == 3 ? 2 : 1), 8, 16384u, 16384u, ID>(tmem, sbase + OFF_AX_HI, sbase + OFF_AX_LO, w, true);
      if (NPASS == 3) issue_part<1, 1, 8, 16384u, 16384u, ID>(tmem, sbase + OFF_AX_HI, sbase + OFF_AX_HI, w + UNIT_PART_BYTES, false);
    } else {
      issue_part_ts<1, (NPASS == 3 ? 2 : 1), 8, 16384u, ID, true>(tmem, tmem + TM_HT, tmem + TM_HT + 16, w, true);
      if (NPASS == 3) issue_part_ts<1, 1, 8, 16384u, ID, true>(tmem, tmem + TM_HT, tmem + TM_HT, w + UNIT_PART_BYTES, false);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  for (int j = 0; j < 4; ++j) {
    float v[32];
    tmem_ld32(trow + 32 * j, v);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) D[r * 128 + 32 * j + c] = v[c];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- host-side packing ---------------------------------------------------------------------------
inline uint16_t bf16_bits(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }
inline float bf16_val(uint16_t b) {
  uint32_t u = static_cast<uint32_t>(b) << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

// W(n, k) for n in [0,NU), k in [0,KU) -> hi image then lo image (each NU*KU bf16, SW128 K-major tiles).
inline uint16_t f16_bits(float x) { return __half_as_ushort(__float2half_rn(x)); }
inline float f16_val(uint16_t b) { return __half2float(__ushort_as_half(b)); }

template <class F>
void pack_unit(F W, int NU, int KU, uint8_t* dst, bool f16 = false) {
  uint16_t* hi = reinterpret_cast<uint16_t*>(dst);
  uint16_t* lo = reinterpret_cast<uint16_t*>(dst + UNIT_PART_BYTES);
  for (int kb = 0; kb < KU / 64; ++kb)
    for (int n = 0; n < NU; ++n)
      for (int k = 0; k < 64; ++k) {
        const float w = W(n, kb * 64 + k);
        const uint16_t h = f16 ? f16_bits(fminf(fmaxf(w, -65504.f), 65504.f)) : bf16_bits(w);
        const uint16_t l = f16 ? f16_bits(w - f16_val(h)) : bf16_bits(w - bf16_val(h));
        const size_t off = (size_t)kb * NU * 128 + sw128_chunk_off(n, k >> 3) + (k & 7) * 2;
        hi[off / 2] = h;
        lo[off / 2] = l;
      }
}

// fp16f8 unit: part A = fp16(w * 2^8) as [2 k-blocks][128 n][64 k]; part B = e4m3(fp16(w) * 2^4) as [128 n][128 k] followed
// by e4m3((w - fp16(w)) * 2^15) as [128 n][128 k]; all tiles K-major / 128-byte swizzle.
template <class F>
void pack_unit_f8(F W, uint8_t* dst) {
  uint16_t* hi = reinterpret_cast<uint16_t*>(dst);
  uint8_t* wh8 = dst + UNIT_PART_BYTES;
  uint8_t* wl8 = dst + UNIT_PART_BYTES + 16384;
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 128; ++k) {
      const float w = W(n, k);
      const float wh = f16_val(f16_bits(w));
      const int kb = k >> 6, kk = k & 63;
      hi[((size_t)kb * 128 * 128 + sw128_chunk_off(n, kk >> 3) + (kk & 7) * 2) / 2] = f16_bits(wh * 256.f);
      const size_t o8 = sw128_chunk_off(n, k >> 4) + (k & 15);
      wh8[o8] = (uint8_t)__nv_cvt_float_to_fp8(wh * 16.f, __NV_SATFINITE, __NV_E4M3);
      wl8[o8] = (uint8_t)__nv_cvt_float_to_fp8((w - wh) * 32768.f, __NV_SATFINITE, __NV_E4M3);
    }
}

// Self-test of the fp16 + 2 x fp8 unit (the FFN contractions of S3D_PREC_FP16F8): D = A . W^T on one tile, single CTA.
// mode 0: A in shared memory (linear1's path), mode 1: A in tensor memory (linear2's path).
__global__ void __launch_bounds__(128, 1) umma_selftest_f8_kernel(const float* __restrict__ A, const uint8_t* wimg, int mode,
                                                                  float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const int warp = threadIdx.x >> 5;
  const int r = threadIdx.x;
  const uint32_t full = sbase + OFF_BAR, done = sbase + OFF_BAR + 8;
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(sbase + OFF_TMEMPTR, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + OFF_TMEMPTR);
  const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(full, UNIT_STRIDE_BYTES);
    for (uint32_t o = 0; o < UNIT_STRIDE_BYTES; o += 16384) bulk_g2s(sbase + OFF_H + o, wimg + o, 16384, full);
  }
  for (int j = 0; j < 4; ++j) {  // 32 channels at a time, like compute thread (r, g = j)
    float v[32];
    for (int i = 0; i < 32; ++i) v[i] = A[r * 128 + 32 * j + i];
    uint32_t h[16], l8[8], h8[8];
    if (mode == 0) {
      for (int c = 0; c < 4; ++c) split8_f8(v + 8 * c, h + 4 * c, l8 + 2 * c, h8 + 2 * c);
    } else {  // linear2's path: the H split from 128 h, as the linear1 epilogue does it
      for (int i = 0; i < 32; ++i) v[i] *= F8_XS;
      for (int c = 0; c < 4; ++c) split8_f8_h(v + 8 * c, h + 4 * c, l8 + 2 * c, h8 + 2 * c);
    }
    if (mode == 0) {
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(sgen + OFF_AX_HI + (j >> 1) * 16384 + sw128_chunk_off(r, (j & 1) * 4 + c)) =
            make_uint4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
      for (int c = 0; c < 2; ++c) {
        *reinterpret_cast<uint4*>(sgen + OFF_AX_LO + sw128_chunk_off(r, 2 * j + c)) =
            make_uint4(l8[4 * c], l8[4 * c + 1], l8[4 * c + 2], l8[4 * c + 3]);
        *reinterpret_cast<uint4*>(sgen + OFF_AX_LO + 16384 + sw128_chunk_off(r, 2 * j + c)) =
            make_uint4(h8[4 * c], h8[4 * c + 1], h8[4 * c + 2], h8[4 * c + 3]);
      }
    } else {
      tmem_st16(trow + TM_HT + 32 * j, h);  // the decoder's in-place layout: [fp16 16 | e4m3 lo 8 | e4m3 hi 8] per block
      tmem_st8(trow + TM_HT + 32 * j + 16, l8);
      tmem_st8(trow + TM_HT + 32 * j + 24, h8);
    }
  }
  if (mode == 0) fence_proxy_async_smem();
  else {
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_wait(full, 0);
    tc_fence_after();
    const uint32_t w = sbase + OFF_H;
    constexpr uint32_t ID = make_idesc_f16(128);
    if (mode == 0) {
      issue_part<1, 1, 8, 16384u, 16384u, ID>(tmem, sbase + OFF_AX_HI, sbase + OFF_AX_HI, w, true);
      for (uint32_t ks = 0; ks < 4; ++ks) {
        umma_f8(tmem, make_desc_sw128(sbase + OFF_AX_LO + 32 * ks), make_desc_sw128(w + UNIT_PART_BYTES + 32 * ks), ID, 1u);
        umma_f8(tmem, make_desc_sw128(sbase + OFF_AX_LO + 16384 + 32 * ks), make_desc_sw128(w + UNIT_PART_BYTES + 16384 + 32 * ks), ID, 1u);
      }
    } else {
      issue_part_ts<1, 1, 8, 16384u, ID, true>(tmem, tmem + TM_HT, tmem + TM_HT, w, true);
      for (uint32_t ks = 0; ks < 4; ++ks) {
        umma_f8_ts(tmem, tmem + TM_HT + 32 * ks + 16, make_desc_sw128(w + UNIT_PART_BYTES + 32 * ks), ID, 1u);
        umma_f8_ts(tmem, tmem + TM_HT + 32 * ks + 24, make_desc_sw128(w + UNIT_PART_BYTES + 16384 + 32 * ks), ID, 1u);
      }
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  for (int j = 0; j < 4; ++j) {
    float v[32];
    tmem_ld32(trow + 32 * j, v);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) D[r * 128 + 32 * j + c] = v[c] * F8_ACC_INV;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

// Build the operand images of the three attention layers from the fp32 [K][N] matrices already
// packed for the fp32 path (DecF32), in the order the MMA issuer consumes them.
int dectc_pack(s3d_model* m, cudaStream_t st) {
  const size_t total = (size_t)3 * UNITS_PER_LAYER * UNIT_STRIDE_BYTES;
  std::vector<uint8_t> img(total);
  auto fetch = [&](const ConvW& cw, std::vector<float>& h) -> int {
    h.resize((size_t)cw.kpad * cw.ncols);
    S3D_CUDA(cudaMemcpyAsync(h.data(), cw.w, h.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    S3D_CUDA(cudaStreamSynchronize(st));
    return S3D_OK;
  };
  // three images: bf16 hi/lo pairs (S3D_PREC_BF16X3, S3D_PREC_BF16), fp16 hi/lo pairs (S3D_PREC_FP16X3), and fp16 pairs for
  // the attention units + fp16 / fp8 / fp8 for the FFN units (S3D_PREC_FP16F8)
  for (int fmt = 0; fmt < 3; ++fmt) {
    const bool f16 = fmt >= 1, f8 = fmt == 2;
    for (int l = 0; l < 3; ++l) {
      const DecLayerF32& L = m->dec32.L[l];
      std::vector<float> win, wo, w1, w2;  // each [K][N]: W(n,k) = w[k*N + n]
      S3D_TRY(fetch(L.in_proj, win));
      S3D_TRY(fetch(L.out_proj, wo));
      S3D_TRY(fetch(L.lin1, w1));
      S3D_TRY(fetch(L.lin2, w2));
      uint8_t* dst = img.data() + (size_t)l * UNITS_PER_LAYER * UNIT_STRIDE_BYTES;
      int g = 0;
      for (int u : {1, 0, 2}) {  // issue order of the QKV projection: K, Q, V
        pack_unit([&](int n, int k) { return win[(size_t)k * 384 + 128 * u + n]; }, 128, 128, dst + (size_t)g * UNIT_STRIDE_BYTES, f16);
        ++g;
      }
      pack_unit([&](int n, int k) { return wo[(size_t)k * 128 + n]; }, 128, 128, dst + (size_t)g * UNIT_STRIDE_BYTES, f16);
      ++g;
      auto pack_w1 = [&](int c) {  // hidden units 128c .. +127 as output columns
        auto W = [&](int n, int k) { return w1[(size_t)k * 2048 + 128 * c + n]; };
        if (f8) pack_unit_f8(W, dst + (size_t)g * UNIT_STRIDE_BYTES);
        else pack_unit(W, 128, 128, dst + (size_t)g * UNIT_STRIDE_BYTES, f16);
        ++g;
      };
      auto pack_w2 = [&](int c) {  // hidden units 128c .. +127 as the contraction index
        auto W = [&](int n, int k) { return w2[(size_t)(128 * c + k) * 128 + n]; };
        if (f8) pack_unit_f8(W, dst + (size_t)g * UNIT_STRIDE_BYTES);
        else pack_unit(W, 128, 128, dst + (size_t)g * UNIT_STRIDE_BYTES, f16);
        ++g;
      };
      // consumption order of the FFN pipeline: W1_0 .. W1_2, then (W2_c, W1_{c+3}) ...
      for (int c = 0; c < NDBUF; ++c) pack_w1(c);
      for (int c = 0; c < NCHUNK; ++c) {
        pack_w2(c);
        if (c + NDBUF < NCHUNK) pack_w1(c + NDBUF);
      }
    }
    void* d = nullptr;
    S3D_CUDA(cudaMalloc(&d, total));
    m->allocs.push_back(d);
    S3D_CUDA(cudaMemcpyAsync(d, img.data(), total, cudaMemcpyHostToDevice, st));
    S3D_CUDA(cudaStreamSynchronize(st));
    if (f8) m->dectc.wimg_f8 = static_cast<uint8_t*>(d);
    else if (f16) m->dectc.wimg_h = static_cast<__half*>(d);
    else m->dectc.wimg = static_cast<__nv_bfloat16*>(d);
  }
  m->dectc.wimg_elems = total / 2;
  // per-layer fp32 vector blocks (biases, LayerNorm affines) in the order of the V_* offsets
  std::vector<float> vecs((size_t)3 * VEC_FLOATS, 0.f);
  auto get = [&](const float* src, int n, float* dst) -> int {
    S3D_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    return S3D_OK;
  };
  for (int l = 0; l < 3; ++l) {
    const DecLayerF32& L = m->dec32.L[l];
    float* v = vecs.data() + (size_t)l * VEC_FLOATS;
    S3D_TRY(get(L.in_proj.shift, 384, v + V_BIN));
    if (l < 2) S3D_TRY(get(m->dec32.L[l + 1].out_proj.shift, 128, v + V_BONEXT));
    S3D_TRY(get(L.n1_w, 128, v + V_LN1W));
    S3D_TRY(get(L.n1_b, 128, v + V_LN1B));
    S3D_TRY(get(L.lin1.shift, 2048, v + V_B1));
    S3D_TRY(get(L.lin2.shift, 128, v + V_B2));
    S3D_TRY(get(L.n2_w, 128, v + V_LN2W));
    S3D_TRY(get(L.n2_b, 128, v + V_LN2B));
  }
  S3D_CUDA(cudaStreamSynchronize(st));
  void* dv = nullptr;
  S3D_CUDA(cudaMalloc(&dv, vecs.size() * sizeof(float)));
  m->allocs.push_back(dv);
  S3D_CUDA(cudaMemcpyAsync(dv, vecs.data(), vecs.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  S3D_CUDA(cudaStreamSynchronize(st));
  m->dectc.vec = static_cast<float*>(dv);
  return S3D_OK;
}

static int g_tc_dbg = 0;  // s3d_debug_set_decoder_flags: timing experiments (bit 0: skip the weight copies -> garbage results)
void decoder_tc_set_debug(int flags) { g_tc_dbg = flags; }

bool decoder_tc_supported(const s3d_model* m) { return m && m->K >= 1 && m->K <= 12 && m->dectc.wimg != nullptr; }

int debug_profile(long long* out32, int reset) {
  unsigned long long h[32];
  S3D_CUDA(cudaMemcpyFromSymbol(h, g_prof, sizeof(h)));
  if (out32)
    for (int i = 0; i < 32; ++i) out32[i] = (long long)h[i];
  if (reset) {
    std::memset(h, 0, sizeof(h));
    S3D_CUDA(cudaMemcpyToSymbol(g_prof, h, sizeof(h)));
  }
  return S3D_OK;
}

size_t decoder_tc_workspace_bytes(int64_t) {
  // per CTA (up to 256): tail scratch [126][256] + double-buffered token scratch 2 x [128][128], fp32
  return (size_t)256 * (TAIL_SLOTS * TILE_Q * 256 + 2 * 128 * 128) * sizeof(float);
}

int decoder_tc(const s3d_model* m, const float* planes, int S, const QueryCtx& q, int64_t n, float out_scale, float* out,
               int precision, void* ws, size_t ws_bytes, cudaStream_t st, const int* n_dev) {
  if (!decoder_tc_supported(m)) {
    set_error("decoder: tensor-core modes need 1 <= n_slices <= 12");
    return S3D_ERR_UNSUPPORTED;
  }
  if (n <= 0) return S3D_OK;
  TcParams p{};
  p.wimg = reinterpret_cast<const uint8_t*>(precision == S3D_PREC_FP16F8   ? (const void*)m->dectc.wimg_f8
                                            : precision == S3D_PREC_FP16X3 ? (const void*)m->dectc.wimg_h
                                                                           : (const void*)m->dectc.wimg);
  p.vecs = m->dectc.vec;
  p.planes = planes;
  p.S = S;
  p.q = q;
  p.n = n;
  p.dbg = g_tc_dbg;
  p.K = m->K;
  p.n_dev = n_dev;
  p.out_scale = out_scale;
  p.out = out;
  const DecF32& d = m->dec32;
  p.fcp_wt = d.fcp_wt; p.fcp_b = d.fcp_b; p.fcs_b = d.fcs_b; p.fco_w = d.fco_w; p.fco_b = d.fco_b;
  p.b_o0 = d.L[0].out_proj.shift;
  p.num_tiles = (n + TILE_Q - 1) / TILE_Q;
  int dev = 0, sms = 148;
  S3D_CUDA(cudaGetDevice(&dev));
  S3D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sms > 256) sms = 256;
  if (ws == nullptr || ws_bytes < decoder_tc_workspace_bytes(n)) {
    set_error("decoder: workspace too small");
    return S3D_ERR_WORKSPACE;
  }
  p.scratch = static_cast<float*>(ws);
  auto launch = [&](auto kern, unsigned g, int cluster) -> int {
    S3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(g);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    S3D_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return S3D_OK;
  };
  {  // CTA pairs: clusters of 2, one pair per TPC
    const long long pairs = (p.num_tiles + 1) / 2;
    const unsigned g2 = 2u * (unsigned)((pairs < sms / 2 && !n_dev) ? pairs : sms / 2);
    if (precision == S3D_PREC_FP16F8) S3D_TRY(launch(decoder_tc_kernel<3, 2, true, true>, g2, 2));
    else if (precision == S3D_PREC_FP16X3) S3D_TRY(launch(decoder_tc_kernel<3, 2, true, false>, g2, 2));
    else if (precision == S3D_PREC_BF16X3) S3D_TRY(launch(decoder_tc_kernel<3, 2, false, false>, g2, 2));
    else S3D_TRY(launch(decoder_tc_kernel<1, 2, false, false>, g2, 2));
  }
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

// Self-test of the UMMA plumbing (descriptors, swizzle, bulk copy, TMEM load): see include/slice3d_b200.h.
int umma_selftest(int mode, int passes, const float* a_dev, const float* w_dev, float* d_dev, cudaStream_t st) {
  const bool f16 = passes == 4;  // passes = 4: the three-pass schedule with fp16 hi/lo pairs (S3D_PREC_FP16X3)
  const bool f8 = passes == 5;   // passes = 5: fp16 leading term + two fp8 cross terms (S3D_PREC_FP16F8's FFN units)
  if ((mode != 0 && mode != 1) || (passes != 1 && passes != 3 && passes != 4 && passes != 5) || !a_dev || !w_dev || !d_dev) {
    set_error("selftest: bad argument");
    return S3D_ERR_BAD_ARG;
  }
  const int N = 128, K = 128;
  std::vector<float> w((size_t)N * K);
  S3D_CUDA(cudaMemcpyAsync(w.data(), w_dev, w.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
  S3D_CUDA(cudaStreamSynchronize(st));
  std::vector<uint8_t> img(UNIT_STRIDE_BYTES);
  if (f8) pack_unit_f8([&](int n, int k) { return w[(size_t)n * K + k]; }, img.data());
  else pack_unit([&](int n, int k) { return w[(size_t)n * K + k]; }, N, K, img.data(), f16);
  void* d = nullptr;
  S3D_CUDA(cudaMalloc(&d, UNIT_STRIDE_BYTES));
  S3D_CUDA(cudaMemcpyAsync(d, img.data(), UNIT_STRIDE_BYTES, cudaMemcpyHostToDevice, st));
  if (f8) {
    cudaFuncSetAttribute(umma_selftest_f8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    umma_selftest_f8_kernel<<<1, 128, SMEM_BYTES, st>>>(a_dev, static_cast<const uint8_t*>(d), mode, d_dev);
  } else if (f16) {
    cudaFuncSetAttribute(umma_selftest_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    umma_selftest_kernel<3, true><<<1, 128, SMEM_BYTES, st>>>(a_dev, static_cast<const uint8_t*>(d), mode, d_dev);
  } else if (passes == 3) {
    cudaFuncSetAttribute(umma_selftest_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    umma_selftest_kernel<3, false><<<1, 128, SMEM_BYTES, st>>>(a_dev, static_cast<const uint8_t*>(d), mode, d_dev);
  } else {
    cudaFuncSetAttribute(umma_selftest_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    umma_selftest_kernel<1, false><<<1, 128, SMEM_BYTES, st>>>(a_dev, static_cast<const uint8_t*>(d), mode, d_dev);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) {
    set_error(std::string("selftest kernel: ") + cudaGetErrorString(e));
    return S3D_ERR_CUDA;
  }
  return S3D_OK;
}

}  // namespace s3d
