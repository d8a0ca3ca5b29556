// Fused tensor-core decoder for sm_100a (S3D_PREC_BF16X3 / S3D_PREC_BF16).
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
// Activations are written by the epilogue warps straight into the canonical K-major/128B-swizzled
// shared-memory operand layout; weights are pre-swizzled "operand images" in global memory streamed
// through a 3-slot ring with bulk async copies (TMA engine) completing on mbarriers.  The residual
// stream never leaves TMEM: the accumulator of out-proj / FFN2 is pre-loaded with x + bias, so the
// MMA result is already residual + bias + contraction and LayerNorm runs in place, one thread per row.
//
// Precision: BF16X3 splits both operands into bf16 hi + lo and issues hi*hi + lo*hi + hi*lo
// (3 passes, ~16 mantissa bits, max-abs error ~2e-5 on sdf_pred); BF16 issues hi*hi only.
//
// Warp roles (192 threads): warps 0-3 = epilogue/compute (thread r owns tile row r = TMEM lane r),
// warp 4 lane 0 = weight producer, warp 5 lane 0 = MMA issuer.
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace s3d {

namespace {

using namespace ptx;

constexpr int TILE_Q = 9;    // queries per tile
constexpr int NTOK = 13;     // tokens per query (K = 12 slices + the query token)
constexpr int NSLOT = 3;     // weight ring depth
constexpr int UNITS_PER_LAYER = 72;  // 6 (in_proj) + 2 (out_proj) + 32 (linear1) + 32 (linear2)
constexpr int UNIT_PART_BYTES = 16384;  // one precision part (hi or lo) of a unit: 64x128 or 128x64 bf16
constexpr int UNIT_STRIDE_BYTES = 2 * UNIT_PART_BYTES;  // hi then lo in global memory
constexpr int NCHUNK = 32;   // FFN hidden chunks of 64

// shared memory map (bytes from the 1024-aligned base)
constexpr uint32_t OFF_AX_HI = 0;                 // [2 k-blocks][128 rows][64] bf16 = 32 KB
constexpr uint32_t OFF_AX_LO = 32768;             // 32 KB
constexpr uint32_t OFF_HKV = 65536;               // H chunk operand (hi 16 KB, lo 16 KB) | K/V staging | gather scratch
constexpr uint32_t HKV_BYTES = 36864;             // 2 x [128][36] fp32
constexpr uint32_t OFF_RING = OFF_HKV + HKV_BYTES;  // 102400, 1024-aligned
constexpr uint32_t OFF_BAR = OFF_RING + NSLOT * UNIT_STRIDE_BYTES;  // 200704
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 1024;               // + alignment slack

// barrier indices (8 bytes each at OFF_BAR)
enum { B_AREADY = 0, B_DDONE, B_D1READY0, B_D1READY1, B_D1FREE0, B_D1FREE1, B_HREADY, B_HFREE, B_FULL0, B_FULL1, B_FULL2,
       B_EMPTY0, B_EMPTY1, B_EMPTY2, B_COUNT };
constexpr uint32_t OFF_TMEMPTR = OFF_BAR + 8 * B_COUNT;

// TMEM columns
constexpr uint32_t TM_R = 0;      // residual / out-proj / FFN2 accumulator, 128 columns
constexpr uint32_t TM_S = 128;    // QKV accumulators (384 columns) | FFN1 chunk accumulators (2 x 64)

struct TcParams {
  const uint8_t* wimg;  // [3 layers][72 units][hi 16 KB | lo 16 KB]
  const float* planes;
  int S;
  QueryCtx q;
  long long n;
  float out_scale;
  float* out;
  const float *fcp_wt, *fcp_b, *fcs_b, *fco_w, *fco_b;
  const float *b_in[3], *b_o[3], *ln1w[3], *ln1b[3], *b1[3], *b2[3], *ln2w[3], *ln2b[3];
  long long num_tiles;
};

// ---- operand writes ------------------------------------------------------------------------
// 8 consecutive k values of row r -> one 16-byte chunk of the hi tile (and of the lo tile).
template <int NPASS>
__device__ __forceinline__ void store_chunk(uint8_t* tile_hi, uint8_t* tile_lo, int r, int kc, const float* v) {
  __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_bf16(v[i], h[i], l[i]);
  const uint32_t off = sw128_chunk_off(r, kc);
  *reinterpret_cast<uint4*>(tile_hi + off) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
  if (NPASS == 3)
    *reinterpret_cast<uint4*>(tile_lo + off) = make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
}

// ---- MMA issue -----------------------------------------------------------------------------
// One weight unit against one activation operand: KS k-steps of 16, NPASS passes.
//   a_hi/a_lo, b_hi/b_lo: shared addresses of the operand tiles; *_kb: byte stride between 64-wide k-blocks.
template <int NPASS>
__device__ __forceinline__ void issue_unit(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_kb, uint32_t b_hi,
                                           uint32_t b_lo, uint32_t b_kb, int KS, uint32_t idesc, bool fresh) {
  uint32_t acc = fresh ? 0u : 1u;
#pragma unroll 1
  for (int pass = 0; pass < NPASS; ++pass) {
    const uint32_t a = (pass == 1) ? a_lo : a_hi;
    const uint32_t b = (pass == 2) ? b_lo : b_hi;
    const uint64_t ad = make_desc_sw128(a), bd = make_desc_sw128(b);
#pragma unroll 1
    for (int ks = 0; ks < KS; ++ks) {
      const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
      umma_bf16(d_tmem, ad + ((kb * a_kb + kin) >> 4), bd + ((kb * b_kb + kin) >> 4), idesc, acc);
      acc = 1u;
    }
  }
}

template <int NPASS>
__global__ void __launch_bounds__(192, 1) decoder_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };

  if (threadIdx.x == 0) {
    mbar_init(bar(B_AREADY), 128);
    mbar_init(bar(B_DDONE), 1);
    mbar_init(bar(B_D1READY0), 1);
    mbar_init(bar(B_D1READY1), 1);
    mbar_init(bar(B_D1FREE0), 128);
    mbar_init(bar(B_D1FREE1), 128);
    mbar_init(bar(B_HREADY), 128);
    mbar_init(bar(B_HFREE), 1);
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(bar(B_FULL0 + s), 1);
      mbar_init(bar(B_EMPTY0 + s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(sbase + OFF_TMEMPTR, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + OFF_TMEMPTR);

  constexpr uint32_t COPY_BYTES = (NPASS == 3) ? UNIT_STRIDE_BYTES : UNIT_PART_BYTES;

  if (warp == 4) {
    // ===================================================================== weight producer
    if (lane == 0) {
      uint32_t ph_empty[NSLOT];
      for (int s = 0; s < NSLOT; ++s) ph_empty[s] = 1;
      int slot = 0;
      for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int g = 0; g < 3 * UNITS_PER_LAYER; ++g) {
          mbar_wait(bar(B_EMPTY0 + slot), ph_empty[slot]);
          ph_empty[slot] ^= 1;
          mbar_arrive_expect_tx(bar(B_FULL0 + slot), COPY_BYTES);
          bulk_g2s(sbase + OFF_RING + slot * UNIT_STRIDE_BYTES, p.wimg + (size_t)g * UNIT_STRIDE_BYTES, COPY_BYTES,
                   bar(B_FULL0 + slot));
          slot = (slot + 1 == NSLOT) ? 0 : slot + 1;
        }
      }
    }
  } else if (warp == 5) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      uint32_t ph_a = 0, ph_full[NSLOT], ph_d1free[2] = {1, 1}, ph_hr = 0;
      for (int s = 0; s < NSLOT; ++s) ph_full[s] = 0;
      int slot = 0;
      const uint32_t ax_hi = sbase + OFF_AX_HI, ax_lo = sbase + OFF_AX_LO;
      const uint32_t h_hi = sbase + OFF_HKV, h_lo = sbase + OFF_HKV + UNIT_PART_BYTES;
      constexpr uint32_t ID64 = make_idesc_bf16(64), ID128 = make_idesc_bf16(128);
      auto wait_full = [&]() -> uint32_t {
        mbar_wait(bar(B_FULL0 + slot), ph_full[slot]);
        ph_full[slot] ^= 1;
        return sbase + OFF_RING + slot * UNIT_STRIDE_BYTES;
      };
      auto release = [&]() {
        umma_commit(bar(B_EMPTY0 + slot));
        slot = (slot + 1 == NSLOT) ? 0 : slot + 1;
      };
      // unit of 64 output columns over K = 128 (in_proj / out_proj / linear1), A = AX
      auto unit_n64 = [&](uint32_t d_col, bool fresh) {
        const uint32_t w = wait_full();
        tc_fence_after();
        issue_unit<NPASS>(tmem + d_col, ax_hi, ax_lo, 16384u, w, w + UNIT_PART_BYTES, 8192u, 8, ID64, fresh);
        release();
      };
      for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int layer = 0; layer < 3; ++layer) {
          // ---- QKV projection: S[:, 0:384] = X . Win^T
          mbar_wait(bar(B_AREADY), ph_a);
          ph_a ^= 1;
          tc_fence_after();
          for (int u = 0; u < 6; ++u) unit_n64(TM_S + 64 * u, true);
          umma_commit(bar(B_DDONE));
          // ---- out-proj: R += O . Wo^T   (R pre-loaded with x + b_o)
          mbar_wait(bar(B_AREADY), ph_a);
          ph_a ^= 1;
          tc_fence_after();
          for (int u = 0; u < 2; ++u) unit_n64(TM_R + 64 * u, false);
          umma_commit(bar(B_DDONE));
          // ---- FFN: D1[c] = X' . W1_c^T (N=64) ; R += relu(D1[c] + b1) . W2_c^T (N=128, K=64)
          mbar_wait(bar(B_AREADY), ph_a);
          ph_a ^= 1;
          tc_fence_after();
          auto issue1 = [&](int c) {
            mbar_wait(bar(B_D1FREE0 + (c & 1)), ph_d1free[c & 1]);
            ph_d1free[c & 1] ^= 1;
            tc_fence_after();
            unit_n64(TM_S + 64 * (c & 1), true);
            umma_commit(bar(B_D1READY0 + (c & 1)));
          };
          auto issue2 = [&](int c) {
            const uint32_t w = wait_full();
            mbar_wait(bar(B_HREADY), ph_hr);
            ph_hr ^= 1;
            tc_fence_after();
            issue_unit<NPASS>(tmem + TM_R, h_hi, h_lo, 0u, w, w + UNIT_PART_BYTES, 0u, 4, ID128, false);
            release();
            umma_commit(bar(B_HFREE));
          };
          issue1(0);
#pragma unroll 1
          for (int c = 0; c < NCHUNK; ++c) {
            if (c + 1 < NCHUNK) issue1(c + 1);
            issue2(c);
          }
          umma_commit(bar(B_DDONE));
        }
      }
    }
  } else {
    // ===================================================================== epilogue / compute warps
    const int r = threadIdx.x;  // tile row == TMEM lane
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    uint8_t* ax_hi = sgen + OFF_AX_HI;
    uint8_t* ax_lo = sgen + OFF_AX_LO;
    uint8_t* h_hi = sgen + OFF_HKV;
    uint8_t* h_lo = sgen + OFF_HKV + UNIT_PART_BYTES;
    uint32_t ph_d = 0, ph_d1r[2] = {0, 0}, ph_hf = 1;
    const int qi = r / NTOK, tk = r - qi * NTOK;
    const unsigned FULL = 0xffffffffu;

    for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const long long q_idx = tile * TILE_Q + qi;
      const bool valid = (qi < TILE_Q) && (q_idx < p.n);
      // ------------------------------------------------------------------ token build
      float px = 0.f, py = 0.f, pz = 0.f, gu = 0.f, gv = 0.f;
      if (valid) load_query(p.q, q_idx, px, py, pz, gu, gv);
      {
        float* scr = reinterpret_cast<float*>(sgen + OFF_HKV) + warp * (32 * 68);
        const int qtr = lane >> 3, l8 = lane & 7;
        const float* b_o0 = p.b_o[0];
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
#pragma unroll 1
          for (int rg = 0; rg < 8; ++rg) {
            const int rl = rg * 4 + qtr;  // row (within this warp) gathered by this quarter-warp
            const float ru = __shfl_sync(FULL, gu, rl), rv = __shfl_sync(FULL, gv, rl);
            const int rvalid = __shfl_sync(FULL, valid ? 1 : 0, rl);
            const int rr = warp * 32 + rl;
            const int rt = rr % NTOK;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            if (rvalid && rt > 0) {
              const int ch = half * 64 + l8 * 4;
              a0 = __ldg(reinterpret_cast<const float4*>(p.fcs_b + ch));
              a1 = __ldg(reinterpret_cast<const float4*>(p.fcs_b + ch + 32));
              size_t off = 0;
#pragma unroll 1
              for (int s = 0; s < 5; ++s) {
                const int R = plane_res(p.S, s);
                const Taps t = make_taps(ru, rv, R);
                const float* P = p.planes + off + (size_t)(rt - 1) * R * R * 128 + ch;
                const float4 c00 = __ldg(reinterpret_cast<const float4*>(P + (size_t)t.o00 * 128));
                const float4 c01 = __ldg(reinterpret_cast<const float4*>(P + (size_t)t.o01 * 128));
                const float4 c10 = __ldg(reinterpret_cast<const float4*>(P + (size_t)t.o10 * 128));
                const float4 c11 = __ldg(reinterpret_cast<const float4*>(P + (size_t)t.o11 * 128));
                const float4 d00 = __ldg(reinterpret_cast<const float4*>(P + (size_t)t.o00 * 128 + 32));
                const float4 d01 = __ldg(reinterpret_cast<const float4*>(P + (size_t)t.o01 * 128 + 32));
                const float4 d10 = __ldg(reinterpret_cast<const float4*>(P + (size_t)t.o10 * 128 + 32));
                const float4 d11 = __ldg(reinterpret_cast<const float4*>(P + (size_t)t.o11 * 128 + 32));
                a0.x += c00.x * t.w00 + c01.x * t.w01 + c10.x * t.w10 + c11.x * t.w11;
                a0.y += c00.y * t.w00 + c01.y * t.w01 + c10.y * t.w10 + c11.y * t.w11;
                a0.z += c00.z * t.w00 + c01.z * t.w01 + c10.z * t.w10 + c11.z * t.w11;
                a0.w += c00.w * t.w00 + c01.w * t.w01 + c10.w * t.w10 + c11.w * t.w11;
                a1.x += d00.x * t.w00 + d01.x * t.w01 + d10.x * t.w10 + d11.x * t.w11;
                a1.y += d00.y * t.w00 + d01.y * t.w01 + d10.y * t.w10 + d11.y * t.w11;
                a1.z += d00.z * t.w00 + d01.z * t.w01 + d10.z * t.w10 + d11.z * t.w11;
                a1.w += d00.w * t.w00 + d01.w * t.w01 + d10.w * t.w10 + d11.w * t.w11;
                off += (size_t)12 * R * R * 128;
              }
            }
            *reinterpret_cast<float4*>(scr + rl * 68 + l8 * 4) = a0;
            *reinterpret_cast<float4*>(scr + rl * 68 + 32 + l8 * 4) = a1;
          }
          __syncwarp();
          float v[64];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 t4 = *reinterpret_cast<const float4*>(scr + lane * 68 + i * 4);
            v[4 * i] = t4.x; v[4 * i + 1] = t4.y; v[4 * i + 2] = t4.z; v[4 * i + 3] = t4.w;
          }
          if (valid && tk == 0) {  // query token: fc_p(q) (models.py:79)
#pragma unroll
            for (int c = 0; c < 64; ++c) {
              const int ch = half * 64 + c;
              v[c] = __ldg(p.fcp_b + ch) + px * __ldg(p.fcp_wt + ch) + py * __ldg(p.fcp_wt + 128 + ch) +
                     pz * __ldg(p.fcp_wt + 256 + ch);
            }
          }
#pragma unroll
          for (int kc = 0; kc < 8; ++kc)
            store_chunk<NPASS>(ax_hi + half * 16384, ax_lo + half * 16384, r, kc, v + 8 * kc);
#pragma unroll
          for (int c = 0; c < 64; ++c) v[c] += __ldg(b_o0 + half * 64 + c);
          tmem_st32(trow + TM_R + half * 64, v);
          tmem_st32(trow + TM_R + half * 64 + 32, v + 32);
          __syncwarp();
        }
        tmem_st_wait();
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(bar(B_AREADY));
      }

#pragma unroll 1
      for (int layer = 0; layer < 3; ++layer) {
        // -------------------------------------------------------------- attention
        mbar_wait(bar(B_DDONE), ph_d);
        ph_d ^= 1;
        tc_fence_after();
        {
          float* Ks = reinterpret_cast<float*>(sgen + OFF_HKV);
          float* Vs = Ks + 128 * 36;
          const float* b_in = p.b_in[layer];
#pragma unroll 1
          for (int h = 0; h < 4; ++h) {
            {
              float kk[32], vv[32];
              tmem_ld32(trow + TM_S + 128 + 32 * h, kk);
              tmem_ld32(trow + TM_S + 256 + 32 * h, vv);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; c += 4) {
                *reinterpret_cast<float4*>(Ks + r * 36 + c) =
                    make_float4(kk[c] + __ldg(b_in + 128 + 32 * h + c), kk[c + 1] + __ldg(b_in + 129 + 32 * h + c),
                                kk[c + 2] + __ldg(b_in + 130 + 32 * h + c), kk[c + 3] + __ldg(b_in + 131 + 32 * h + c));
                *reinterpret_cast<float4*>(Vs + r * 36 + c) =
                    make_float4(vv[c] + __ldg(b_in + 256 + 32 * h + c), vv[c + 1] + __ldg(b_in + 257 + 32 * h + c),
                                vv[c + 2] + __ldg(b_in + 258 + 32 * h + c), vv[c + 3] + __ldg(b_in + 259 + 32 * h + c));
              }
            }
            named_bar_sync(1, 128);
            float qq[32], o[32];
            tmem_ld32(trow + TM_S + 32 * h, qq);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              qq[c] = (qq[c] + __ldg(b_in + 32 * h + c)) * 0.17677669529663687f;
              o[c] = 0.f;
            }
            if (valid) {
              const float* kb = Ks + qi * NTOK * 36;
              const float* vb = Vs + qi * NTOK * 36;
              float sc[NTOK];
              float mx = -3.0e38f;
#pragma unroll
              for (int j = 0; j < NTOK; ++j) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                  const float4 k4 = *reinterpret_cast<const float4*>(kb + j * 36 + c);
                  s = fmaf(qq[c], k4.x, s);
                  s = fmaf(qq[c + 1], k4.y, s);
                  s = fmaf(qq[c + 2], k4.z, s);
                  s = fmaf(qq[c + 3], k4.w, s);
                }
                sc[j] = s;
                mx = fmaxf(mx, s);
              }
              float sum = 0.f;
#pragma unroll
              for (int j = 0; j < NTOK; ++j) {
                sc[j] = __expf(sc[j] - mx);
                sum += sc[j];
              }
              const float inv = 1.f / sum;
#pragma unroll
              for (int j = 0; j < NTOK; ++j) {
                const float pj = sc[j] * inv;
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                  const float4 v4 = *reinterpret_cast<const float4*>(vb + j * 36 + c);
                  o[c] = fmaf(pj, v4.x, o[c]);
                  o[c + 1] = fmaf(pj, v4.y, o[c + 1]);
                  o[c + 2] = fmaf(pj, v4.z, o[c + 2]);
                  o[c + 3] = fmaf(pj, v4.w, o[c + 3]);
                }
              }
            }
            // O[:, 32h:32h+32] -> operand A (k-block h/2, chunks 4*(h&1)..+3)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
              store_chunk<NPASS>(ax_hi + (h >> 1) * 16384, ax_lo + (h >> 1) * 16384, r, (h & 1) * 4 + cc, o + 8 * cc);
            named_bar_sync(1, 128);  // everyone done with Ks/Vs before the next head overwrites them
          }
          fence_proxy_async_smem();
          mbar_arrive(bar(B_AREADY));
        }
        // -------------------------------------------------------------- residual + LayerNorm 1 (in place in TMEM)
        mbar_wait(bar(B_DDONE), ph_d);
        ph_d ^= 1;
        tc_fence_after();
        {
          float v[128];
#pragma unroll
          for (int j = 0; j < 4; ++j) tmem_ld32(trow + TM_R + 32 * j, v + 32 * j);
          tmem_ld_wait();
          float s = 0.f;
#pragma unroll
          for (int c = 0; c < 128; ++c) s += v[c];
          const float mean = s * (1.f / 128.f);
          float d2 = 0.f;
#pragma unroll
          for (int c = 0; c < 128; ++c) {
            v[c] -= mean;
            d2 = fmaf(v[c], v[c], d2);
          }
          const float rstd = rsqrtf(d2 * (1.f / 128.f) + 1e-5f);
          const float* w = p.ln1w[layer];
          const float* b = p.ln1b[layer];
#pragma unroll
          for (int c = 0; c < 128; ++c) v[c] = fmaf(v[c] * rstd, __ldg(w + c), __ldg(b + c));
#pragma unroll
          for (int kc = 0; kc < 16; ++kc)
            store_chunk<NPASS>(ax_hi + (kc >> 3) * 16384, ax_lo + (kc >> 3) * 16384, r, kc & 7, v + 8 * kc);
          const float* b2 = p.b2[layer];
#pragma unroll
          for (int c = 0; c < 128; ++c) v[c] += __ldg(b2 + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) tmem_st32(trow + TM_R + 32 * j, v + 32 * j);
          tmem_st_wait();
          tc_fence_before();
          fence_proxy_async_smem();
          mbar_arrive(bar(B_AREADY));
        }
        // -------------------------------------------------------------- FFN hidden chunks
        {
          const float* b1 = p.b1[layer];
#pragma unroll 1
          for (int c = 0; c < NCHUNK; ++c) {
            const int bsel = c & 1;
            mbar_wait(bar(B_D1READY0 + bsel), ph_d1r[bsel]);
            ph_d1r[bsel] ^= 1;
            tc_fence_after();
            float d[64];
            tmem_ld32(trow + TM_S + 64 * bsel, d);
            tmem_ld32(trow + TM_S + 64 * bsel + 32, d + 32);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(bar(B_D1FREE0 + bsel));
#pragma unroll
            for (int j = 0; j < 64; ++j) d[j] = fmaxf(d[j] + __ldg(b1 + c * 64 + j), 0.f);
            mbar_wait(bar(B_HFREE), ph_hf);
            ph_hf ^= 1;
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) store_chunk<NPASS>(h_hi, h_lo, r, kc, d + 8 * kc);
            fence_proxy_async_smem();
            mbar_arrive(bar(B_HREADY));
          }
        }
        // -------------------------------------------------------------- residual + LayerNorm 2
        mbar_wait(bar(B_DDONE), ph_d);
        ph_d ^= 1;
        tc_fence_after();
        {
          float v[128];
#pragma unroll
          for (int j = 0; j < 4; ++j) tmem_ld32(trow + TM_R + 32 * j, v + 32 * j);
          tmem_ld_wait();
          float s = 0.f;
#pragma unroll
          for (int c = 0; c < 128; ++c) s += v[c];
          const float mean = s * (1.f / 128.f);
          float d2 = 0.f;
#pragma unroll
          for (int c = 0; c < 128; ++c) {
            v[c] -= mean;
            d2 = fmaf(v[c], v[c], d2);
          }
          const float rstd = rsqrtf(d2 * (1.f / 128.f) + 1e-5f);
          const float* w = p.ln2w[layer];
          const float* b = p.ln2b[layer];
#pragma unroll
          for (int c = 0; c < 128; ++c) v[c] = fmaf(v[c] * rstd, __ldg(w + c), __ldg(b + c));
          if (layer < 2) {
#pragma unroll
            for (int kc = 0; kc < 16; ++kc)
              store_chunk<NPASS>(ax_hi + (kc >> 3) * 16384, ax_lo + (kc >> 3) * 16384, r, kc & 7, v + 8 * kc);
            const float* bo = p.b_o[layer + 1];
#pragma unroll
            for (int c = 0; c < 128; ++c) v[c] += __ldg(bo + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) tmem_st32(trow + TM_R + 32 * j, v + 32 * j);
            tmem_st_wait();
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(bar(B_AREADY));
          } else if (valid && tk == 0) {  // fc_out on token 0 (models.py:83-84)
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 128; ++c) acc = fmaf(v[c], __ldg(p.fco_w + c), acc);
            p.out[q_idx] = p.out_scale * (acc + __ldg(p.fco_b));
          }
        }
      }
      // all warps must be done with TMEM R / the scratch before the next tile's token build
      named_bar_sync(1, 128);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- self-test: one 128-row UMMA tile against one weight unit ------------------------------------
// mode 0: D[128][64]  = A[128][128] . W[64][128]^T   (unit shape of in_proj / out_proj / linear1)
// mode 1: D[128][128] = A[128][64]  . W[128][64]^T   (unit shape of linear2)
template <int NPASS>
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A, const uint8_t* wimg, int mode,
                                                               float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const int warp = threadIdx.x >> 5;
  const int r = threadIdx.x;
  const int K = mode == 0 ? 128 : 64, N = mode == 0 ? 64 : 128;
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
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(full, UNIT_STRIDE_BYTES);
    bulk_g2s(sbase + OFF_RING, wimg, UNIT_STRIDE_BYTES, full);
  }
  for (int kc = 0; kc < K / 8; ++kc) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = A[r * K + kc * 8 + i];
    store_chunk<NPASS>(sgen + OFF_AX_HI + (kc >> 3) * 16384, sgen + OFF_AX_LO + (kc >> 3) * 16384, r, kc & 7, v);
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_wait(full, 0);
    tc_fence_after();
    const uint32_t w = sbase + OFF_RING;
    if (mode == 0)
      issue_unit<NPASS>(tmem, sbase + OFF_AX_HI, sbase + OFF_AX_LO, 16384u, w, w + UNIT_PART_BYTES, 8192u, 8,
                        make_idesc_bf16(64), true);
    else
      issue_unit<NPASS>(tmem, sbase + OFF_AX_HI, sbase + OFF_AX_LO, 0u, w, w + UNIT_PART_BYTES, 0u, 4,
                        make_idesc_bf16(128), true);
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int j = 0; j < N / 32; ++j) {
    float v[32];
    tmem_ld32(trow + 32 * j, v);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) D[r * N + 32 * j + c] = v[c];
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
template <class F>
void pack_unit(F W, int NU, int KU, uint8_t* dst) {
  uint16_t* hi = reinterpret_cast<uint16_t*>(dst);
  uint16_t* lo = reinterpret_cast<uint16_t*>(dst + UNIT_PART_BYTES);
  for (int kb = 0; kb < KU / 64; ++kb)
    for (int n = 0; n < NU; ++n)
      for (int k = 0; k < 64; ++k) {
        const float w = W(n, kb * 64 + k);
        const uint16_t h = bf16_bits(w);
        const uint16_t l = bf16_bits(w - bf16_val(h));
        const size_t off = (size_t)kb * NU * 128 + sw128_chunk_off(n, k >> 3) + (k & 7) * 2;
        hi[off / 2] = h;
        lo[off / 2] = l;
      }
}

}  // namespace

// Build the operand images of the three attention layers from the fp32 [K][N] matrices already
// packed for the fp32 path (DecF32), in the order the MMA issuer consumes them.
int dectc_pack(s3d_model* m, cudaStream_t st) {
  if (m->K != 12) return S3D_OK;  // the tensor-core tile layout is specialised for 13 tokens; fp32 path serves other K
  const size_t total = (size_t)3 * UNITS_PER_LAYER * UNIT_STRIDE_BYTES;
  std::vector<uint8_t> img(total);
  auto fetch = [&](const ConvW& cw, std::vector<float>& h) -> int {
    h.resize((size_t)cw.kpad * cw.ncols);
    S3D_CUDA(cudaMemcpyAsync(h.data(), cw.w, h.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    S3D_CUDA(cudaStreamSynchronize(st));
    return S3D_OK;
  };
  for (int l = 0; l < 3; ++l) {
    const DecLayerF32& L = m->dec32.L[l];
    std::vector<float> win, wo, w1, w2;  // each [K][N]: W(n,k) = w[k*N + n]
    S3D_TRY(fetch(L.in_proj, win));
    S3D_TRY(fetch(L.out_proj, wo));
    S3D_TRY(fetch(L.lin1, w1));
    S3D_TRY(fetch(L.lin2, w2));
    uint8_t* dst = img.data() + (size_t)l * UNITS_PER_LAYER * UNIT_STRIDE_BYTES;
    int g = 0;
    for (int u = 0; u < 6; ++u, ++g)
      pack_unit([&](int n, int k) { return win[(size_t)k * 384 + 64 * u + n]; }, 64, 128, dst + (size_t)g * UNIT_STRIDE_BYTES);
    for (int u = 0; u < 2; ++u, ++g)
      pack_unit([&](int n, int k) { return wo[(size_t)k * 128 + 64 * u + n]; }, 64, 128, dst + (size_t)g * UNIT_STRIDE_BYTES);
    auto pack_w1 = [&](int c) {
      pack_unit([&](int n, int k) { return w1[(size_t)k * 2048 + 64 * c + n]; }, 64, 128, dst + (size_t)g * UNIT_STRIDE_BYTES);
      ++g;
    };
    auto pack_w2 = [&](int c) {
      pack_unit([&](int n, int k) { return w2[(size_t)(64 * c + k) * 128 + n]; }, 128, 64, dst + (size_t)g * UNIT_STRIDE_BYTES);
      ++g;
    };
    // consumption order of the FFN pipeline: W1_0, then (W1_{c+1}, W2_c) ...
    pack_w1(0);
    for (int c = 0; c < NCHUNK; ++c) {
      if (c + 1 < NCHUNK) pack_w1(c + 1);
      pack_w2(c);
    }
  }
  void* d = nullptr;
  S3D_CUDA(cudaMalloc(&d, total));
  m->allocs.push_back(d);
  S3D_CUDA(cudaMemcpyAsync(d, img.data(), total, cudaMemcpyHostToDevice, st));
  S3D_CUDA(cudaStreamSynchronize(st));
  m->dectc.wimg = static_cast<__nv_bfloat16*>(d);
  m->dectc.wimg_elems = total / 2;
  return S3D_OK;
}

size_t decoder_tc_workspace_bytes(int64_t) { return 256; }

int decoder_tc(const s3d_model* m, const float* planes, int S, const QueryCtx& q, int64_t n, float out_scale, float* out,
               int precision, void*, size_t, cudaStream_t st) {
  if (m->K != 12 || m->dectc.wimg == nullptr) {
    set_error("decoder: tensor-core modes need n_slices == 12 (use S3D_PREC_FP32 otherwise)");
    return S3D_ERR_UNSUPPORTED;
  }
  if (n <= 0) return S3D_OK;
  TcParams p{};
  p.wimg = reinterpret_cast<const uint8_t*>(m->dectc.wimg);
  p.planes = planes;
  p.S = S;
  p.q = q;
  p.n = n;
  p.out_scale = out_scale;
  p.out = out;
  const DecF32& d = m->dec32;
  p.fcp_wt = d.fcp_wt; p.fcp_b = d.fcp_b; p.fcs_b = d.fcs_b; p.fco_w = d.fco_w; p.fco_b = d.fco_b;
  for (int l = 0; l < 3; ++l) {
    p.b_in[l] = d.L[l].in_proj.shift; p.b_o[l] = d.L[l].out_proj.shift;
    p.ln1w[l] = d.L[l].n1_w; p.ln1b[l] = d.L[l].n1_b;
    p.b1[l] = d.L[l].lin1.shift; p.b2[l] = d.L[l].lin2.shift;
    p.ln2w[l] = d.L[l].n2_w; p.ln2b[l] = d.L[l].n2_b;
  }
  p.num_tiles = (n + TILE_Q - 1) / TILE_Q;
  int dev = 0, sms = 148;
  S3D_CUDA(cudaGetDevice(&dev));
  S3D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)(p.num_tiles < sms ? p.num_tiles : sms);
  if (precision == S3D_PREC_BF16X3) {
    S3D_CUDA(cudaFuncSetAttribute(decoder_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    decoder_tc_kernel<3><<<grid, 192, SMEM_BYTES, st>>>(p);
  } else {
    S3D_CUDA(cudaFuncSetAttribute(decoder_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    decoder_tc_kernel<1><<<grid, 192, SMEM_BYTES, st>>>(p);
  }
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

// Self-test of the UMMA plumbing (descriptors, swizzle, bulk copy, TMEM load): see include/slice3d_b200.h.
int umma_selftest(int mode, int passes, const float* a_dev, const float* w_dev, float* d_dev, cudaStream_t st) {
  if ((mode != 0 && mode != 1) || (passes != 1 && passes != 3) || !a_dev || !w_dev || !d_dev) {
    set_error("selftest: bad argument");
    return S3D_ERR_BAD_ARG;
  }
  const int N = mode == 0 ? 64 : 128, K = mode == 0 ? 128 : 64;
  std::vector<float> w((size_t)N * K);
  S3D_CUDA(cudaMemcpyAsync(w.data(), w_dev, w.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
  S3D_CUDA(cudaStreamSynchronize(st));
  std::vector<uint8_t> img(UNIT_STRIDE_BYTES);
  pack_unit([&](int n, int k) { return w[(size_t)n * K + k]; }, N, K, img.data());
  void* d = nullptr;
  S3D_CUDA(cudaMalloc(&d, UNIT_STRIDE_BYTES));
  S3D_CUDA(cudaMemcpyAsync(d, img.data(), UNIT_STRIDE_BYTES, cudaMemcpyHostToDevice, st));
  if (passes == 3) {
    cudaFuncSetAttribute(umma_selftest_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    umma_selftest_kernel<3><<<1, 128, SMEM_BYTES, st>>>(a_dev, static_cast<const uint8_t*>(d), mode, d_dev);
  } else {
    cudaFuncSetAttribute(umma_selftest_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    umma_selftest_kernel<1><<<1, 128, SMEM_BYTES, st>>>(a_dev, static_cast<const uint8_t*>(d), mode, d_dev);
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
