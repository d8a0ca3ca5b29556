// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (bf16 -> fp32) on sm_100a for the shapes the decoder
// uses.  One CTA (or CTA pair) per SM; one thread issues `iters` groups of `KS` MMAs (K = 16 each) into the
// same accumulator, commits, waits, and reports SM cycles per MMA.  Operand contents are zeros (timing only).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I slice3d_b200/csrc tools/mma_rate.cu -o /tmp/mma_rate
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_ptx.cuh"

using namespace s3d::ptx;

template <int CG, bool TS, int N, bool F8 = false, int MIX = 0>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int ks, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  // layout: A tile [128][64] bf16 SW128 (16 KB) | B tile [N/CG][64] (<= 32 KB) | barrier | tmem ptr
  const uint32_t a_s = sbase, b_s = sbase + 16384, bar = sbase + 16384 + 32768, tptr = bar + 16;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sgen)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (CG == 2) tmem_alloc_pair(tptr, 512);
    else tmem_alloc(tptr, 512);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + (tptr - sbase));
  const bool leader = (CG == 1) || cluster_ctarank() == 0;
  if (threadIdx.x == 0 && leader) {
    constexpr uint32_t idesc = make_idesc_bf16(N, 128 * CG);
    const uint64_t ad = make_desc_sw128(a_s), bd = make_desc_sw128(b_s);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < ks; ++k) {
        const uint32_t kin = (k & 3) * 32;
        if (F8 && (MIX == 0 || (MIX == -1 ? (it > 0 || k > 0) : MIX == -2 ? (it < iters / 2) : (k % (2 * MIX)) >= MIX))) {  // kind::f8f6f4, E4M3, K = 32 per instruction (CTA pairs only); MIX: alternate with kind::f16 every MIX MMAs
          if (TS) umma_f8_ts_pair_lo(tmem, tmem + 256 + 8 * (k & 3), make_desc_lo(b_s) + (kin >> 4), idesc, k ? 1u : 0u);
          else umma_f8_pair_lo(tmem, make_desc_lo(a_s) + (kin >> 4), make_desc_lo(b_s) + (kin >> 4), idesc, k ? 1u : 0u);
        } else if (TS) {
          if (CG == 2) umma_bf16_ts_pair(tmem, tmem + 256 + 8 * (k & 7), bd + (kin >> 4), idesc, k ? 1u : 0u);
          else umma_bf16_ts(tmem, tmem + 256 + 8 * (k & 7), bd + (kin >> 4), idesc, k ? 1u : 0u);
        } else {
          if (CG == 2) umma_bf16_pair(tmem, ad + (kin >> 4), bd + (kin >> 4), idesc, k ? 1u : 0u);
          else umma_bf16(tmem, ad + (kin >> 4), bd + (kin >> 4), idesc, k ? 1u : 0u);
        }
      }
    }
    if (CG == 2) umma_commit_pair(bar);
    else umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  } else if (threadIdx.x == 0) {
    mbar_wait(bar, 0);  // peer: the multicast commit arrives here too
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  if (warp == 0) {
    if (CG == 2) tmem_dealloc_pair(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

template <int CG, bool TS, int N, bool F8 = false, int MIX = 0>
void run(int grid, int iters, int ks) {
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  cudaMemset(d, 0, grid * sizeof(long long));
  auto kern = rate_kernel<CG, TS, N, F8, MIX>;
  const int smem = 16384 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, iters, ks, d);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("CG=%d %s N=%3d: %s\n", CG, TS ? "TS" : "SS", N, cudaGetErrorString(e));
      return;
    }
  }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; i += CG) mx = h[i] > mx ? h[i] : mx;
  const double per = (double)mx / ((double)iters * ks);
  printf("CG=%d %s %s N=%3d ks=%2d grid=%3d: %7.1f cycles/MMA  (floor %3d)  -> %5.1f %% of the %d-FLOP/clk/SM pipe\n", CG,
         TS ? "TS" : "SS", F8 ? (MIX ? (MIX == 8 ? "mix 8/8  " : MIX == 16 ? "mix 16/16" : MIX == 64 ? "mix 64/64" : MIX == 256 ? "mix 256  " : MIX == -1 ? "f16 once " : MIX == -2 ? "halves   " : "mix 1024 ") : "f8 K=32 ") : "f16 K=16", N, ks, grid, per, N / 2, 100.0 * (N / 2) / per, F8 ? 16384 : 8192);
  cudaFree(d);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  const int iters = 2000;
  for (int ks : {4, 8, 24}) {
    run<1, false, 64>(grid, iters, ks);
    run<1, true, 64>(grid, iters, ks);
    run<1, false, 128>(grid, iters, ks);
    run<1, true, 128>(grid, iters, ks);
    run<1, false, 256>(grid, iters, ks);
    run<1, true, 256>(grid, iters, ks);
    run<2, false, 64>(grid, iters, ks);
    run<2, true, 64>(grid, iters, ks);
    run<2, false, 128>(grid, iters, ks);
    run<2, true, 128>(grid, iters, ks);
    run<2, false, 256>(grid, iters, ks);
    run<2, true, 256>(grid, iters, ks);
    run<2, false, 128, true>(grid, iters, ks);
    run<2, true, 128, true>(grid, iters, ks);
    run<2, false, 256, true>(grid, iters, ks);
    run<2, true, 256, true>(grid, iters, ks);
  }
  // alternating kinds (the fp16f8 FFN unit: 8 kind::f16 MMAs, then 8 kind::f8f6f4 MMAs): cost of switching
  run<2, false, 128, true, 8>(grid, iters, 32);
  run<2, true, 128, true, 8>(grid, iters, 32);
  run<2, false, 128, true, 16>(grid, iters, 32);
  run<2, true, 128, true, 16>(grid, iters, 32);
  run<2, false, 128, true, 64>(grid, 500, 128);
  run<2, false, 128, true, 256>(grid, 200, 512);
  run<2, false, 128, true, 1024>(grid, 50, 2048);
  run<2, false, 128, true, -1>(grid, iters, 24);   // one kind::f16 MMA first, then only kind::f8f6f4
  run<2, false, 128, true, -2>(grid, iters, 24);   // first half of the iterations f8, second half f16 (one switch)
  return 0;
}
