// Micro-benchmark: does alternating tcgen05.mma kinds (kind::f16 / kind::f8f6f4) on one accumulator cost anything?
// Fully unrolled groups of 16 MMAs (cta_group::2, M = 256, N = 128), no per-MMA control flow in the issuing thread.
//   PATTERN 0: 16 x f16      1: 16 x f8      2: 8 x f16 then 8 x f8 (the fp16f8 FFN unit)      3: alternate every MMA
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I slice3d_b200/csrc tools/mma_mix.cu -o tools/_bin/mma_mix
#include <cstdio>
#include <vector>

#include "tc_ptx.cuh"
using namespace s3d::ptx;

template <int PATTERN, bool TS>
__global__ void __launch_bounds__(128, 1) mix_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const uint32_t a_s = sbase, b_s = sbase + 16384, bar = sbase + 16384 + 32768, tptr = bar + 16;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sgen)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if ((threadIdx.x >> 5) == 0) tmem_alloc_pair(tptr, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + (tptr - sbase));
  const bool leader = cluster_ctarank() == 0;
  if (threadIdx.x == 0 && leader) {
    constexpr uint32_t idesc = make_idesc_f16(128, 256);
    const uint32_t al = make_desc_lo(a_s), bl = make_desc_lo(b_s);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const bool f8 = PATTERN == 1 || (PATTERN == 2 && k >= 8) || (PATTERN == 3 && (k & 1));
        const uint32_t ko = 2 * (k & 3);
        if (TS) {
          if (f8) umma_f8_ts_pair_lo(tmem, tmem + 256 + 8 * (k & 3), bl + ko, idesc, 1u);
          else umma_ts_pair_lo(tmem, tmem + 256 + 8 * (k & 7), bl + ko, idesc, 1u);
        } else {
          if (f8) umma_f8_pair_lo(tmem, al + ko, bl + ko, idesc, 1u);
          else umma_pair_lo(tmem, al + ko, bl + ko, idesc, 1u);
        }
      }
    }
    umma_commit_pair(bar);
    mbar_wait(bar, 0);
    out[blockIdx.x] = clock64() - t0;
  } else if (threadIdx.x == 0) {
    mbar_wait(bar, 0);
  }
  tc_fence_before();
  cluster_sync_all();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc_pair(tmem, 512);
}

template <int PATTERN, bool TS>
void run(int grid, int iters) {
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  auto kern = mix_kernel<PATTERN, TS>;
  const int smem = 16384 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, iters, d);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("pattern %d: %s\n", PATTERN, cudaGetErrorString(e));
      return;
    }
  }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; i += 2) mx = h[i] > mx ? h[i] : mx;
  const char* names[4] = {"16 x f16", "16 x f8", "8 x f16 + 8 x f8", "f16 / f8 alternating"};
  printf("%s %-22s: %6.1f cycles per MMA (M=256, N=128; floor 64)\n", TS ? "TS" : "SS", names[PATTERN], (double)mx / (16.0 * iters));
  cudaFree(d);
}

int main() {
  run<0, false>(148, 3000); run<1, false>(148, 3000); run<2, false>(148, 3000); run<3, false>(148, 3000);
  run<0, true>(148, 3000);  run<1, true>(148, 3000);  run<2, true>(148, 3000);  run<3, true>(148, 3000);
  return 0;
}
