// Micro-benchmark: cost of the signalling primitives the decoder's MMA issuer / compute warps hand work over with,
// on a CTA pair (cluster of 2) of sm_100a: tcgen05.commit issue cost and latency, mbarrier try_wait on a completed
// phase, elect + syncwarp, local and cluster-remote arrive -> waiter wake-up, and the issue cost of 8 MMAs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I slice3d_b200/csrc tools/sync_cost.cu -o tools/_bin/sync_cost
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_ptx.cuh"

using namespace s3d::ptx;

enum { T_CLOCK = 0, T_COMMIT_PAIR, T_COMMIT_ONE, T_COMMIT_LAT, T_TRYWAIT_DONE, T_ELECT_SYNC, T_FENCE_AFTER, T_MMA8_ISSUE,
       T_MMA8_COMMIT_WAIT, T_PING_LOCAL, T_PING_REMOTE, T_PING_COMMIT_REMOTE, T_COMMIT_LAT_ONE, T_MMA16_2COMMIT, T_PING_COMMIT_32, T_PING_COMMIT_16L, T_PING_COMMIT_NB, T_COUNT };
const char* NAMES[T_COUNT] = {"clock() pair", "commit.cta_group::2.multicast issue (back to back)", "commit.cta_group::1 issue (back to back)",
                              "commit(pair) -> own barrier flips (nothing outstanding)", "try_wait on a completed phase",
                              "elect_one + __syncwarp", "tcgen05.fence::after_thread_sync", "issue 8 MMAs (M256 N128 K16, SS)",
                              "8 MMAs + commit(pair) + wait (floor 512)", "ping-pong local arrive/wait (round trip)",
                              "ping-pong cluster-remote arrive/wait (round trip)", "commit(pair) -> peer waits -> remote arrive -> leader (round trip)",
                              "commit(cta_group::1) -> own barrier flips", "2 x (8 MMAs + commit(pair)) then wait both (floor 1024)",
                              "commit(pair) -> 16 + 16 warps wait -> 32 arrives (16 remote) -> leader (round trip)",
                              "commit(pair) -> 16 local warps wait -> 16 local arrives -> leader (round trip)",
                              "commit(pair) -> 16 + 16 warps wait -> bar.sync per CTA -> 2 arrives (1 remote) -> leader"};

__global__ void __launch_bounds__(544, 1) cost_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const uint32_t a_s = sbase, b_s = sbase + 32768, bar0 = sbase + 65536, tptr = bar0 + 128;
  auto bar = [&](int i) { return bar0 + 8u * i; };
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sgen)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(bar(i), 1);
    fence_barrier_init();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc_pair(tptr, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + (tptr - sbase));
  const bool leader = cluster_ctarank() == 0;
  long long* o = out + (blockIdx.x >> 1) * T_COUNT;
  constexpr uint32_t idesc = make_idesc_f16(128, 256);

  // ---- single-thread measurements on the leader
  if (leader && threadIdx.x == 0) {
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) acc += (uint32_t)clock();
    o[T_CLOCK] = clock64() - t0 + (acc == 12345u);
    // back-to-back commits (the barrier has count 1: every arrival completes a phase)
    t0 = clock64();
    for (int i = 0; i < iters; ++i) umma_commit_pair(bar(0));
    o[T_COMMIT_PAIR] = clock64() - t0;
  }
  cluster_sync_all();  // (the peer's copy of barrier 0 has flipped `iters` times too; parity tracked below)
  if (leader && threadIdx.x == 0) {
    // drain: wait until barrier 0 has seen all arrivals (iters even -> parity back to 0 ... unknown; just spin a while)
    for (int i = 0; i < 20000; ++i) asm volatile("nanosleep.u32 20;");
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) umma_commit(bar(1));
    o[T_COMMIT_ONE] = clock64() - t0;
    for (int i = 0; i < 20000; ++i) asm volatile("nanosleep.u32 20;");
    // commit -> own barrier flips (fresh barrier 2, parity sequence 0, 1, 0, ...)
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      umma_commit_pair(bar(2));
      mbar_wait(bar(2), i & 1);
    }
    o[T_COMMIT_LAT] = clock64() - t0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      umma_commit(bar(3));
      mbar_wait(bar(3), i & 1);
    }
    o[T_COMMIT_LAT_ONE] = clock64() - t0;
    // try_wait on a completed phase: barrier 3 has completed `iters` phases; parity of the LAST completed phase
    const uint32_t par = (iters - 1) & 1;
    t0 = clock64();
    uint32_t ok = 0;
    for (int i = 0; i < iters; ++i) ok += mbar_try_wait(bar(3), par);
    o[T_TRYWAIT_DONE] = clock64() - t0 + (ok == 7u);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) tc_fence_after();
    o[T_FENCE_AFTER] = clock64() - t0;
    // 8 MMAs issue cost (pipe kept short of saturation by waiting afterwards)
    const uint32_t al = make_desc_lo(a_s), bl = make_desc_lo(b_s);
    long long issue = 0, total = 0;
    for (int i = 0; i < iters; ++i) {
      const long long s0 = clock64();
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_pair_lo(tmem, al + 2 * (k & 3) + (k >> 2) * 1024, bl + 2 * (k & 3) + (k >> 2) * 512, idesc, k ? 1u : 0u);
      const long long s1 = clock64();
      umma_commit_pair(bar(4));
      mbar_wait(bar(4), i & 1);
      const long long s2 = clock64();
      issue += s1 - s0;
      total += s2 - s0;
    }
    o[T_MMA8_ISSUE] = issue;
    o[T_MMA8_COMMIT_WAIT] = total;
    total = 0;
    for (int i = 0; i < iters; ++i) {
      const long long s0 = clock64();
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_pair_lo(tmem, al + 2 * (k & 3) + (k >> 2) * 1024, bl + 2 * (k & 3) + (k >> 2) * 512, idesc, k ? 1u : 0u);
      umma_commit_pair(bar(5));
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_pair_lo(tmem + 128, al + 2 * (k & 3) + (k >> 2) * 1024, bl + 2 * (k & 3) + (k >> 2) * 512, idesc, k ? 1u : 0u);
      umma_commit_pair(bar(6));
      mbar_wait(bar(5), i & 1);
      mbar_wait(bar(6), i & 1);
      total += clock64() - s0;
    }
    o[T_MMA16_2COMMIT] = total;
  }
  // elect + syncwarp: whole warp 1 of the leader
  if (leader && warp == 1) {
    const long long t0 = clock64();
    uint32_t e = 0;
    for (int i = 0; i < iters; ++i) {
      e += elect_one();
      __syncwarp();
    }
    if (lane == 0) o[T_ELECT_SYNC] = clock64() - t0 + (e == 99999999u);
  }
  cluster_sync_all();
  // barriers 4.. were used; re-initialise a clean set for the ping-pongs
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(bar(i), 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  // ---- local ping-pong: leader warp 2 lane 0 <-> leader warp 3 lane 0
  if (leader && lane == 0 && warp == 2) {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      mbar_arrive(bar(0));
      mbar_wait(bar(1), i & 1);
    }
    o[T_PING_LOCAL] = clock64() - t0;
  }
  if (leader && lane == 0 && warp == 3) {
    for (int i = 0; i < iters; ++i) {
      mbar_wait(bar(0), i & 1);
      mbar_arrive(bar(1));
    }
  }
  cluster_sync_all();
  // ---- remote ping-pong: leader thread 0 <-> peer thread 0 (each waits on its own barrier 2, arrives remotely)
  if (threadIdx.x == 0) {
    const uint32_t remote = mapa_u32(bar(2), leader ? 1 : 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (leader) {
        mbar_arrive_cluster(remote);
        mbar_wait_cluster(bar(2), i & 1);
      } else {
        mbar_wait_cluster(bar(2), i & 1);
        mbar_arrive_cluster(remote);
      }
    }
    if (leader) o[T_PING_REMOTE] = clock64() - t0;
  }
  cluster_sync_all();
  // ---- commit(pair, multicast) -> peer's barrier 3 -> peer arrives remotely on the leader's barrier 4
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (leader) {
        umma_commit_pair(bar(3));
        mbar_wait_cluster(bar(4), i & 1);
      } else {
        mbar_wait(bar(3), i & 1);
        mbar_arrive_cluster(mapa_u32(bar(4), 0));
      }
    }
    if (leader) o[T_PING_COMMIT_REMOTE] = clock64() - t0;
  }
  cluster_sync_all();
  // ---- the decoder's H hand-off: commit(pair) wakes 16 warps in each CTA, every warp arrives on the leader's barrier
  // (bounded polling everywhere: a protocol error shows up as a failure count instead of a hang)
  auto bwait = [&](uint32_t b, uint32_t parity) -> bool {
    for (int k = 0; k < 2000000; ++k)
      if (mbar_try_wait(b, parity)) return true;
    return false;
  };
  for (int variant = 0; variant < 3; ++variant) {
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_init(bar(5), 1);
      mbar_init(bar(6), variant == 0 ? 32 : variant == 1 ? 16 : 2);
      fence_barrier_init();
    }
    cluster_sync_all();
    const int n_it = 200;
    if (warp == 0) {
      if (leader && lane == 0) {
        const long long t0 = clock64();
        int fails = 0;
        for (int i = 0; i < n_it; ++i) {
          umma_commit_pair(bar(5));
          if (!bwait(bar(6), i & 1)) { ++fails; break; }
        }
        o[T_PING_COMMIT_32 + variant] = fails ? -1 : (clock64() - t0) * (1000 / n_it);
      }
    } else if (leader || variant != 1) {
      const uint32_t dst = leader ? bar(6) : mapa_u32(bar(6), 0);
      for (int i = 0; i < n_it; ++i) {
        if (!bwait(bar(5), i & 1)) break;
        if (variant == 2) {
          asm volatile("bar.sync 1, 512;" ::: "memory");
          if (warp == 1 && lane == 0) {
            if (leader) mbar_arrive(dst);
            else mbar_arrive_cluster(dst);
          }
        } else {
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(dst);
            else mbar_arrive_cluster(dst);
          }
        }
      }
    }
    __syncwarp();
    cluster_sync_all();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tmem, 512);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  const int iters = 1000;
  long long* d;
  cudaMalloc(&d, (grid / 2) * T_COUNT * sizeof(long long));
  cudaMemset(d, 0, (grid / 2) * T_COUNT * sizeof(long long));
  const int smem = 65536 + 256 + 1024;
  cudaFuncSetAttribute(cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(544);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, cost_kernel, iters, d);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("error: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<long long> h((grid / 2) * T_COUNT);
  cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  for (int t = 0; t < T_COUNT; ++t) {
    long long mn = 1ll << 60, mx = 0;
    for (int c = 0; c < grid / 2; ++c) {
      const long long v = h[c * T_COUNT + t];
      mn = v < mn ? v : mn;
      mx = v > mx ? v : mx;
    }
    printf("%-75s %8.1f .. %8.1f cycles\n", NAMES[t], (double)mn / iters, (double)mx / iters);
  }
  return 0;
}
