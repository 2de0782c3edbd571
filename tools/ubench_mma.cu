// tcgen05.mma issue/throughput micro-benchmark on B200: A operand from shared memory (.ss) vs from tensor memory (.ts),
// N = 256 / 128, cta_group 1 / 2, optionally with epilogue-like TMEM traffic (tcgen05.ld + tcgen05.st) from 16 other
// warps.  One thread per CTA (pair) issues `iters` groups of 4 MMAs (K = 16 each) + one commit, as the field kernel does.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nsff_pl_b200/csrc -o tools/bin/ubench_mma tools/ubench_mma.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "nsff_sm100.cuh"

using namespace nsff::sm100;

struct Args {
  int iters, n, ts, traffic;
  unsigned long long* cycles;
  unsigned int* epi_iters;
};

template <int kCtas>
__global__ void __launch_bounds__(576, 1) bench(Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem, sB = smem + 65536, sBar = smem + 65536 + 65536, sPtr = sBar + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kCtas == 2 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    mbar_init(sBar, 1);        // all MMAs done
    mbar_init(sBar + 8, 1);    // scratch commits
    fence_mbar_init();
  }
  if (warp == 17) {
    if constexpr (kCtas == 1) tmem_alloc(sPtr, 512);
    else tmem_alloc_pair(sPtr, 512);
  }
  tc_fence_before_sync();
  if constexpr (kCtas == 1) __syncthreads();
  else cluster_sync_all();
  tc_fence_after_sync();
  uint32_t tbase;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tbase) : "r"(sPtr));
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  unsigned long long t0 = 0, t1 = 0;
  if (warp == 17) {
    if (rank == 0 && elect_one()) {
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = ((sA & 0x3FFFFu) >> 4) | (1u << 16), b_lo0 = ((sB & 0x3FFFFu) >> 4) | (1u << 16);
      auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
      const uint32_t idesc = make_idesc_bf16(128 * kCtas, (uint32_t)a.n);
      t0 = clock64();
      for (int it = 0; it < a.iters; ++it) {
        const uint32_t d = tbase + (uint32_t)((it & 1) * 256);            // D region alternates, A in the other one
        const uint32_t at = tbase + (uint32_t)(((it & 1) ^ 1) * 256) + (uint32_t)((it >> 1) & 1) * 8u;
        const uint32_t a_lo = a_lo0 + (uint32_t)((it & 3) * 1024), b_lo = b_lo0 + (uint32_t)((it & 1) * 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (a.ts) {
            if constexpr (kCtas == 1) mma_bf16_ts(d, at + 16 * k, desc(b_lo + 2 * k), idesc, k > 0);
            else mma_bf16_ts_pair(d, at + 16 * k, desc(b_lo + 2 * k), idesc, k > 0);
          } else {
            if constexpr (kCtas == 1) mma_bf16_ss(d, desc(a_lo + 2 * k), desc(b_lo + 2 * k), idesc, k > 0);
            else mma_bf16_ss_pair(d, desc(a_lo + 2 * k), desc(b_lo + 2 * k), idesc, k > 0);
          }
        }
        if constexpr (kCtas == 1) mma_commit(sBar + 8);
        else mma_commit_pair_mc(sBar + 8, 1);
      }
      if constexpr (kCtas == 1) mma_commit(sBar);
      else mma_commit_pair_mc(sBar, 1);
      mbar_wait(sBar, 0);
      t1 = clock64();
      a.cycles[blockIdx.x] = t1 - t0;
      stop = 1;
    }
    __syncwarp();
  } else if (warp < 16 && a.traffic && rank == 0) {
    // epilogue-like TMEM traffic: ld.x16 + st.x8 per iteration per warp, until the issuer is done
    const int q = warp & 3, part = warp >> 2;
    const uint32_t t_row = tbase + ((uint32_t)(q * 32) << 16);
    int it = 0;
    while (!stop) {
      uint32_t v[16];
      tmem_ld16(t_row + (uint32_t)(((it & 3) * 64 + part * 16) + 256 * ((it >> 2) & 1)), v);
      tmem_ld_wait();
      tmem_st8(t_row + (uint32_t)(448 + part * 16), v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
      tmem_st_wait();
      ++it;
    }
    if (lane == 0 && warp == 0) a.epi_iters[blockIdx.x] = (unsigned)it;
  }
  tc_fence_before_sync();
  if constexpr (kCtas == 1) __syncthreads();
  else cluster_sync_all();
  if (warp == 17) {
    tc_fence_after_sync();
    if constexpr (kCtas == 1) tmem_dealloc(tbase, 512);
    else tmem_dealloc_pair(tbase, 512);
  }
}

template <int kCtas>
static void run(int n, int ts, int traffic, unsigned long long* d_cyc, unsigned int* d_it) {
  const int smem = 65536 + 65536 + 1024 + 1024, iters = 2048;
  cudaFuncSetAttribute(bench<kCtas>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  Args a{iters, n, ts, traffic, d_cyc, d_it};
  for (int rep = 0; rep < 2; ++rep) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(576);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCtas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, bench<kCtas>, a);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("cta_group %d n %d ts %d: CUDA error %s\n", kCtas, n, ts, cudaGetErrorString(e));
    exit(1);
  }
  unsigned long long h[148];
  cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  int cnt = 0;
  for (int i = 0; i < 148; i += kCtas) mean += (double)h[i], ++cnt;
  mean /= cnt;
  const double per_mma = mean / (iters * 4.0);
  unsigned int hi[148];
  cudaMemcpy(hi, d_it, sizeof(hi), cudaMemcpyDeviceToHost);
  double its = 0;
  for (int i = 0; i < 148; i += kCtas) its += hi[i];
  its /= cnt;
  const double floor_cyc = 128.0 * n / 256.0;      // per-SM floor for M = 128 rows per CTA
  printf("cta_group::%d  N=%3d  A from %s  %s : %7.1f cycles per MMA (floor %5.1f) -> %5.1f %% of the tensor peak\n", kCtas, n,
         ts ? "TMEM" : "smem", traffic ? "+ epilogue TMEM traffic" : "                       ", per_mma, floor_cyc,
         100.0 * floor_cyc / per_mma);
  if (traffic) printf("      epilogue warps: %.0f cycles per (ld.x16 + st.x8) iteration while the MMAs run\n", mean / its);
}

int main() {
  unsigned long long* d_cyc;
  cudaMalloc(&d_cyc, 148 * sizeof(unsigned long long));
  unsigned int* d_it;
  cudaMalloc(&d_it, 148 * sizeof(unsigned int));
  for (int traffic = 1; traffic < 2; ++traffic)
    for (int n : {256, 128, 64, 32, 16})
      for (int ts = 0; ts < 2; ++ts) {
        run<1>(n, ts, traffic, d_cyc, d_it);
        run<2>(n, ts, traffic, d_cyc, d_it);
      }
  return 0;
}
