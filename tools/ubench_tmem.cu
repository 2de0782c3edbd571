// Micro-benchmarks that size the field kernel's epilogue on B200 (sm_100a):
//   ld     : tcgen05.ld 32x32b.xN throughput vs epilogue warp count       (is the accumulator drain TMEM-read-bound?)
//   st     : tcgen05.st 32x32b.x8 throughput
//   epi_*  : one 64-column epilogue slice per iteration, as field_forward_kernel does it:
//            ld.x16 -> +bias -> cvt.relu.bf16x2 -> st.shared.v4 x2 -> fence.proxy.async -> arrive
//            variants: no fence / all four slices loaded up front / tcgen05.st instead of st.shared
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nsff_pl_b200/csrc -o tools/bin/ubench_tmem tools/ubench_tmem.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "nsff_sm100.cuh"

using namespace nsff::sm100;

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
        "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
        "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
        "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
        "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

enum Mode { LD16 = 0, LD32, LD64, ST8, EPI, EPI_NOFENCE, EPI_FRONT, EPI_TST, EPI_TST_FRONT, N_MODES };
static const char* kNames[] = {"ld.x16", "ld.x32", "ld.x64", "st.x8", "epi(smem+fence)", "epi(smem,no fence)",
                               "epi(4 slices loaded first)", "epi(tcgen05.st)", "epi(tcgen05.st, 4 loaded first)"};

// every "iteration" moves one 64-column x 128-row fp32 slice (32 KB) per CTA: warps = 4 quadrants x parts
template <int kMode>
__global__ void __launch_bounds__(512, 1) bench(int iters, unsigned long long* cycles, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sBar = smem + 65536, sPtr = sBar + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int q = warp & 3, part = warp >> 2, nparts = nw >> 2;      // columns of a 64-slice split over nparts warps
  const int cols = 64 / nparts;                                     // 64 (4 warps), 32 (8), 16 (16)
  if (threadIdx.x == 0) {
    mbar_init(sBar, (1u << 20) - 1u);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(sPtr, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tbase;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tbase) : "r"(sPtr));
  const uint32_t t_row = tbase + ((uint32_t)(q * 32) << 16);
  const int row = q * 32 + lane;
  const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
  const int rsw = row & 7;
  float acc = 0.f;
  __syncthreads();
  const unsigned long long t0 = clock64();
  if constexpr (kMode == LD16 || kMode == LD32 || kMode == LD64) {
    constexpr int kN = kMode == LD16 ? 16 : kMode == LD32 ? 32 : 64;
    // each warp reads `cols` columns per iteration in kN-column instructions
    for (int it = 0; it < iters; ++it) {
      const uint32_t c0 = (uint32_t)(((it & 3) * 64 + part * cols) & 255) + (uint32_t)((it >> 2) & 1) * 256u;
      for (int c = 0; c < cols; c += kN) {
        if constexpr (kN == 16) {
          uint32_t v[16];
          tmem_ld16(t_row + c0 + c, v);
          tmem_ld_wait();
          acc += __uint_as_float(v[0] ^ v[15]);
        } else if constexpr (kN == 32) {
          uint32_t v[32];
          tmem_ld32(t_row + c0 + c, v);
          tmem_ld_wait();
          acc += __uint_as_float(v[0] ^ v[31]);
        } else {
          uint32_t v[64];
          tmem_ld64(t_row + c0 + c, v);
          tmem_ld_wait();
          acc += __uint_as_float(v[0] ^ v[63]);
        }
      }
    }
  } else if constexpr (kMode == ST8) {
    // bf16 image of the slice: 32 columns per 64-column slice, cols/2 per warp
    for (int it = 0; it < iters; ++it) {
      const uint32_t c0 = (uint32_t)((it & 3) * 32 + part * (cols / 2));
      for (int c = 0; c < cols / 2; c += 8) {
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (uint32_t)(it + i);
        tmem_st8(t_row + c0 + c, v);
      }
      tmem_st_wait();
    }
  } else {
    // epilogue slice: this build assumes 16 warps (16 columns per warp per slice)
    constexpr bool kFront = kMode == EPI_FRONT || kMode == EPI_TST_FRONT;
    constexpr bool kTst = kMode == EPI_TST || kMode == EPI_TST_FRONT;
    constexpr bool kFence = kMode != EPI_NOFENCE;
    const float bias = (float)lane * 1e-3f;
    for (int it = 0; it < iters; it += 4) {            // one "layer" = 4 slices
      const uint32_t t_acc = t_row + (uint32_t)((it >> 2) & 1) * 256u;
      uint32_t v[4][16];
      if constexpr (kFront) {
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16(t_acc + (uint32_t)(c * 64 + part * 16), v[c]);
      } else {
        tmem_ld16(t_acc + (uint32_t)(part * 16), v[0]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        if constexpr (!kFront)
          if (c < 3) tmem_ld16(t_acc + (uint32_t)((c + 1) * 64 + part * 16), v[(c + 1) & 3]);
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          w[i] = pack_bf16x2_relu(__uint_as_float(v[c][2 * i]) + bias, __uint_as_float(v[c][2 * i + 1]) + bias);
        if constexpr (kTst) {
          // bf16 hidden image written back to TMEM (in place, into the already drained columns of this accumulator)
          tmem_st8(t_acc + (uint32_t)(c * 32 + part * 8), w);
          tmem_st_wait();
          tc_fence_before_sync();
        } else {
          const uint32_t row_base = smem + (uint32_t)c * 16384u + row_off;
          st_shared_v4(row_base + (uint32_t)((((part * 2) ^ rsw) & 7) << 4), w[0], w[1], w[2], w[3]);
          st_shared_v4(row_base + (uint32_t)((((part * 2 + 1) ^ rsw) & 7) << 4), w[4], w[5], w[6], w[7]);
          if constexpr (kFence) fence_proxy_async_smem();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sBar);
      }
    }
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 12345.678f) sink[0] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tbase, 512);
  }
}

template <int kMode>
static void run(int nw, int iters, unsigned long long* d_cyc, float* d_sink) {
  const int smem = 65536 + 1024 + 1024;
  cudaFuncSetAttribute(bench<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<kMode><<<148, nw * 32, smem>>>(iters, d_cyc, d_sink);   // warm-up
  bench<kMode><<<148, nw * 32, smem>>>(iters, d_cyc, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-34s nw=%2d  CUDA error %s\n", kNames[kMode], nw, cudaGetErrorString(e));
    exit(1);
  }
  unsigned long long h[148];
  cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < 148; ++i) mean += (double)h[i];
  mean /= 148.0;
  const double per_slice = mean / iters;
  const double bytes = kMode == ST8 ? 16384.0 : 32768.0;       // fp32 slice read / bf16 slice written
  printf("%-34s nw=%2d  %8.1f cycles per 64-col slice  (%6.1f B/clk/SM; a 128x256 layer = %6.0f cycles)\n", kNames[kMode],
         nw, per_slice, bytes / per_slice, per_slice * 4);
}

int main() {
  unsigned long long* d_cyc;
  float* d_sink;
  cudaMalloc(&d_cyc, 148 * sizeof(unsigned long long));
  cudaMalloc(&d_sink, 4);
  const int iters = 4096;
  for (int nw : {4, 8, 16}) {
    run<LD16>(nw, iters, d_cyc, d_sink);
    if (nw <= 8) run<LD32>(nw, iters, d_cyc, d_sink);
    if (nw <= 4) run<LD64>(nw, iters, d_cyc, d_sink);   // 16 warps own only 16 columns each
    run<ST8>(nw, iters, d_cyc, d_sink);
  }
  run<EPI>(16, iters, d_cyc, d_sink);
  run<EPI_NOFENCE>(16, iters, d_cyc, d_sink);
  run<EPI_FRONT>(16, iters, d_cyc, d_sink);
  run<EPI_TST>(16, iters, d_cyc, d_sink);
  run<EPI_TST_FRONT>(16, iters, d_cyc, d_sink);
  return 0;
}
