// Micro-benchmark: MUFU.EX2 and FFMA+EX2 throughput per SM, and tcgen05.ld bandwidth, on one CTA of 384 threads
// (the softmax warp count of the attention forward kernel).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I weathermodel_b200/csrc -o /tmp/mufubench tools/mufubench.cu && /tmp/mufubench
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "wm_common.cuh"
using namespace wm;

__device__ long long g_t[8];
__device__ float g_sink[1024];

__global__ void __launch_bounds__(384, 1) k_ex2(int iters, int with_fma) {
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = 0.001f * (threadIdx.x + j);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = with_fma ? fmaf(x[j], 0.999f, -0.5f) : x[j];
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[j]) : "f"(a));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  g_sink[threadIdx.x] = s;
  if (threadIdx.x == 0) g_t[0] = t1 - t0;
}

// mode 0: cvt.rn.bf16x2.f32 alone; mode 1: two ex2 + one pack (the softmax inner loop's ratio); mode 2: two ex2 + pack +
// fma + add + min (all of the inner loop's arithmetic)
__global__ void __launch_bounds__(384, 1) k_mix(int iters, int mode) {
  float x[8];
  uint32_t acc = 0;
  float sum = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = 0.001f * (threadIdx.x + j);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float a = x[j], b = x[j + 1];
      if (mode == 2) {
        a = fminf(fmaf(a, 0.999f, -0.5f), 64.0f);
        b = fminf(fmaf(b, 0.999f, -0.5f), 64.0f);
      }
      if (mode >= 1) {
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(a));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(b));
      }
      if (mode == 2) sum += a + b;
      uint32_t pk;
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(b), "f"(a));
      acc ^= pk;
      x[j] = a * 0.5f;
      x[j + 1] = b * 0.5f;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  g_sink[threadIdx.x] = __uint_as_float(acc) + sum;
  if (threadIdx.x == 0) g_t[2] = t1 - t0;
}

// 12 warps read 128 lanes x 384 columns of TMEM `reps` times with 32-column loads
__global__ void __launch_bounds__(416, 1) k_ldtm(int reps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 12) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < 12) {
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int sl = warp >> 2;
    asm volatile("bar.sync 1, 384;" ::: "memory");
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tm + lane_sel + c * 96 + sl * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j];
      }
    }
    asm volatile("bar.sync 1, 384;" ::: "memory");
    t1 = clock64();
    g_sink[threadIdx.x] = __uint_as_float(acc);
    if (threadIdx.x == 0) g_t[1] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
  long long t[8];
  for (int f = 0; f < 2; ++f) {
    const int iters = 4096;
    k_ex2<<<1, 384>>>(iters, f);
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(t, g_t, sizeof(t));
    printf("ex2%s: %.2f per clock per SM (384 threads x %d)\n", f ? " + ffma" : "", 384.0 * iters * 8 / t[0], iters * 8);
  }
  for (int mode = 0; mode < 3; ++mode) {
    const int iters = 4096;
    k_mix<<<1, 384>>>(iters, mode);
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(t, g_t, sizeof(t));
    const char* nm[3] = {"bf16x2 pack alone", "2 ex2 + 1 pack", "2 (fma, min, ex2, add) + 1 pack"};
    printf("%s: %.2f ELEMENTS per clock per SM\n", nm[mode], 384.0 * iters * 8 / t[2]);
  }
  const int reps = 256;
  k_ldtm<<<1, 416>>>(reps);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(t, g_t, sizeof(t));
  printf("tcgen05.ld 32x32b.x32 by 12 warps: %.1f B/clk per SM, %.0f cycles per 128x384 fp32 tile (%s)\n",
         128.0 * 384 * 4 * reps / t[1], (double)t[1] / reps, cudaGetErrorString(e));
  return 0;
}
