// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16) for the operand layouts the attention
// kernels use, as a function of N. The issuing warp runs convergently (umma_*_warp): a single thread inside divergent
// code pays ~100 cycles of issue overhead per MMA, which hides everything below that.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I weathermodel_b200/csrc -o /tmp/mmabench tools/mmabench.cu && /tmp/mmabench
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "wm_common.cuh"
using namespace wm;

__device__ long long g_t[4];

// MODE 0: A, B unswizzled K-major            1: A unswizzled K-major, B unswizzled MN-major
//      2: A, B 128B-swizzled K-major         3: A, B unswizzled MN-major
//      4: A from tensor memory, B MN-major   5: A from tensor memory, B K-major
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int n, int reps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  const int warp = warp_idx_uniform();
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 64 * 1024);
    uint64_t da = 0, db;
    uint32_t idesc, astep = 0, bstep;
    if (MODE == 0) { da = umma_smem_desc(a, 2048, 128, UMMA_SWZ_NONE); db = umma_smem_desc(b, 2048, 128, UMMA_SWZ_NONE); idesc = umma_idesc_bf16(128, n, 0, 0); astep = 4096 >> 4; bstep = 4096 >> 4; }
    else if (MODE == 1) { da = umma_smem_desc(a, 2048, 128, UMMA_SWZ_NONE); db = umma_smem_desc(b, 128, 2048, UMMA_SWZ_NONE); idesc = umma_idesc_bf16(128, n, 0, 1); astep = 4096 >> 4; bstep = 256 >> 4; }
    else if (MODE == 2) { da = umma_smem_desc(a, 16, 1024, UMMA_SWZ_128B); db = umma_smem_desc(b, 16, 1024, UMMA_SWZ_128B); idesc = umma_idesc_bf16(128, n, 0, 0); astep = 32 >> 4; bstep = 32 >> 4; }
    else if (MODE == 3) { da = umma_smem_desc(a, 128, 2048, UMMA_SWZ_NONE); db = umma_smem_desc(b, 128, 2048, UMMA_SWZ_NONE); idesc = umma_idesc_bf16(128, n, 1, 1); astep = 256 >> 4; bstep = 256 >> 4; }
    else if (MODE == 4) { db = umma_smem_desc(b, 128, 2048, UMMA_SWZ_NONE); idesc = umma_idesc_bf16(128, n, 0, 1); bstep = 256 >> 4; }
    else { db = umma_smem_desc(b, 2048, 128, UMMA_SWZ_NONE); idesc = umma_idesc_bf16(128, n, 0, 0); bstep = 4096 >> 4; }
    const long long t0 = clock64();
#pragma unroll 4
    for (int r = 0; r < reps; ++r) {
      const int s = r & 3;
      if (MODE >= 4) umma_ts_warp(tm + (r & 1) * 128, tm + 256 + s * 8, db + s * bstep, idesc, 1);
      else umma_ss_warp(tm + (r & 1) * 128, da + s * astep, db + s * bstep, idesc, 1);
    }
    const long long t1 = clock64();
    umma_commit_warp(&bar);
    mbar_wait(&bar, 0, 99);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { g_t[0] = t1 - t0; g_t[1] = t2 - t0; }
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int MODE>
static void run(const char* name) {
  const int smem = 161 * 1024 + 1024;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int n : {16, 48, 64, 96, 128}) {
    const int reps = 2048;
    k<MODE><<<1, 128, smem>>>(n, reps);
    cudaError_t e = cudaDeviceSynchronize();
    long long t[4];
    cudaMemcpyFromSymbol(t, g_t, sizeof(t));
    printf("%-36s N=%3d  issue %6.1f cyc/mma  complete %6.1f cyc/mma  (%s)\n", name, n, (double)t[0] / reps,
           (double)t[1] / reps, cudaGetErrorString(e));
  }
}

int main() {
  run<0>("A,B unswizzled K-major");
  run<1>("A unswz K-major, B unswz MN-major");
  run<2>("A,B SW128 K-major");
  run<3>("A,B unswizzled MN-major");
  run<4>("A TMEM, B unswz MN-major");
  run<5>("A TMEM, B unswz K-major");
  return 0;
}
