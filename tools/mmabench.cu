// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16) for the operand layouts the attention
// kernels use (unswizzled core-matrix columns) against 128B-swizzled K-major tiles, as a function of N.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I weathermodel_b200/csrc -o /tmp/mmabench tools/mmabench.cu && /tmp/mmabench
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "wm_common.cuh"
using namespace wm;

__device__ long long g_t[4];

// mode 0: A, B unswizzled K-major (LBO = rows*16 between k chunks, SBO = 128)
// mode 1: A unswizzled K-major, B unswizzled MN-major (LBO = 128 along k, SBO = rows*16 along n)
// mode 2: A, B 128B-swizzled K-major (row pitch 128 B, SBO = 1024)
// mode 3: A, B unswizzled MN-major
__global__ void __launch_bounds__(128, 1) k(int mode, int n, int reps, int distinct) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 64 * 1024);
    uint64_t da, db;
    uint32_t idesc, astep, bstep;
    if (mode == 0) { da = umma_smem_desc(a, 2048, 128, UMMA_SWZ_NONE); db = umma_smem_desc(b, 6144, 128, UMMA_SWZ_NONE); idesc = umma_idesc_bf16(128, n, 0, 0); astep = 4096; bstep = 12288; }
    else if (mode == 1) { da = umma_smem_desc(a, 2048, 128, UMMA_SWZ_NONE); db = umma_smem_desc(b, 128, 6144, UMMA_SWZ_NONE); idesc = umma_idesc_bf16(128, n, 0, 1); astep = 4096; bstep = 256; }
    else if (mode == 2) { da = umma_smem_desc(a, 16, 1024, UMMA_SWZ_128B); db = umma_smem_desc(b, 16, 1024, UMMA_SWZ_128B); idesc = umma_idesc_bf16(128, n, 0, 0); astep = 32; bstep = 32; }
    else { da = umma_smem_desc(a, 128, 2048, UMMA_SWZ_NONE); db = umma_smem_desc(b, 128, 6144, UMMA_SWZ_NONE); idesc = umma_idesc_bf16(128, n, 1, 1); astep = 256; bstep = 256; }
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int s = distinct ? (r & 3) : 0;
      umma_ss(tm + (r & 1) * 256, umma_desc_advance(da, s * astep), umma_desc_advance(db, s * bstep), idesc, 1);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0, 99);
    long long t2 = clock64();
    if (blockIdx.x == 0) { g_t[0] = t1 - t0; g_t[1] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
  const int smem = 161 * 1024 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[4] = {"A,B unswizzled K-major", "A unswz K-major, B unswz MN-major", "A,B SW128 K-major", "A,B unswizzled MN-major"};
  for (int grid : {1, 148})
    for (int mode = 0; mode < 4; ++mode)
      for (int n : {16, 48, 96, 128, 256}) {
        const int reps = 2048;
        k<<<grid, 128, smem>>>(mode, n, reps, 1);
        cudaError_t e = cudaDeviceSynchronize();
        long long t[4];
        cudaMemcpyFromSymbol(t, g_t, sizeof(t));
        printf("grid %3d  %-36s N=%3d  issue %6.1f cyc/mma  complete %6.1f cyc/mma  (%s)\n", grid, names[mode], n,
               (double)t[0] / reps, (double)t[1] / reps, cudaGetErrorString(e));
      }
  return 0;
}
