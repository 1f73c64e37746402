// Micro-benchmark: how fast can one CTA stage 3 head slices [365 rows x 72 B, row pitch 3456 B] into shared memory?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ldbench tools/ldbench.cu && /tmp/ldbench
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

constexpr int S = 365, H = 16, DH = 36, D = H * DH, LD = 3 * D, NTH = 416;

__device__ long long g_t[8];

template <int MODE>
__global__ void __launch_bounds__(NTH, 1) k(const __nv_bfloat16* __restrict__ qkv, int* sink) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int bh = blockIdx.x % (9 * H), b = bh / H, h = bh % H;
  const uint8_t* base = reinterpret_cast<const uint8_t*>(qkv + (size_t)b * S * LD + h * DH);
  const int tid = threadIdx.x;
  long long t0 = clock64();
  if (MODE == 0) {  // 8-byte pieces, all loads first (27 per thread max), then stores
    uint2 v[24]; int off[24];
#pragma unroll
    for (int u = 0; u < 24; ++u) {
      int idx = tid + u * NTH; off[u] = -1;
      if (idx < 3 * S * 9) { int t = idx / (S * 9), i = idx % (S * 9), r = i / 9, p = i % 9;
        v[u] = __ldg(reinterpret_cast<const uint2*>(base + (size_t)r * LD * 2 + t * D * 2) + p); off[u] = t * 36864 + r * 96 + p * 8; }
    }
#pragma unroll
    for (int u = 0; u < 24; ++u) if (off[u] >= 0) *reinterpret_cast<uint2*>(sm + off[u]) = v[u];
  } else if (MODE == 1) {  // warp per row: lane < 18 loads 4 B -> one coalesced 72 B request per row
    const int warp = tid >> 5, lane = tid & 31, nw = NTH / 32;
    for (int rr = warp; rr < 3 * S; rr += nw * 8) {
      uint32_t v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { int row = rr + u * nw; if (row < 3 * S && lane < 18) { int t = row / S, r = row % S;
          v[u] = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)r * LD * 2 + t * D * 2) + lane); } }
#pragma unroll
      for (int u = 0; u < 8; ++u) { int row = rr + u * nw; if (row < 3 * S && lane < 18) { int t = row / S, r = row % S;
          *reinterpret_cast<uint32_t*>(sm + t * 36864 + r * 96 + lane * 4) = v[u]; } }
    }
  } else if (MODE == 2) {  // 16-byte aligned chunks covering the row segment (6 per row), all in flight
    uint4 v[16]; int off[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      int idx = tid + u * NTH; off[u] = -1;
      if (idx < 3 * S * 6) { int t = idx / (S * 6), i = idx % (S * 6), r = i / 6, c = i % 6;
        const uint8_t* rowp = base + (size_t)r * LD * 2 + t * D * 2; const uint8_t* al = (const uint8_t*)((uintptr_t)rowp & ~(uintptr_t)15);
        v[u] = __ldg(reinterpret_cast<const uint4*>(al) + c); off[u] = t * 36864 + r * 96 + c * 16; }
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) if (off[u] >= 0) *reinterpret_cast<uint4*>(sm + off[u]) = v[u];
  } else if (MODE == 3) {  // contiguous source: same bytes but packed [3][365][72 B] (what a head-major layout would give), 16 B loads
    const uint8_t* packed = reinterpret_cast<const uint8_t*>(qkv) + (size_t)bh * 3 * 36864;
    uint4 v[16]; int off[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) { int idx = tid + u * NTH; off[u] = -1; if (idx < 3 * 36864 / 16) { v[u] = __ldg(reinterpret_cast<const uint4*>(packed) + idx); off[u] = idx * 16; } }
#pragma unroll
    for (int u = 0; u < 16; ++u) if (off[u] >= 0) *reinterpret_cast<uint4*>(sm + off[u]) = v[u];
    for (int idx = tid + 16 * NTH; idx < 3 * 36864 / 16; idx += NTH) *reinterpret_cast<uint4*>(sm + idx * 16) = __ldg(reinterpret_cast<const uint4*>(packed) + idx);
  } else if (MODE == 4) {  // cp.async.bulk of the packed image: 3 instructions
    __shared__ uint64_t bar;
    const uint8_t* packed = reinterpret_cast<const uint8_t*>(qkv) + (size_t)bh * 3 * 36864;
    uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(3 * 36864));
      for (int t = 0; t < 3; ++t)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(sm + t * 36864)), "l"(packed + t * 36864), "r"(36864), "r"(bar_a) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar_a) : "memory");
  }
  __syncthreads();
  long long t1 = clock64();
  if (blockIdx.x == 0 && tid == 0) { g_t[0] = t1 - t0; }
  if (sm[tid] == 123 && sink) *sink = 1;
}

int main() {
  const int B = 9;  // 144 CTAs: one per SM
  size_t n = (size_t)B * S * LD + 3 * 36864 * B * H;  // also big enough for the packed view
  __nv_bfloat16* d; cudaMalloc(&d, n * 2); cudaMemset(d, 0, n * 2);
  int smem = 3 * 36864 + 1024;
  auto run = [&](auto kern, const char* name) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 3; ++rep) { kern<<<B * H, NTH, smem>>>(d, nullptr); cudaDeviceSynchronize(); }
    long long t; cudaMemcpyFromSymbol(&t, g_t, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); for (int rep = 0; rep < 20; ++rep) kern<<<B * H * 8, NTH, smem>>>(d, nullptr); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s CTA0 %7lld cycles; %d CTAs x20: %.3f ms -> %.1f us per wave-of-148\n", name, t, B * H * 8, ms / 20, ms / 20 / (B * H * 8 / 148.0) * 1e3);
  };
  run(k<0>, "8B pieces all-in-flight");
  run(k<1>, "warp-per-row 4B coalesced");
  run(k<2>, "16B aligned superset");
  run(k<3>, "packed contiguous 16B");
  run(k<4>, "packed cp.async.bulk");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
