// wm_elementwise.cu -- the HBM-bound kernels of the WeatherModel training step (sm_100a).
//
//   mask_bert / mask_former : bit-exact replay of torch.rand on the CUDA generator
//                             (reference: src/pretraining/dataloader/pretraining_dataloader.py:56-84;
//                              torch/include/ATen/native/cuda/DistributionTemplates.h:50-89)
//   embed_fwd               : normalise + mask + Linear(34->D) + sinusoidal PE, one pass
//                             (reference: src/pretraining/models/weatherbert.py:101-115,
//                              src/utils/utils.py:63-74, src/base_models/vanilla_pos_encoding.py:57)
//   layernorm_fwd / bwd     : post-LN of nn.TransformerEncoderLayer (torch/nn/modules/transformer.py:951-957)
//   loss_bert / loss_former : fused loss + dLoss/dY (reference: src/pretraining/trainers/
//                             weatherbert_trainer.py:55-60, weatherformer_trainer.py:68-111,
//                             src/utils/losses.py:10-47, src/pretraining/models/weatherformer.py:87-92)
//   adam                    : torch.optim.Adam single step over a flat parameter bucket
//                             (reference: src/base_trainer/base_trainer.py:337)
#include "wm_kernels.h"

namespace wm {

// ------------------------------------------------------------------------------------------------
// torch.rand replay. ATen launches grid_x blocks of 256 threads; thread idx draws Philox blocks
// (seed, subsequence = idx, offset/4 + iter) and writes component ii to element
// idx + iter*4*T + ii*T with T = 256*grid_x.
// ------------------------------------------------------------------------------------------------
WM_DEVICE float torch_rand_element(uint64_t seed, uint64_t offset_blocks, int64_t T, int64_t li) {
  const int64_t iter = li / (4 * T);
  const int64_t rem = li - iter * 4 * T;
  const int ii = static_cast<int>(rem / T);
  const int64_t idx = rem - static_cast<int64_t>(ii) * T;
  const Philox4 r = philox4x32_10(seed, static_cast<uint64_t>(idx), offset_blocks + static_cast<uint64_t>(iter));
  const uint32_t u = ii == 0 ? r.x : ii == 1 ? r.y : ii == 2 ? r.z : r.w;
  const float v = curand_uniform_from_u32(u);
  return v == 1.0f ? 0.0f : v;  // ATen reverses (0,1] -> [0,1)
}

__global__ void __launch_bounds__(256)
mask_bert_kernel(uint64_t seed, uint64_t offset_blocks, float p, int64_t numel, uint8_t* __restrict__ mask,
                 float* __restrict__ rand_out) {
  const int64_t T = static_cast<int64_t>(gridDim.x) * 256;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  uint64_t iter = 0;
  for (int64_t base = idx; base < numel; base += 4 * T, ++iter) {
    const Philox4 r = philox4x32_10(seed, static_cast<uint64_t>(idx), offset_blocks + iter);
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int64_t li = base + T * ii;
      if (li < numel) {
        float v = curand_uniform_from_u32(u[ii]);
        v = v == 1.0f ? 0.0f : v;
        mask[li] = v < p ? 1 : 0;
        if (rand_out) rand_out[li] = v;
      }
    }
  }
}

int launch_mask_bert(uint64_t seed, uint64_t philox_offset, int grid_x, float p, int64_t numel,
                     uint8_t* mask, float* rand_out, cudaStream_t stream) {
  if (numel <= 0) return WM_OK;
  if (grid_x <= 0 || (philox_offset & 3)) return WM_ERR_ARG;
  mask_bert_kernel<<<grid_x, 256, 0, stream>>>(seed, philox_offset / 4, p, numel, mask, rand_out);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// one warp per sample: lane i holds v_i = rand[b, i]; mask[b, rank(v_i)] = (i < n_masked)
__global__ void __launch_bounds__(256)
mask_former_kernel(uint64_t seed, uint64_t offset_blocks, int64_t T, int n_masked, int64_t n_samples,
                   int F, uint8_t* __restrict__ mask) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= n_samples) return;
  float v = 2.0f;
  if (lane < F) v = torch_rand_element(seed, offset_blocks, T, b * F + lane);
  int rank = 0;
  for (int k = 0; k < F; ++k) {
    const float vk = __shfl_sync(0xffffffffu, v, k);
    rank += (vk < v) || (vk == v && k < lane);  // stable sort tie order
  }
  // position `rank` holds original index `lane`; argsort(...) < n  <=>  lane < n
  if (lane < F) mask[b * F + rank] = lane < n_masked ? 1 : 0;
}

int launch_mask_former(uint64_t seed, uint64_t philox_offset, int grid_x, int n_masked, int64_t n_samples,
                       int n_features, uint8_t* mask, cudaStream_t stream) {
  if (n_samples <= 0) return WM_OK;
  if (grid_x <= 0 || (philox_offset & 3) || n_features > 32 || n_features <= 0) return WM_ERR_ARG;
  const int blocks = static_cast<int>((n_samples + 7) / 8);
  mask_former_kernel<<<blocks, 256, 0, stream>>>(seed, philox_offset / 4, static_cast<int64_t>(grid_x) * 256,
                                                  n_masked, n_samples, n_features, mask);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// embed_fwd: out[t, :] = bf16( ([w * ~m, year', lat', lon'] . W_in^T + b_in) + PE[t % S] ), fp32 FMA throughout.
// Register-tiled like an SGEMM with K = 34: a CTA owns a strip of CT <= 320 output columns, keeps that strip of
// W_in^T in shared memory for its whole life and walks over 64-token tiles (persistent: one launch wave, two CTAs
// per SM so that one CTA's tile load overlaps the other's arithmetic). A thread computes 8 tokens x 8 columns with
// packed fp32 FMAs (FFMA2: one issue slot per two multiply-adds): per input channel 4 + 2 shared-memory loads feed
// 32 FFMA2 -- ~22 issued instructions per output element-pair... the first version (one column per thread, scalar
// FMAs, 2-byte stores) needed 55 per element and ran at 0.64-0.76 ms for the large shape (5 % of the HBM roofline).
// The input tile is stored DUPLICATED ((x, x) pairs) so that the broadcast operand of an FFMA2 needs no MOV.
// ------------------------------------------------------------------------------------------------
constexpr int kEmbTok = 64;   // tokens per tile
constexpr int kXinLd = 64;    // padded bf16 copy of the 34-channel input row (wgrad operand)
constexpr int kEmbMaxCT = 320;

__global__ void __launch_bounds__(kEmbMaxCT, 2)
embed_fwd_kernel(const float* __restrict__ weather, const uint8_t* __restrict__ mask, int64_t msb, int64_t mss,
                 const float* __restrict__ year, const float* __restrict__ coords,
                 const float* __restrict__ w_in, const float* __restrict__ b_in, const float* __restrict__ pe,
                 __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ xin, int B, int S, int F, int D, int CT) {
  extern __shared__ __align__(16) float smf[];
  const int Fin = F + 3;
  float* sW = smf;                       // [Fin][CT]      W_in^T strip
  float* sX = smf + Fin * CT;            // [Fin][kEmbTok] input tile, channel-major
  float* sRaw = sX + Fin * kEmbTok;      // [kEmbTok][F]   the tile's weather rows as they lie in memory
  const int M = B * S;                   // (host-checked to fit 31 bits)
  const int c_base = blockIdx.y * CT;    // first column of this CTA's strip
  const int ncols = min(CT, D - c_base); // a multiple of 8
  const int ntx = CT >> 3, half = CT >> 1;
  const int tx = threadIdx.x % ntx, ty = threadIdx.x / ntx;  // column group / token group (8 x 8 outputs per thread)
  // A thread's 8 columns are two runs of 4, half a strip apart: consecutive lanes then read consecutive 16-byte pieces
  // of a W row (conflict-free LDS.128; 8 consecutive columns per lane made every such load a 2-way bank conflict)
  const int colA = tx * 4, colB = half + tx * 4;
  const bool active = ty < (kEmbTok >> 3);
  const bool okA = active && colA < ncols, okB = active && colB < ncols;
  for (int i = threadIdx.x; i < Fin * CT; i += blockDim.x) {  // coalesced over d within a channel row
    const int c = i / CT, d = i - c * CT;
    sW[i] = d < ncols ? w_in[static_cast<size_t>(c_base + d) * Fin + c] : 0.0f;
  }
  const int ntiles = (M + kEmbTok - 1) / kEmbTok;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int t0 = tile * kEmbTok;
    const int nt = min(kEmbTok, M - t0);
    __syncthreads();  // previous tile's readers are done with sX / sRaw (and, first time round, sW is complete)
    // Both staging loops keep EIGHT independent loads in flight per thread: one load per iteration exposed the full
    // global-memory latency ~15 times per tile (0.40 ms for the large shape, 4x the arithmetic).
    {  // the tile's weather block is contiguous in memory: straight coalesced copy
      const float* src = weather + static_cast<size_t>(t0) * F;
      const int n = nt * F;
      for (int i0 = threadIdx.x; i0 < n; i0 += 8 * blockDim.x) {
        float r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * blockDim.x;
          r[u] = i < n ? __ldg(src + i) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * blockDim.x;
          if (i < n) sRaw[i] = r[u];
        }
      }
    }
    __syncthreads();
    for (int i0 = threadIdx.x; i0 < kEmbTok * Fin; i0 += 8 * blockDim.x) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        const int c = i / kEmbTok, tt = i - c * kEmbTok;  // token fastest: conflict-free smem on both sides (F is odd)
        const int t = t0 + tt;
        v[u] = 0.0f;
        if (i < kEmbTok * Fin && t < M) {
          const int b = t / S, sq = t - b * S;
          if (c < F) {
            const float w = sRaw[tt * F + c];
            const uint8_t m = mask[b * msb + sq * mss + c];
            v[u] = m ? w * 0.0f : w;  // weather * (~mask): keeps NaN/Inf semantics of the multiply
          } else if (c == F) {
            v[u] = (__ldg(year + t) - 1970.0f) / 100.0f;
          } else if (c == F + 1) {
            v[u] = __ldg(coords + b * 2) / 360.0f;
          } else {
            v[u] = __ldg(coords + b * 2 + 1) / 180.0f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < kEmbTok * Fin) sX[i] = v[u];
      }
    }
    __syncthreads();
    if (xin && blockIdx.y == 0) {  // bf16 copy of the input rows for the in_proj weight gradient
      // lanes run over tokens (conflict-free reads of the channel-major tile); a lane writes 16 bytes = 8 channels
      for (int i = threadIdx.x; i < kEmbTok * (kXinLd / 8); i += blockDim.x) {
        const int c8 = (i / kEmbTok) * 8, tt = i - (i / kEmbTok) * kEmbTok;
        const int t = t0 + tt;
        if (t < M) {
          float a[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = c8 + j < Fin ? sX[(c8 + j) * kEmbTok + tt] : 0.0f;
          uint4 o;
          o.x = pack_bf16x2(a[0], a[1]);
          o.y = pack_bf16x2(a[2], a[3]);
          o.z = pack_bf16x2(a[4], a[5]);
          o.w = pack_bf16x2(a[6], a[7]);
          *reinterpret_cast<uint4*>(xin + static_cast<size_t>(t) * kXinLd + c8) = o;
        }
      }
    }
    if (!okA) continue;
    uint64_t acc[8][4];  // [token][column pair]: pairs 0, 1 = run A, pairs 2, 3 = run B
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][j] = 0ull;
    const float* xrow = sX + ty * 8;
#pragma unroll 1
    for (int c = 0; c < Fin; ++c) {
      const ulonglong2 wa = *reinterpret_cast<const ulonglong2*>(sW + c * CT + colA);
      const ulonglong2 wb = *reinterpret_cast<const ulonglong2*>(sW + c * CT + (okB ? colB : colA));
      const float4 x0 = *reinterpret_cast<const float4*>(xrow + c * kEmbTok);      // tokens 0-3 (warp broadcast)
      const float4 x1 = *reinterpret_cast<const float4*>(xrow + c * kEmbTok + 4);  // tokens 4-7
      const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const uint64_t xx = f2_pack(xs[t], xs[t]);
        acc[t][0] = f2_fma(xx, wa.x, acc[t][0]);
        acc[t][1] = f2_fma(xx, wa.y, acc[t][1]);
        acc[t][2] = f2_fma(xx, wb.x, acc[t][2]);
        acc[t][3] = f2_fma(xx, wb.y, acc[t][3]);
      }
    }
    // epilogue: (acc + bias) + PE[position]; 4 bf16 = one 8-byte store per run and token; position kept incrementally
    const int tfirst = t0 + ty * 8;
    int sidx = tfirst % S;
    const float4 bA = __ldg(reinterpret_cast<const float4*>(b_in + c_base + colA));
    const float4 bB = okB ? __ldg(reinterpret_cast<const float4*>(b_in + c_base + colB)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int tok = tfirst + t;
      if (tok < M) {
        const float* prow = pe + static_cast<size_t>(sidx) * D + c_base;
        __nv_bfloat16* orow = out + static_cast<size_t>(tok) * D + c_base;
        float a0, a1, a2, a3;
        f2_unpack(acc[t][0], a0, a1);
        f2_unpack(acc[t][1], a2, a3);
        const float4 pA = __ldg(reinterpret_cast<const float4*>(prow + colA));
        uint2 o;
        o.x = pack_bf16x2((a0 + bA.x) + pA.x, (a1 + bA.y) + pA.y);
        o.y = pack_bf16x2((a2 + bA.z) + pA.z, (a3 + bA.w) + pA.w);
        *reinterpret_cast<uint2*>(orow + colA) = o;
        if (okB) {
          f2_unpack(acc[t][2], a0, a1);
          f2_unpack(acc[t][3], a2, a3);
          const float4 pB = __ldg(reinterpret_cast<const float4*>(prow + colB));
          o.x = pack_bf16x2((a0 + bB.x) + pB.x, (a1 + bB.y) + pB.y);
          o.y = pack_bf16x2((a2 + bB.z) + pB.z, (a3 + bB.w) + pB.w);
          *reinterpret_cast<uint2*>(orow + colB) = o;
        }
      }
      if (++sidx == S) sidx = 0;
    }
  }
}

int launch_embed_fwd(const float* weather, const uint8_t* mask, int64_t msb, int64_t mss, const float* year,
                     const float* coords, const float* w_in, const float* b_in, const float* pe,
                     __nv_bfloat16* out, __nv_bfloat16* xin, int B, int S, int F, int D, cudaStream_t stream) {
  if (B <= 0 || S <= 0 || F <= 0 || F + 3 > 40 || (D & 7)) return WM_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(pe) | reinterpret_cast<uintptr_t>(b_in)) & 15u) return WM_ERR_ALIGN;
  const int64_t M = static_cast<int64_t>(B) * S;
  if (M >= (1ll << 31) - kEmbTok) return WM_ERR_SHAPE;  // 32-bit token indices inside the kernel
  const int Fin = F + 3;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int strips = (D + kEmbMaxCT - 1) / kEmbMaxCT;
  const int CT = ((D / 8 + strips - 1) / strips) * 8;   // columns per strip: a multiple of 8, <= 320
  const int threads = ((CT + 31) / 32) * 32;            // CT/8 column groups x 8 token groups (+ idle lanes of the last warp)
  const int smem = (Fin * CT + Fin * kEmbTok + kEmbTok * F) * 4;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(embed_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return WM_ERR_CUDA;
  const int64_t ntiles = (M + kEmbTok - 1) / kEmbTok;
  // persistent: as many CTAs per strip as stay resident (registers: two 320-thread CTAs per SM; narrow strips: more)
  const int per_sm = CT >= 160 ? 2 : (CT >= 64 ? 4 : 8);
  const int gx = static_cast<int>(ntiles < static_cast<int64_t>(sms) * per_sm / strips ? ntiles : static_cast<int64_t>(sms) * per_sm / strips);
  dim3 grid(gx > 0 ? gx : 1, strips);
  embed_fwd_kernel<<<grid, threads, smem, stream>>>(weather, mask, msb, mss, year, coords, w_in, b_in, pe, out, xin, B,
                                                    S, F, D, CT);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row; lane owns 8-element chunks c = lane + 32*i (D % 8 == 0, D <= 768).
// ------------------------------------------------------------------------------------------------
constexpr int kLnMaxChunks = 3;

WM_DEVICE void ln_load_row(const __nv_bfloat16* row, int nchunks, int lane, float (&v)[kLnMaxChunks][8]) {
#pragma unroll
  for (int i = 0; i < kLnMaxChunks; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(row) + c);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[i][2 * j] = bf16_lo(aw[j]);
        v[i][2 * j + 1] = bf16_hi(aw[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.0f;
    }
  }
}

WM_DEVICE void ln_load_raw(const __nv_bfloat16* row, int nchunks, int lane, uint4 (&v)[kLnMaxChunks]) {
#pragma unroll
  for (int i = 0; i < kLnMaxChunks; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < nchunks ? __ldg(reinterpret_cast<const uint4*>(row) + c) : make_uint4(0u, 0u, 0u, 0u);
  }
}
WM_DEVICE void ln_unpack(const uint4 (&r)[kLnMaxChunks], float (&v)[kLnMaxChunks][8]) {
#pragma unroll
  for (int i = 0; i < kLnMaxChunks; ++i) {
    const uint32_t aw[4] = {r[i].x, r[i].y, r[i].z, r[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[i][2 * j] = bf16_lo(aw[j]);
      v[i][2 * j + 1] = bf16_hi(aw[j]);
    }
  }
}

// Chunk width: a lane owns kW consecutive elements per pass. At D = 576 the 72 eight-element chunks need three passes of
// which the last keeps 8 of 32 lanes busy (24 element slots per lane for 18 elements); 144 four-element chunks need
// five passes = 20 slots: 17 % fewer issued instructions in kernels that are issue-bound (profiles/r01_ln_bwd_full.txt).
template <int kW> struct LnVec;
template <> struct LnVec<8> { using T = uint4; };
template <> struct LnVec<4> { using T = uint2; };
template <int kW>
WM_DEVICE void ln_unpack_w(const typename LnVec<kW>::T& r, float (&v)[kW]) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
  for (int j = 0; j < kW / 2; ++j) {
    v[2 * j] = bf16_lo(w[j]);
    v[2 * j + 1] = bf16_hi(w[j]);
  }
}
template <int kW>
WM_DEVICE typename LnVec<kW>::T ln_pack_w(const float (&o)[kW]) {
  typename LnVec<kW>::T pk;
  uint32_t* w = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
  for (int j = 0; j < kW / 2; ++j) w[j] = pack_bf16x2(o[2 * j], o[2 * j + 1]);
  return pk;
}
// passes that hold an element of a row of D elements, and the chunk width that needs fewer element slots per lane
static inline int ln_slots(int D, int w) { return ((D / w + 31) / 32) * w; }
int g_ln_bwd_width = 0;  // 0 = default (8), 4 or 8 forced -- wm_set_option("ln_bwd_width", v)
int g_ln_bwd_rows = 0;   // 0 = default, 14 or 15 row warps per CTA in the encoder form -- wm_set_option("ln_bwd_rows", v)
int g_ln_fwd_width = 0;  // 0 = auto (ln_pick_width), 4 or 8 forced -- wm_set_option("ln_fwd_width", v)
static inline int ln_pick_width(int D) { return ln_slots(D, 4) < ln_slots(D, 8) ? 4 : 8; }

// kRows rows per warp: the loads of ALL its rows are issued before the first reduction, so a warp keeps kRows x 1152 B
// (D = 576) in flight instead of one row's -- at one row per warp and 32 resident warps an SM had 36 KB of reads in
// flight, short of what 6.5 TB/s at ~1 us latency asks for (the kernel ran at 4.7 TB/s).
#ifndef WM_LN_FWD_ROWS
#define WM_LN_FWD_ROWS 2
#endif
constexpr int kLnFwdRows = WM_LN_FWD_ROWS;
template <int kW>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int M, int D, float eps) {
  using Vec = typename LnVec<kW>::T;
  constexpr int kIt = (kLnMaxChunks * 8) / kW;  // passes for the widest supported row
  const int row0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * kLnFwdRows;
  const int lane = threadIdx.x & 31;
  if (row0 >= M) return;
  const int nchunks = D / kW;
  Vec raw[kLnFwdRows][kIt];
#pragma unroll
  for (int r = 0; r < kLnFwdRows; ++r) {
    const Vec* xr = reinterpret_cast<const Vec*>(x + static_cast<size_t>(row0 + r) * D);
#pragma unroll
    for (int i = 0; i < kIt; ++i) {
      const int c = lane + 32 * i;
      raw[r][i] = (c < nchunks && row0 + r < M) ? __ldg(xr + c) : Vec{};
    }
  }
#pragma unroll
  for (int r = 0; r < kLnFwdRows; ++r) {
    const int row = row0 + r;
    if (row >= M) break;  // warp-uniform
    float v[kIt][kW];
#pragma unroll
    for (int i = 0; i < kIt; ++i) ln_unpack_w<kW>(raw[r][i], v[i]);  // (chunks past the row are zero)
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < kIt; ++i)
#pragma unroll
      for (int j = 0; j < kW; ++j) s += v[i][j];
    const float mean = warp_sum(s) / static_cast<float>(D);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < kIt; ++i) {
      if (lane + 32 * i < nchunks) {
#pragma unroll
        for (int j = 0; j < kW; ++j) {
          const float d = v[i][j] - mean;
          q = fmaf(d, d, q);
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < kIt; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float g[kW], b[kW], o[kW];
#pragma unroll
        for (int j4 = 0; j4 < kW / 4; ++j4) {
          const float4 gv = __ldg(reinterpret_cast<const float4*>(gamma + c * kW) + j4);
          const float4 bv = __ldg(reinterpret_cast<const float4*>(beta + c * kW) + j4);
          g[4 * j4] = gv.x; g[4 * j4 + 1] = gv.y; g[4 * j4 + 2] = gv.z; g[4 * j4 + 3] = gv.w;
          b[4 * j4] = bv.x; b[4 * j4 + 1] = bv.y; b[4 * j4 + 2] = bv.z; b[4 * j4 + 3] = bv.w;
        }
#pragma unroll
        for (int j = 0; j < kW; ++j) o[j] = fmaf((v[i][j] - mean) * rstd, g[j], b[j]);
        reinterpret_cast<Vec*>(y + static_cast<size_t>(row) * D)[c] = ln_pack_w<kW>(o);
      }
    }
  }
}

int launch_layernorm_fwd(const __nv_bfloat16* x, const float* gamma, const float* beta, __nv_bfloat16* y,
                         float* mean, float* rstd, int M, int D, float eps, cudaStream_t stream) {
  if (M <= 0 || D <= 0 || (D & 7) || D > kLnMaxChunks * 256) return WM_ERR_SHAPE;
  if ((g_ln_fwd_width ? g_ln_fwd_width : ln_pick_width(D)) == 4) layernorm_fwd_kernel<4><<<(M + 8 * kLnFwdRows - 1) / (8 * kLnFwdRows), 256, 0, stream>>>(x, gamma, beta, y, mean, rstd, M, D, eps);
  else layernorm_fwd_kernel<8><<<(M + 8 * kLnFwdRows - 1) / (8 * kLnFwdRows), 256, 0, stream>>>(x, gamma, beta, y, mean, rstd, M, D, eps);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// LayerNorm backward. Persistent: one CTA per SM, 8 consumer warps + 1 producer warp. The producer streams blocks of
// 8 consecutive rows of x and dy (one contiguous span each: the activations are token-major) into a shared-memory
// ring with 1-D bulk copies (cp.async.bulk, mbarrier completion); consumer warp w owns row w of every block and
// keeps per-column partial sums of dgamma, dbeta and the bias gradient of the producing linear in registers. CTA
// partials go to a workspace and a second kernel folds them in a fixed order (deterministic).
// Why the ring: the column partials cost ~150 registers per thread, so only 8 consumer warps fit on an SM, and with
// register prefetch (the previous version: two rows ahead = 37 KB in flight per SM) Little's law capped the kernel
// at ~3.1 TB/s. The ring keeps up to 8 blocks x 2 tensors x 8 rows (147 KB at D = 576) in flight per SM.
constexpr int kLnBwdCtas = 148 * 4;  // workspace bound (the kernel launches min(row blocks, SMs) CTAs)
constexpr int kLnBwdMaxStages = 8;
// kBias: also accumulate the bias gradient of the producing linear (column sums of the bf16-rounded dx / dx_drop).
// The encoder takes that gradient from the wgrad GEMM instead (its fused all-ones chunk), which frees 24 registers
// per thread: 15 consumer warps (rows per block) fit instead of 8 -- the kernel is issue/latency-bound, so warps count.
template <bool kBias, int kRows, int kW>
__global__ void __launch_bounds__(32 * (kRows + 1), 1)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                     const float* __restrict__ gamma, const float* __restrict__ mean,
                     const float* __restrict__ rstd, __nv_bfloat16* __restrict__ dx,
                     __nv_bfloat16* __restrict__ dx_drop, int M, int D, uint32_t drop_thresh, float drop_scale,
                     DropKeys dkeys_in, float* __restrict__ partial, int stages) {
  const DropKeys dkeys = drop_keys_live(dkeys_in);  // (+ the per-replay words of a captured step; zero otherwise)
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const uint32_t row_bytes = static_cast<uint32_t>(D) * 2u;
  const uint32_t tens_bytes = static_cast<uint32_t>(kRows) * row_bytes;
  const uint32_t stage_bytes = 2u * tens_bytes;  // [x rows | dy rows]
  uint8_t* ring = ln_smem;
  constexpr int kSums = kBias ? 3 : 2;
  float* sred = reinterpret_cast<float*>(ring + static_cast<size_t>(stages) * stage_bytes);  // [kRows warps][kSums][D]
  float* sgam = sred + kRows * kSums * D;  // gamma (fp32): read per row from here instead of 24 registers per thread
  uint64_t* full = reinterpret_cast<uint64_t*>(sgam + D);
  uint64_t* empty = full + stages;
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  using Vec = typename LnVec<kW>::T;
  constexpr int kIt = (kLnMaxChunks * 8) / kW;  // lane passes for the widest supported row
  const int nchunks = D / kW;
  const int nblk = (M + kRows - 1) / kRows;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kRows);  // one arrive per consumer warp
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < D; i += blockDim.x) sgam[i] = gamma[i];
  __syncthreads();

  if (warp == kRows) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int b = blockIdx.x; b < nblk; b += gridDim.x) {
        mbar_wait(&empty[s], ph ^ 1u, 41);
        const int r0 = b * kRows;
        const uint32_t bytes = static_cast<uint32_t>(min(kRows, M - r0)) * row_bytes;
        uint8_t* dst = ring + static_cast<size_t>(s) * stage_bytes;
        mbar_arrive_expect_tx(&full[s], 2u * bytes);
        bulk_load_1d(dst, x + static_cast<size_t>(r0) * D, bytes, &full[s]);
        bulk_load_1d(dst + tens_bytes, dy + static_cast<size_t>(r0) * D, bytes, &full[s]);
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // Packed fp32 arithmetic throughout (FADD2 / FMUL2 / FFMA2 on register pairs: even element low, odd element high):
    // ~11 instead of ~20 issued instructions per element. By itself that changed nothing (0.207 -> 0.209 ms at the large
    // shape): the ncu source view then showed what did hold the kernel back -- 18 spilled registers reloaded in the row
    // loop, and the row statistics' global-load latency exposed on every row (see below). With both gone: 0.146 ms,
    // 5.9 TB/s = 90 % of the measured copy bandwidth (profiles/r02_ln_bwd_v2.txt).
    constexpr int kP = kW / 2;
    uint64_t ag2[kIt][kP], ab2[kIt][kP];
    float ad[kBias ? kIt : 1][kW];
#pragma unroll
    for (int i = 0; i < kIt; ++i) {
      const int c = lane + 32 * i;
#pragma unroll
      for (int j = 0; j < kP; ++j) {
        ag2[i][j] = f2_pack(0.0f, 0.0f);
        ab2[i][j] = f2_pack(0.0f, 0.0f);
      }
      if constexpr (kBias) {
#pragma unroll
        for (int j = 0; j < kW; ++j) ad[i][j] = 0.0f;
      }
    }
    const float invD = 1.0f / static_cast<float>(D);
    const uint64_t sc2 = f2_pack(drop_scale, drop_scale);
    const uint32_t add2 = drop_add2(drop_thresh);
    int s = 0;
    uint32_t ph = 0;
    // Row statistics: lane l holds those of this warp's row 32 * (it / 32) + l blocks ahead, one shuffle per row hands
    // them out. (Loaded one block ahead as scalars they are warp-uniform, the compiler moves them to uniform registers
    // right behind the load -- and that R2UR waited out the whole global-load latency on every row: 19 % of the
    // kernel's stall samples in the ncu source view.)
    float mu_v = 0.0f, rs_v = 0.0f;
    int it = 0;
    for (int b = blockIdx.x; b < nblk; b += gridDim.x, ++it) {
      const int row = b * kRows + warp;
      if ((it & 31) == 0) {
        const long long r = (static_cast<long long>(b) + static_cast<long long>(lane) * gridDim.x) * kRows + warp;
        mu_v = r < M ? __ldg(mean + r) : 0.0f;
        rs_v = r < M ? __ldg(rstd + r) : 0.0f;
      }
      const float mu = __shfl_sync(0xffffffffu, mu_v, it & 31), rs = __shfl_sync(0xffffffffu, rs_v, it & 31);
      mbar_wait(&full[s], ph, 42);
      const bool live = row < M;  // warp-uniform (tail block)
      uint64_t xh2[kIt][kP], gd2[kIt][kP];
      uint64_t s1p = f2_pack(0.0f, 0.0f), s2p = f2_pack(0.0f, 0.0f);
      const uint64_t nmu2 = f2_pack(-mu, -mu), rs2 = f2_pack(rs, rs);
      {
        // (rows are read from the ring pass by pass, not all up front: the raw vectors of all passes next to the
        // products built from them were the registers ptxas spilled -- the reloads showed up as long-scoreboard stalls)
        const Vec* sx = reinterpret_cast<const Vec*>(ring + static_cast<size_t>(s) * stage_bytes + warp * row_bytes);
        const Vec* sd = reinterpret_cast<const Vec*>(ring + static_cast<size_t>(s) * stage_bytes + tens_bytes + warp * row_bytes);
#pragma unroll
        for (int i = 0; i < kIt; ++i) {
          const int c = lane + 32 * i;
          if (c < nchunks && live) {
            const Vec xr = sx[c], dr = sd[c];
            const uint32_t* xw = reinterpret_cast<const uint32_t*>(&xr);
            const uint32_t* dw = reinterpret_cast<const uint32_t*>(&dr);
            float gv[kW];
#pragma unroll
            for (int j4 = 0; j4 < kW / 4; ++j4) {
              const float4 g4 = reinterpret_cast<const float4*>(sgam + c * kW)[j4];
              gv[4 * j4] = g4.x; gv[4 * j4 + 1] = g4.y; gv[4 * j4 + 2] = g4.z; gv[4 * j4 + 3] = g4.w;
            }
#pragma unroll
            for (int j = 0; j < kP; ++j) {
              const uint64_t d2 = f2_pack(bf16_lo(dw[j]), bf16_hi(dw[j]));
              const uint64_t xh = f2_mul(f2_add(f2_pack(bf16_lo(xw[j]), bf16_hi(xw[j])), nmu2), rs2);
              const uint64_t gd = f2_mul(d2, f2_pack(gv[2 * j], gv[2 * j + 1]));
              s1p = f2_add(s1p, gd);
              s2p = f2_fma(gd, xh, s2p);
              ag2[i][j] = f2_fma(d2, xh, ag2[i][j]);
              ab2[i][j] = f2_add(ab2[i][j], d2);
              xh2[i][j] = xh;
              gd2[i][j] = gd;
            }
          }
        }
      }
      __syncwarp();  // every lane has read its part: hand the slot back to the producer
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == stages) { s = 0; ph ^= 1u; }
      if (!live) continue;
      float s1, s2;
      {
        float a0, a1, b0, b1;
        f2_unpack(s1p, a0, a1);
        f2_unpack(s2p, b0, b1);
        s1 = warp_sum(a0 + a1) * invD;
        s2 = warp_sum(b0 + b1) * invD;
      }
      const uint64_t ns1 = f2_pack(-s1, -s1), ns2 = f2_pack(-s2, -s2);
#pragma unroll
      for (int i = 0; i < kIt; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunks) {
          uint64_t o2[kP];
          Vec pk;
          uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
          for (int j = 0; j < kP; ++j) {  // rs * (dy * gamma - s1 - xhat * s2)
            o2[j] = f2_mul(f2_fma(xh2[i][j], ns2, f2_add(gd2[i][j], ns1)), rs2);
            float lo, hi;
            f2_unpack(o2[j], lo, hi);
            pw[j] = pack_bf16x2(lo, hi);
          }
          reinterpret_cast<Vec*>(dx + static_cast<size_t>(row) * D)[c] = pk;
          if (drop_thresh) {  // the mask the GEMM epilogue drew for these columns: counter = (row, 16-column group, 4-element word)
            const int col = c * kW;
            const uint32_t x0 = (static_cast<uint32_t>(row) * static_cast<uint32_t>((D + 15) >> 4) + static_cast<uint32_t>(col >> 4)) * 4u +
                                static_cast<uint32_t>((col & 15) >> 2);
#pragma unroll
            for (int w = 0; w < kW / 4; ++w) {  // scale in fp32, round, then mask the packed pair (same bits as masking first)
              const DropWords fl = drop_flags4(x0 + w, dkeys, add2);
              float a0, a1, a2, a3;
              f2_unpack(f2_mul(o2[2 * w], sc2), a0, a1);
              f2_unpack(f2_mul(o2[2 * w + 1], sc2), a2, a3);
              pw[2 * w] = pack_bf16x2(a0, a1) & drop_pair_mask(fl.a);
              pw[2 * w + 1] = pack_bf16x2(a2, a3) & drop_pair_mask(fl.b);
            }
            reinterpret_cast<Vec*>(dx_drop + static_cast<size_t>(row) * D)[c] = pk;
          }
          // bias gradient of the producing linear sums what that linear's output actually received;
          // use the bf16-rounded values so it matches the wgrad operand exactly
          if constexpr (kBias) {
#pragma unroll
            for (int j = 0; j < kW / 2; ++j) {
              ad[i][2 * j] += bf16_lo(pw[j]);
              ad[i][2 * j + 1] += bf16_hi(pw[j]);
            }
          }
        }
      }
    }
    // CTA reduce: warp w writes its registers, then the CTA folds the row warps per column
#pragma unroll
    for (int i = 0; i < kIt; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
#pragma unroll
        for (int j = 0; j < kP; ++j) {
          float lo, hi;
          f2_unpack(ag2[i][j], lo, hi);
          sred[(warp * kSums + 0) * D + c * kW + 2 * j] = lo;
          sred[(warp * kSums + 0) * D + c * kW + 2 * j + 1] = hi;
          f2_unpack(ab2[i][j], lo, hi);
          sred[(warp * kSums + 1) * D + c * kW + 2 * j] = lo;
          sred[(warp * kSums + 1) * D + c * kW + 2 * j + 1] = hi;
        }
        if constexpr (kBias) {
#pragma unroll
          for (int j = 0; j < kW; ++j) sred[(warp * kSums + 2) * D + c * kW + j] = ad[i][j];
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSums * D; i += blockDim.x) {
    float sum = 0.0f;
#pragma unroll
    for (int w = 0; w < kRows; ++w) sum += sred[w * kSums * D + i];
    partial[static_cast<size_t>(blockIdx.x) * 3 * D + i] = sum;
  }
}

// out_k[c] (+)= sum_{cta} partial[cta][k][c]: 64 columns per CTA, four threads per column each folding a quarter of
// the CTAs in order, then a fixed-order fold of the four
__global__ void __launch_bounds__(256)
ln_bwd_finalize_kernel(const float* __restrict__ partial, int ncta, int D, float* dgamma, float* dbeta, float* dbias) {
  __shared__ float part[4][64];
  const int i = blockIdx.x * 64 + (threadIdx.x & 63);
  const int q = threadIdx.x >> 6;
  const int per = (ncta + 3) >> 2;
  const int nsum = dbias ? 3 : 2;  // without dbias the main kernel wrote only the first two [D] blocks of each partial
  float s = 0.0f;
  if (i < nsum * D)
    for (int c = q * per; c < min(ncta, (q + 1) * per); ++c) s += partial[static_cast<size_t>(c) * 3 * D + i];
  part[q][threadIdx.x & 63] = s;
  __syncthreads();
  if (q == 0 && i < nsum * D) {
    const int t = threadIdx.x;
    const float sum = ((part[0][t] + part[1][t]) + part[2][t]) + part[3][t];
    const int k = i / D, col = i - k * D;
    float* dst = k == 0 ? dgamma : k == 1 ? dbeta : dbias;
    if (dst) dst[col] = sum;
  }
}

size_t layernorm_bwd_workspace_bytes(int M, int D) {
  (void)M;
  return static_cast<size_t>(kLnBwdCtas) * 3 * D * sizeof(float);
}

int launch_layernorm_bwd(const __nv_bfloat16* dy, const __nv_bfloat16* x, const float* gamma, const float* mean,
                         const float* rstd, __nv_bfloat16* dx, __nv_bfloat16* dx_drop, float* dgamma,
                         float* dbeta, float* dbias, int M, int D, uint32_t drop_thresh, float drop_scale,
                         uint64_t seed, uint64_t stream_id, float* workspace, cudaStream_t stream) {
  if (M <= 0 || D <= 0 || (D & 7) || D > kLnMaxChunks * 256) return WM_ERR_SHAPE;
  if (drop_thresh && !dx_drop) return WM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) |
       reinterpret_cast<uintptr_t>(dx_drop)) & 15u)
    return WM_ERR_ALIGN;  // bulk copies and 16-byte stores
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int rows = dbias ? 8 : (g_ln_bwd_rows == 14 ? 14 : 15);  // (encoder form: 15 row warps at 122 registers, no spills)
  const int nsum = dbias ? 3 : 2;
  int ctas = (M + rows - 1) / rows;
  if (ctas > sms) ctas = sms;
  if (ctas > kLnBwdCtas) ctas = kLnBwdCtas;
  const int stage_bytes = 2 * rows * D * 2;
  const int fixed = rows * nsum * D * 4 + D * 4 + 2 * kLnBwdMaxStages * 8 + 128;
  int stages = (227 * 1024 - fixed) / stage_bytes;
  if (stages > kLnBwdMaxStages) stages = kLnBwdMaxStages;
  if (stages < 2) return WM_ERR_SHAPE;
  const int smem = stages * stage_bytes + fixed;
  if (drop_thresh && static_cast<uint64_t>(M) * static_cast<uint64_t>((D + 15) / 16) * 4ull > 0xFFFFFFFFull) return WM_ERR_SHAPE;
  const bool w4 = g_ln_bwd_width ? g_ln_bwd_width == 4 : false;  // ("ln_bwd_width" option; A/B in profiles/r02_ln_width_ab.txt)
  auto kern = dbias ? (w4 ? layernorm_bwd_kernel<true, 8, 4> : layernorm_bwd_kernel<true, 8, 8>)
                    : rows == 15 ? (w4 ? layernorm_bwd_kernel<false, 15, 4> : layernorm_bwd_kernel<false, 15, 8>)
                                 : (w4 ? layernorm_bwd_kernel<false, 14, 4> : layernorm_bwd_kernel<false, 14, 8>);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return WM_ERR_CUDA;
  kern<<<ctas, 32 * (rows + 1), smem, stream>>>(dy, x, gamma, mean, rstd, dx, dx_drop, M, D, drop_thresh, drop_scale,
                                                drop_keys(seed, stream_id), workspace, stages);
  WM_COUNT_LAUNCH();
  if (cudaGetLastError() != cudaSuccess) return WM_ERR_CUDA;
  ln_bwd_finalize_kernel<<<(3 * D + 63) / 64, 256, 0, stream>>>(workspace, ctas, D, dgamma, dbeta, dbias);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// colsum: out[n] = sum_m x[m, n] (bias gradients). blockDim (32, 8); each x-thread owns 8 columns.
// ------------------------------------------------------------------------------------------------
constexpr int kColsumSlabs = 296;

__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, int ld, int M, int N, float* __restrict__ partial) {
  __shared__ float sred[8][32][9];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;  // 8-column chunk index
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c * 8 < N) {
    for (int r = r0 + ty; r < r1; r += 8) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * ld) + c);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] += bf16_lo(aw[j]);
        acc[2 * j + 1] += bf16_hi(aw[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sred[ty][tx][j] = acc[j];
  __syncthreads();
  if (ty == 0 && c * 8 < N) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += sred[w][tx][j];
      partial[static_cast<size_t>(blockIdx.y) * N + c * 8 + j] = s;
    }
  }
}
__global__ void colsum_finalize_kernel(const float* __restrict__ partial, int slabs, int N, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float s = 0.0f;
  for (int k = 0; k < slabs; ++k) s += partial[static_cast<size_t>(k) * N + i];
  out[i] = s;
}
static int colsum_slabs(int M, int N) {
  const int gx = (N / 8 + 31) / 32;
  int slabs = (kColsumSlabs * 4 + gx - 1) / gx;
  if (slabs > (M + 7) / 8) slabs = (M + 7) / 8;
  return slabs < 1 ? 1 : slabs;
}
size_t colsum_workspace_bytes(int M, int N) { return static_cast<size_t>(colsum_slabs(M, N)) * N * sizeof(float); }

int launch_colsum(const __nv_bfloat16* x, int ld, int M, int N, float* out, float* workspace, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || (N & 7) || (ld & 7)) return WM_ERR_SHAPE;
  const int slabs = colsum_slabs(M, N);
  dim3 grid((N / 8 + 31) / 32, slabs);
  colsum_kernel<<<grid, 256, 0, stream>>>(x, ld, M, N, workspace);
  WM_COUNT_LAUNCH();
  if (cudaGetLastError() != cudaSuccess) return WM_ERR_CUDA;
  colsum_finalize_kernel<<<(N + 255) / 256, 256, 0, stream>>>(workspace, slabs, N, out);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// Loss heads. Phase 1: per-CTA partial sums -> scratch; phase 2: every CTA folds the partials in a
// fixed order (deterministic), CTA 0 publishes the scalars, all CTAs write dY (bf16, padded ld).
// scratch layout: [kLossCtas][4] floats.
// ------------------------------------------------------------------------------------------------
constexpr int kLossCtas = 592;

WM_DEVICE void block_reduce4(float (&v)[4], float* sbuf /* [8][4] */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) sbuf[warp * 4 + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float s = 0.0f;
    for (int w = 0; w < 8; ++w) s += sbuf[w * 4 + k];
    v[k] = s;
  }
  __syncthreads();
}
// every thread folds the CTA partials in the same fixed order
WM_DEVICE void fold_partials(const float* scratch, int nctas, float (&tot)[4], float* sbuf) {
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < nctas; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] += scratch[i * 4 + k];
  }
  block_reduce4(v, sbuf);
#pragma unroll
  for (int k = 0; k < 4; ++k) tot[k] = v[k];
}

__global__ void __launch_bounds__(256)
loss_bert_partial_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ weather,
                         const uint8_t* __restrict__ mask, int64_t M, int F, float* __restrict__ scratch) {
  __shared__ float sbuf[32];
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  const int64_t n = M * F;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (mask[i]) {
      const int64_t t = i / F;
      const int f = static_cast<int>(i - t * F);
      const float d = weather[i] - y[t * ldy + f];
      v[0] = fmaf(d, d, v[0]);
      v[1] += 1.0f;
    }
  }
  block_reduce4(v, sbuf);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) scratch[blockIdx.x * 4 + k] = v[k];
  }
}
__global__ void __launch_bounds__(256)
loss_bert_grad_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ weather,
                      const uint8_t* __restrict__ mask, int64_t M, int F, const float* __restrict__ scratch,
                      int nctas, float* __restrict__ loss_out, const float* __restrict__ gscale,
                      __nv_bfloat16* __restrict__ dy, int lddy) {
  __shared__ float sbuf[32];
  float tot[4];
  fold_partials(scratch, nctas, tot, sbuf);
  float inv = 1.0f / tot[1];  // NaN/Inf if nothing is masked, exactly as the reference's mean over []
  if (loss_out && blockIdx.x == 0 && threadIdx.x == 0) {
    loss_out[0] = tot[0] * inv;
    loss_out[1] = tot[1];
  }
  if (!dy) return;
  if (gscale) inv *= __ldg(gscale);  // upstream gradient of the scalar loss (a device scalar: no host read-back)
  const int64_t n = M * lddy;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / lddy;
    const int f = static_cast<int>(i - t * lddy);
    float gval = 0.0f;
    if (f < F && mask[t * F + f]) gval = 2.0f * (y[t * ldy + f] - weather[t * F + f]) * inv;
    dy[i] = __float2bfloat16(gval);
  }
}

// loss_out != nullptr: partial sums -> scratch, then the fold (+ dy if given). loss_out == nullptr: dy only, from the
// partial sums an earlier call left in scratch (the autograd backward of the loss: dy = grad_scale[0] * dLoss/dY).
int launch_loss_bert(const float* y, int ldy, const float* weather, const uint8_t* mask, int64_t M, int F,
                     float* scratch, float* loss_out, const float* grad_scale, __nv_bfloat16* dy, int lddy,
                     cudaStream_t stream) {
  if (M <= 0 || F <= 0) return WM_ERR_SHAPE;
  if (!loss_out && !dy) return WM_ERR_ARG;
  int ctas = static_cast<int>((M * F + 255) / 256);
  if (ctas > kLossCtas) ctas = kLossCtas;
  if (loss_out) {
    loss_bert_partial_kernel<<<ctas, 256, 0, stream>>>(y, ldy, weather, mask, M, F, scratch);
    WM_COUNT_LAUNCH();
    if (cudaGetLastError() != cudaSuccess) return WM_ERR_CUDA;
  }
  loss_bert_grad_kernel<<<dy ? kLossCtas : 1, 256, 0, stream>>>(y, ldy, weather, mask, M, F, scratch, ctas, loss_out,
                                                               grad_scale, dy, lddy);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// WeatherFormer ELBO. y[t, 0:F] = mu, y[t, F:2F] = log-variance.
struct FormerTerms {
  float nll, kl, dmu_r, dmu_k, dl_r, dl_k;
};
WM_DEVICE FormerTerms former_terms(float x, float mu, float lv) {
  const float e = expf(lv);
  const float v = fminf(fmaxf(e, 1e-6f), 1.0f);
  const float d = x - mu;
  FormerTerms t;
  t.nll = 0.5f * logf(2.0f * 3.14159265358979323846f * v) + 0.5f * d * d / v;
  t.kl = 0.5f * (logf(1.0f / v) + v + mu * mu - 1.0f);
  const float gate = (e >= 1e-6f && e <= 1.0f) ? e : 0.0f;  // d clamp(exp(l)) / dl
  t.dmu_r = -d / v;
  t.dmu_k = mu;
  t.dl_r = (0.5f / v - 0.5f * d * d / (v * v)) * gate;
  t.dl_k = 0.5f * (1.0f - 1.0f / v) * gate;
  return t;
}

__global__ void __launch_bounds__(256)
loss_former_partial_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ weather,
                           const uint8_t* __restrict__ mask, int64_t msb, int64_t mss, int B, int S, int F,
                           float* __restrict__ scratch, float* __restrict__ mu_out, float* __restrict__ var_out) {
  __shared__ float sbuf[32];
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  const int64_t n = static_cast<int64_t>(B) * S * F;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / F;
    const int f = static_cast<int>(i - t * F);
    const int64_t b = t / S, s = t - b * S;
    const float mu = y[t * ldy + f], lv = y[t * ldy + F + f];
    if (mu_out) {
      mu_out[i] = mu;
      var_out[i] = fminf(fmaxf(expf(lv), 1e-6f), 1.0f);
    }
    if (mask[b * msb + s * mss + f]) {
      const FormerTerms tm = former_terms(weather[i], mu, lv);
      v[0] += tm.nll;
      v[1] += tm.kl;
      v[2] += 1.0f;
    }
  }
  block_reduce4(v, sbuf);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) scratch[blockIdx.x * 4 + k] = v[k];
  }
}
__global__ void __launch_bounds__(256)
loss_former_grad_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ weather,
                        const uint8_t* __restrict__ mask, int64_t msb, int64_t mss, int B, int S, int F,
                        float beta, const float* __restrict__ scratch, int nctas, float* __restrict__ loss_out,
                        const float* __restrict__ gscale, __nv_bfloat16* __restrict__ dy, int lddy) {
  __shared__ float sbuf[32];
  float tot[4];
  fold_partials(scratch, nctas, tot, sbuf);
  // n_bar = tot[2] / B ; recon = (1/B) sum_b nll_b / n_bar = tot[0] / tot[2] ; kl likewise * beta
  float inv = 1.0f / tot[2];
  if (loss_out && blockIdx.x == 0 && threadIdx.x == 0) {
    const float recon = tot[0] * inv, kl = beta * tot[1] * inv;
    loss_out[0] = recon + kl;
    loss_out[1] = recon;
    loss_out[2] = kl;
    loss_out[3] = tot[2];
  }
  if (!dy) return;
  if (gscale) inv *= __ldg(gscale);  // upstream gradient of the scalar loss (device scalar)
  const int64_t M = static_cast<int64_t>(B) * S;
  const int64_t n = M * F;
  // zero the padding columns [2F, lddy)
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < M * (lddy - 2 * F);
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / (lddy - 2 * F);
    dy[t * lddy + 2 * F + (i - t * (lddy - 2 * F))] = __float2bfloat16(0.0f);
  }
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / F;
    const int f = static_cast<int>(i - t * F);
    const int64_t b = t / S, s = t - b * S;
    float gm = 0.0f, gl = 0.0f;
    if (mask[b * msb + s * mss + f]) {
      const FormerTerms tm = former_terms(weather[i], y[t * ldy + f], y[t * ldy + F + f]);
      gm = (tm.dmu_r + beta * tm.dmu_k) * inv;
      gl = (tm.dl_r + beta * tm.dl_k) * inv;
    }
    dy[t * lddy + f] = __float2bfloat16(gm);
    dy[t * lddy + F + f] = __float2bfloat16(gl);
  }
}

int launch_loss_former(const float* y, int ldy, const float* weather, const uint8_t* mask, int64_t msb,
                       int64_t mss, int B, int S, int F, float beta, float* scratch, float* loss_out,
                       const float* grad_scale, __nv_bfloat16* dy, int lddy, float* mu_out, float* var_out,
                       cudaStream_t stream) {
  if (B <= 0 || S <= 0 || F <= 0 || (dy && lddy < 2 * F)) return WM_ERR_SHAPE;
  if (!loss_out && !dy) return WM_ERR_ARG;
  const int64_t n = static_cast<int64_t>(B) * S * F;
  int ctas = static_cast<int>((n + 255) / 256);
  if (ctas > kLossCtas) ctas = kLossCtas;
  if (loss_out) {  // (see launch_loss_bert: loss_out == nullptr is the gradient-only second call)
    loss_former_partial_kernel<<<ctas, 256, 0, stream>>>(y, ldy, weather, mask, msb, mss, B, S, F, scratch, mu_out, var_out);
    WM_COUNT_LAUNCH();
    if (cudaGetLastError() != cudaSuccess) return WM_ERR_CUDA;
  }
  loss_former_grad_kernel<<<dy ? kLossCtas : 1, 256, 0, stream>>>(y, ldy, weather, mask, msb, mss, B, S, F, beta, scratch,
                                                                 ctas, loss_out, grad_scale, dy, lddy);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam, amsgrad=False, maximize=False). Also refreshes the bf16 shadow weights.
// ------------------------------------------------------------------------------------------------
// One Adam update of four consecutive elements (float4 accesses: the scalar version ran at 2.8 TB/s, 42 % of the HBM
// roofline, on seven streams of 4-byte accesses); the bf16 shadow of the new parameters is written from the same registers.
WM_DEVICE void adam_update4(float4& p4, const float4& g4, float4& m4, float4& v4, float omb1, float beta2, float omb2, float eps,
                            float weight_decay, float step_size, float bc2_sqrt, float grad_scale) {
  float* p = &p4.x; const float* g = &g4.x; float* m = &m4.x; float* v = &v4.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float gi = g[k] * grad_scale;
    float pi = p[k];
    if (weight_decay != 0.0f) gi = fmaf(weight_decay, pi, gi);
    float mi = m[k], vi = v[k];
    mi = mi + omb1 * (gi - mi);                      // exp_avg.lerp_(grad, 1 - beta1), weight < 0.5
    vi = vi * beta2 + omb2 * gi * gi;                // mul_(beta2).addcmul_(g, g, 1 - beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);              // addcdiv_(exp_avg, denom, value=-step_size)
    p[k] = pi; m[k] = mi; v[k] = vi;
  }
}
WM_DEVICE void adam_body(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                         __nv_bfloat16* __restrict__ shadow, int64_t n, float lr, float omb1, float beta2, float omb2, float eps,
                         float weight_decay, float bc1, float bc2_sqrt, float grad_scale) {
  // omb1 = float(1 - beta1), omb2 = float(1 - beta2) formed in double on the host, as torch forms its lerp / addcmul weights
  const float step_size = lr / bc1;
  const int64_t n4 = n >> 2;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15u) == 0 && (reinterpret_cast<uintptr_t>(shadow) & 7u) == 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x, tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (vec) {
    for (int64_t i = tid; i < n4; i += stride) {
      float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
      const float4 g4 = reinterpret_cast<const float4*>(g)[i];
      adam_update4(p4, g4, m4, v4, omb1, beta2, omb2, eps, weight_decay, step_size, bc2_sqrt, grad_scale);
      reinterpret_cast<float4*>(p)[i] = p4;
      reinterpret_cast<float4*>(m)[i] = m4;
      reinterpret_cast<float4*>(v)[i] = v4;
      if (shadow) reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16x2(p4.x, p4.y), pack_bf16x2(p4.z, p4.w));
    }
  }
  for (int64_t i = (vec ? (n4 << 2) : 0) + tid; i < n; i += stride) {  // tail (or everything, if unaligned)
    float gi = g[i] * grad_scale;
    float pi = p[i];
    if (weight_decay != 0.0f) gi = fmaf(weight_decay, pi, gi);
    float mi = m[i], vi = v[i];
    mi = mi + omb1 * (gi - mi);
    vi = vi * beta2 + omb2 * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (shadow) shadow[i] = __float2bfloat16(pi);
  }
}
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            __nv_bfloat16* __restrict__ shadow, int64_t n, float lr, float omb1, float beta2, float omb2, float eps,
            float weight_decay, float bc1, float bc2_sqrt, float grad_scale) {
  adam_body(p, g, m, v, shadow, n, lr, omb1, beta2, omb2, eps, weight_decay, bc1, bc2_sqrt, grad_scale);
}

// The same update with the step-dependent scalars read from DEVICE memory: hyper = {lr, 1 - beta1^t, sqrt(1 - beta2^t)}.
// A captured training step (CUDA graph) replays this launch unchanged while the host refreshes the three numbers.
__global__ void __launch_bounds__(256)
adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                __nv_bfloat16* __restrict__ shadow, int64_t n, const float* __restrict__ hyper, float omb1, float beta2,
                float omb2, float eps, float weight_decay, float grad_scale) {
  adam_body(p, g, m, v, shadow, n, hyper[0], omb1, beta2, omb2, eps, weight_decay, hyper[1], hyper[2], grad_scale);
}
int launch_adam_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, __nv_bfloat16* shadow, int64_t n,
                    const float* hyper_dev, double beta1, double beta2, float eps, float weight_decay, float grad_scale,
                    cudaStream_t stream) {
  if (n <= 0) return WM_OK;
  if (!hyper_dev) return WM_ERR_ARG;
  int blocks = static_cast<int>((n / 4 + 255) / 256) + 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_dev_kernel<<<blocks, 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, shadow, n, hyper_dev,
                                              static_cast<float>(1.0 - beta1), static_cast<float>(beta2),
                                              static_cast<float>(1.0 - beta2), eps, weight_decay, grad_scale);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

__global__ void step_params_apply_kernel(const uint32_t* __restrict__ words) {
  if (threadIdx.x < 3) g_wm_drop_mix[threadIdx.x] = words ? words[threadIdx.x] : 0u;
}
int launch_step_params_apply(const uint32_t* dev_words, cudaStream_t stream) {
  step_params_apply_kernel<<<1, 32, 0, stream>>>(dev_words);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

int launch_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, __nv_bfloat16* shadow,
                int64_t n, float lr, double beta1, double beta2, float eps, float weight_decay, int step,
                float grad_scale, cudaStream_t stream) {
  if (n <= 0) return WM_OK;
  if (step < 1) return WM_ERR_ARG;
  const double bc1 = 1.0 - pow(beta1, step);
  const double bc2 = 1.0 - pow(beta2, step);
  int blocks = static_cast<int>((n / 4 + 255) / 256) + 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<blocks, 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, shadow, n, lr, static_cast<float>(1.0 - beta1),
                                          static_cast<float>(beta2), static_cast<float>(1.0 - beta2), eps, weight_decay,
                                          static_cast<float>(bc1), static_cast<float>(sqrt(bc2)), grad_scale);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// wt[c, r] = bf16(w[r, c]); wt row pitch ld_out >= rows, pad columns zeroed by the caller (memset once)
__global__ void __launch_bounds__(256)
cast_transpose_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wt, int rows, int cols, int ld_out) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < rows && c < cols) ? w[static_cast<size_t>(r) * cols + c] : 0.0f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (c < cols && r < rows) wt[static_cast<size_t>(c) * ld_out + r] = __float2bfloat16(tile[tx][j]);
  }
}
// All transposed bf16 weight copies of an encoder in ONE launch (4 per layer + the head: 33 launches per step at
// WeatherFormer large were 33 x ~6 us of kernel time plus as many launch gaps, and a third of the nodes of a recorded
// mini step). The job table travels as a kernel parameter; a block finds its job by scanning the cumulative tile counts.
__global__ void __launch_bounds__(256)
cast_transpose_multi_kernel(const TransposeJobs jobs) {
  __shared__ float tile[32][33];
  int j = 0;
  while (j + 1 < jobs.n && static_cast<int>(blockIdx.x) >= jobs.job[j + 1].tile0) ++j;
  const TransposeJob& J = jobs.job[j];
  const int local = blockIdx.x - J.tile0;
  const int tiles_x = (J.cols + 31) / 32;
  const int bx = local % tiles_x, by = local / tiles_x;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = by * 32, c0 = bx * 32;
  for (int k = ty; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + tx;
    tile[k][tx] = (r < J.rows && c < J.cols) ? J.w[static_cast<size_t>(r) * J.cols + c] : 0.0f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + tx;
    if (c < J.cols && r < J.rows) J.wt[static_cast<size_t>(c) * J.ld_out + r] = __float2bfloat16(tile[tx][k]);
  }
}
int launch_cast_transpose_multi(TransposeJobs& jobs, cudaStream_t stream) {
  if (jobs.n <= 0 || jobs.n > kMaxTransposeJobs) return WM_ERR_ARG;
  int tiles = 0;
  for (int j = 0; j < jobs.n; ++j) {
    TransposeJob& J = jobs.job[j];
    if (J.rows <= 0 || J.cols <= 0 || J.ld_out < J.rows) return WM_ERR_SHAPE;
    J.tile0 = tiles;
    tiles += ((J.cols + 31) / 32) * ((J.rows + 31) / 32);
  }
  cast_transpose_multi_kernel<<<tiles, 256, 0, stream>>>(jobs);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

int launch_cast_transpose(const float* w, __nv_bfloat16* wt, int rows, int cols, int ld_out, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0 || ld_out < rows) return WM_ERR_SHAPE;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  cast_transpose_kernel<<<grid, 256, 0, stream>>>(w, wt, rows, cols, ld_out);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}
int launch_cast_bf16(const float* src, __nv_bfloat16* dst, int64_t n, cudaStream_t stream) {
  if (n <= 0) return WM_OK;
  int blocks = static_cast<int>((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cast_bf16_kernel<<<blocks, 256, 0, stream>>>(src, dst, n);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

}  // namespace wm
