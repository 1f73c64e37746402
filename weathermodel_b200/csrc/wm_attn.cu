// wm_attn.cu -- fused multi-head self-attention forward / backward for S <= 384 on tcgen05 (sm_100a).
//
// Replaces F.scaled_dot_product_attention(q, k, v, None, dropout_p, False) as reached from
// nn.TransformerEncoderLayer (torch/nn/functional.py:6666-6696; reference call site
// src/pretraining/models/weatherbert.py:45-54,116-118). Head dims of this model family are 12/20/28/36:
// Q/K/V rows are copied from the token-major [M, 3D] QKV activation into zero-padded, UNSWIZZLED
// canonical core-matrix tiles in shared memory (8 rows x 16 B per core matrix), which one and the same
// tile can feed to tcgen05.mma either K-major (contract over head dim) or MN-major (contract over rows).
//
// One CTA per (batch, head); thread t owns query row t of the current 128-row tile (TMEM lane t), so
// row max / row sum / LSE / delta are plain per-thread scalars -- no shuffles, no atomics.
//   fwd : pass 1 row max over 64-key score chunks, pass 2 exp2 + dropout + P(bf16)->smem, O += P V.
//   bwd : per (kv tile j, q tile i): S = Q K^T, dP = dO V^T in TMEM; P, dS -> smem (bf16);
//         dV_j += P^T dO, dK_j += dS^T Q, dQ_i += dS K; all five accumulators live in TMEM (<= 496 cols).
// Dropout on P uses Philox4x32-10 keyed by (seed, stream, (bh*S + q)*ceil(S/16) + k/16), 8 bits / element
// (keep iff byte >= thresh8), regenerated identically in backward.
#include "wm_kernels.h"

namespace wm {

constexpr int kAttThreads = 128;
constexpr int kSP = 384;  // key rows staged per head (S <= 384)

// canonical unswizzled tile [rows, DHP]: elem(r, d) at (r/8)*RS + (d/8)*128 + (r%8)*16 + (d%8)*2
template <int DHP>
struct TileGeom {
  static constexpr uint32_t RS = (DHP / 8) * 128;
};

// copy rows [0, S) of one head slice (row pitch ld elements, dh valid columns) into a canonical tile of
// `rows_alloc` rows, zero-filling pad columns and pad rows.
// Asynchronous (LDGSTS) 8-byte pieces: every thread queues all of its pieces back to back and the
// caller waits once (cp_async_wait_all) -- the first version used ld.global + st.shared per piece and paid one
// DRAM round trip per loop iteration (~36 per thread), which was most of the kernel time.
WM_DEVICE void cp_async8(void* smem_dst, const void* gmem_src, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int sz = valid ? 8 : 0;  // src-size 0 => destination is zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
WM_DEVICE void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Stage NT head slices ([S rows, dh] each, row pitch ld[t]) into canonical tiles.
//  phase 1: every thread issues ALL of its 8-byte global loads into registers. Indexing is warp-structured
//           (lane -> (row within the warp's row group, piece), rows advance by a constant) so a piece costs a
//           handful of integer instructions; the first versions spent ~10k cycles per CTA on index arithmetic
//           and per-piece zero fills (profiles/r01_attn_phase_ticks.txt).
//  phase 2: the padding (16-byte chunk columns at/after dh, rows >= S) is zeroed with 16-byte stores while the
//           loads are in flight;  phase 3: barrier, then the data pieces are stored (they overlap the first
//           zeroed chunk when dh % 8 == 4).
template <int DHP, int NT, int kIters>
WM_DEVICE void load_head_tiles(uint8_t* const (&tile)[NT], const __nv_bfloat16* const (&src)[NT], const int (&ld)[NT],
                               int S, int rows_alloc, int dh) {
  constexpr uint32_t RS = TileGeom<DHP>::RS;
  const int pv = dh >> 2;                   // valid 8-byte pieces per row
  const int rpi = 32 / pv;                  // rows covered by one warp instruction
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int rsub = lane / pv, p = lane - rsub * pv;
  const bool lane_on = rsub < rpi;
  const int rstep = rpi * nwarps;
  const uint32_t poff = (p >> 1) * 128 + (p & 1) * 8;
  uint2 v[NT][kIters];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const uint2* g = reinterpret_cast<const uint2*>(src[t] + static_cast<size_t>(warp * rpi + rsub) * ld[t]) + p;
    const size_t gstep = static_cast<size_t>(rstep) * ld[t] / 4;  // in uint2 units (ld % 4 == 0)
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int r = warp * rpi + rsub + it * rstep;
      if (lane_on && r < S) v[t][it] = __ldg(g + it * gstep);
    }
  }
  // zero the padding: chunk columns [dh / 8, DHP / 8) of every row, and whole rows [S, rows_alloc)
  const int cz = dh >> 3, nz = (DHP >> 3) - cz;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    for (int i = threadIdx.x; i < (rows_alloc >> 3) * nz * 8; i += blockDim.x) {
      const int g = i / (nz * 8), w = i - g * (nz * 8);  // w = chunk-in-pad * 8 + row-in-group
      *reinterpret_cast<uint4*>(tile[t] + g * RS + (cz + (w >> 3)) * 128 + (w & 7) * 16) = z4;
    }
    for (int i = threadIdx.x; i < (rows_alloc - S) * cz; i += blockDim.x) {
      const int rr = S + i / cz, c = i - (i / cz) * cz;
      *reinterpret_cast<uint4*>(tile[t] + (rr >> 3) * RS + c * 128 + (rr & 7) * 16) = z4;
    }
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < NT; ++t) {
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int r = warp * rpi + rsub + it * rstep;
      if (lane_on && r < S) *reinterpret_cast<uint2*>(tile[t] + (r >> 3) * RS + (r & 7) * 16 + poff) = v[t][it];
    }
  }
}

// Dropout on attention probabilities: one Philox4x32-7 block (the 7-round variant is the Crush-resistant
// minimum of the Random123 paper; nothing here has to match torch's stream) decides 16 consecutive keys of one
// query row, 8 bits each: keep iff byte >= thresh8. m[w] holds 0xFF in every kept byte of word w.
constexpr int kAttnPhiloxRounds = 7;
WM_DEVICE void keep_masks16(uint64_t seed, uint64_t stream, uint64_t grp, uint32_t thresh4, uint32_t (&m)[4]) {
  const Philox4 r = philox4x32<kAttnPhiloxRounds>(seed, stream, grp);
  m[0] = __vcmpgeu4(r.x, thresh4);
  m[1] = __vcmpgeu4(r.y, thresh4);
  m[2] = __vcmpgeu4(r.z, thresh4);
  m[3] = __vcmpgeu4(r.w, thresh4);
}
// pins a value in a register at this point of the instruction stream: without it the compiler sinks the (pure)
// Philox arithmetic below the mbarrier wait it is supposed to overlap with
#define WM_PIN(x) asm volatile("" : "+r"(x))
// all-ones / all-zeros 32-bit mask of element j (0..15) of the group
#define WM_KEEP32(m, j) __byte_perm((m)[(j) >> 2], 0u, 0x1111u * ((j) & 3))

WM_DEVICE void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
WM_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------------
// forward: 12 warps. Warp w reads TMEM lane quarter w%4 (query rows) and owns key tile w/4 (128 score
// columns) of the current 128-query tile. All scores of the tile sit in TMEM at once (384 columns), so
// the softmax is exact two-pass (max, then exp/sum) with one MMA round trip per tile; two issuing threads
// (PV of this tile, QK^T of the next) keep descriptor arithmetic off the critical path.
// ------------------------------------------------------------------------------------------------
constexpr int kFwdThreads = 416;  // 12 softmax warps + 1 MMA-issue warp (its lane 0 never shares a warp with spinning waiters)

template <int DHP, bool DROP>
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse_out,
                int S, int H, int dh, float scale, uint32_t thresh8, float drop_scale, uint64_t seed,
                uint64_t stream_id) {
  constexpr uint32_t RS = TileGeom<DHP>::RS;
  constexpr uint32_t RS_P = (kSP / 8) * 128;  // P tile [128 q, 384 keys]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kSP * DHP * 2;
  uint8_t* sV = sK + kSP * DHP * 2;
  uint8_t* sP = sV + kSP * DHP * 2;
  float* sMax = reinterpret_cast<float*>(sP + 128 * kSP * 2);  // [3][128]
  float* sSum = sMax + 3 * 128;                                 // [3][128]
  uint8_t* sOut = reinterpret_cast<uint8_t*>(sSum + 3 * 128);   // [128][dh] bf16 output staging
  __shared__ uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_slot;

  const int D = H * dh;
  const int ld = 3 * D;
  const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = warp >> 2;                 // key tile owned by this warp
  const int row = (warp & 3) * 32 + lane;    // query row inside the tile == TMEM lane
  const __nv_bfloat16* qbase = qkv + static_cast<size_t>(b) * S * ld + h * dh;
  const int ntq = (S + 127) / 128;           // query tiles == key tiles
  const int nk16 = (S + 15) / 16;            // PV k-steps / dropout groups per row
  const uint32_t thresh4 = thresh8 * 0x01010101u;

  if (warp == 0) WM_TICK(0);
  {
    uint8_t* const tiles[3] = {sQ, sK, sV};
    const __nv_bfloat16* const srcs[3] = {qbase, qbase + D, qbase + 2 * D};
    const int lds[3] = {ld, ld, ld};
    load_head_tiles<DHP, 3, (kSP + (32 / (DHP / 4)) * (kFwdThreads / 32) - 1) / ((32 / (DHP / 4)) * (kFwdThreads / 32))>(tiles, srcs, lds, S, kSP, dh);
  }
  if (warp == 0) WM_TICK(1);
  cp_async_wait_all();
  if (warp == 0) WM_TICK(2);
  if (tid == 0) {
    mbar_init(&bar_s, 1);
    mbar_init(&bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) WM_TICK(3);
  const uint32_t tS = tmem, tO = tmem + kSP;
  const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
  const uint32_t idesc_o = umma_idesc_bf16(128, DHP, 0, 1);
  const float c2 = scale * 1.4426950408889634f;
  const bool issue_warp = warp == 12;
  const uint64_t dQ0 = umma_smem_desc(smem_u32(sQ), 128, RS, UMMA_SWZ_NONE);
  const uint64_t dK0 = umma_smem_desc(smem_u32(sK), 128, RS, UMMA_SWZ_NONE);
  const uint64_t dP0 = umma_smem_desc(smem_u32(sP), 128, RS_P, UMMA_SWZ_NONE);
  // V as MN-major B: mn = head dim (SBO = 128 between 8-column groups), k = key rows (LBO = RS)
  const uint64_t dV0 = umma_smem_desc(smem_u32(sV), RS, 128, UMMA_SWZ_NONE);

  auto issue_scores = [&](int it) {  // S[128, 128*ntq] = Q_it K^T
    for (int jt = 0; jt < ntq; ++jt) {
#pragma unroll
      for (int k = 0; k < DHP / 16; ++k)
        umma_ss(tS + jt * 128, umma_desc_advance(dQ0, (it * 16) * RS + k * 256),
                umma_desc_advance(dK0, (jt * 16) * RS + k * 256), idesc_s, k != 0);
    }
    umma_commit(&bar_s);
  };
  if (issue_warp) {
    // ---- MMA-issue warp: mirrors the three block barriers of every tile, issues after the second one
    if (lane == 0) issue_scores(0);
    for (int it = 0; it < ntq; ++it) {
      __syncthreads();  // row maxima exchanged
      __syncthreads();  // P tile complete in smem, all score reads done
      if (lane == 0) {
        tc_fence_after();
        for (int k = 0; k < nk16; ++k)
          umma_ss(tO, umma_desc_advance(dP0, k * 256), umma_desc_advance(dV0, (k * 2) * RS), idesc_o, k != 0);
        umma_commit(&bar_o);
        if (it + 1 < ntq) issue_scores(it + 1);  // runs under this tile's epilogue
      }
      __syncwarp();
      tc_fence_before();
      __syncthreads();  // epilogue done: O and P reusable
    }
  } else {
  uint32_t ph_s = 0, ph_o = 0;
  // Dropout keep masks of a whole query tile (8 Philox blocks per thread) are generated while the thread would
  // otherwise sit in an mbarrier wait (first tile: under the score MMAs; later tiles: under the previous PV).
  uint32_t kmask[8][4];
  auto gen_masks = [&](int it) {
    if (!DROP || grp >= ntq) return;
    const int qq = it * 128 + row;
    const uint64_t rowbase = (static_cast<uint64_t>(bh) * S + (qq < S ? qq : 0)) * nk16;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int k0 = grp * 128 + c * 16;
      if (k0 < nk16 * 16) {
        keep_masks16(seed, stream_id, rowbase + (k0 >> 4), thresh4, kmask[c]);
        WM_PIN(kmask[c][0]); WM_PIN(kmask[c][1]); WM_PIN(kmask[c][2]); WM_PIN(kmask[c][3]);
      }
    }
  };
  gen_masks(0);
  for (int it = 0; it < ntq; ++it) {
    const int q = it * 128 + row;
    const bool qvalid = q < S;
    if (warp == 0) WM_TICK(4 + it * 8);
    mbar_wait(&bar_s, ph_s, 41);
    ph_s ^= 1u;
    tc_fence_after();
    if (warp == 0) WM_TICK(5 + it * 8);
    // ---- pass 1: max over this warp's key tile
    float mloc = -INFINITY;
    if (grp < ntq) {
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        const int kbase = grp * 128 + c0;
        if (kbase >= S) break;
        uint32_t v[32];
        tmem_ld32(tS + lane_sel + kbase, v);
        tmem_ld_wait();
        if (kbase + 32 <= S) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mloc = fmaxf(mloc, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (kbase + j < S) mloc = fmaxf(mloc, __uint_as_float(v[j]));
        }
      }
    }
    sMax[grp * 128 + row] = mloc;
    if (warp == 0) WM_TICK(6 + it * 8);
    __syncthreads();
    if (warp == 0) WM_TICK(7 + it * 8);
    const float mrow = fmaxf(fmaxf(sMax[row], sMax[128 + row]), sMax[256 + row]);
    const float mneg = -mrow * c2;
    // ---- pass 2: exp2, row sum, dropout, P -> smem (bf16, K-major over keys)
    float lsum = 0.0f;
    if (grp < ntq) {
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 16) {
        const int k0 = grp * 128 + c0;
        if (k0 >= nk16 * 16) break;
        uint32_t v[16];
        tmem_ld16(tS + lane_sel + k0, v);
        tmem_ld_wait();
        const uint32_t(&km)[4] = kmask[c0 >> 4];
        uint32_t pb[16];  // P as fp32 bit patterns (masked by the keep decision)
        if (k0 + 16 <= S) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float e = fast_exp2(fmaf(__uint_as_float(v[j]), c2, mneg));
            lsum += e;
            pb[j] = DROP ? (__float_as_uint(e) & WM_KEEP32(km, j)) : __float_as_uint(e);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float e = 0.0f;
            if (k0 + j < S) e = fast_exp2(fmaf(__uint_as_float(v[j]), c2, mneg));
            lsum += e;
            pb[j] = DROP ? (__float_as_uint(e) & WM_KEEP32(km, j)) : __float_as_uint(e);
          }
        }
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(pb[g8 * 8 + 0]), __uint_as_float(pb[g8 * 8 + 1]));
          pk.y = pack_bf16x2(__uint_as_float(pb[g8 * 8 + 2]), __uint_as_float(pb[g8 * 8 + 3]));
          pk.z = pack_bf16x2(__uint_as_float(pb[g8 * 8 + 4]), __uint_as_float(pb[g8 * 8 + 5]));
          pk.w = pack_bf16x2(__uint_as_float(pb[g8 * 8 + 6]), __uint_as_float(pb[g8 * 8 + 7]));
          const int kc = k0 + g8 * 8;
          *reinterpret_cast<uint4*>(sP + (row >> 3) * RS_P + (kc >> 3) * 128 + (row & 7) * 16) = pk;
        }
      }
    }
    sSum[grp * 128 + row] = lsum;
    if (warp == 0) WM_TICK(8 + it * 8);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) WM_TICK(9 + it * 8);
    if (it + 1 < ntq) gen_masks(it + 1);
    mbar_wait(&bar_o, ph_o, 43);
    ph_o ^= 1u;
    tc_fence_after();
    if (warp == 0) WM_TICK(10 + it * 8);
    // ---- epilogue: warp group g scales head-dim columns [16g, 16g+16) of its row and parks them in a compact
    // [rows, dh] staging tile; the tile then leaves with row-contiguous 8-byte pieces (about 3 rows per warp
    // store instead of 32 different rows: thread-per-row global stores cost ~1 LSU cycle per 32-byte sector)
    const float tot = sSum[row] + sSum[128 + row] + sSum[256 + row];
    if (grp * 16 < DHP) {
      const float inv = drop_scale / tot;
      uint32_t v[16];
      tmem_ld16(tO + lane_sel + grp * 16, v);
      tmem_ld_wait();
      uint8_t* srow = sOut + row * (dh * 2) + grp * 32;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        if (grp * 16 + j < dh) {  // dh % 4 == 0
          uint2 pk;
          pk.x = pack_bf16x2(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
          pk.y = pack_bf16x2(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
          *reinterpret_cast<uint2*>(srow + j * 2) = pk;
        }
      }
      if (grp == 0 && qvalid && lse_out) lse_out[static_cast<size_t>(bh) * S + q] = mrow * scale + logf(tot);
    }
    named_bar_sync(1, kFwdThreads - 32);
    {
      const int pv = dh >> 2;
      const int nrows = min(128, S - it * 128);
      const float inv_pv = 1.0f / static_cast<float>(pv);
      __nv_bfloat16* obase = ctx + (static_cast<size_t>(b) * S + it * 128) * D + h * dh;
      for (int i = tid; i < nrows * pv; i += kFwdThreads - 32) {
        const int r = static_cast<int>((static_cast<float>(i) + 0.5f) * inv_pv);
        const int pp = i - r * pv;
        *reinterpret_cast<uint2*>(obase + static_cast<size_t>(r) * D + pp * 4) = *reinterpret_cast<const uint2*>(sOut + i * 8);
      }
    }
    if (warp == 0) WM_TICK(11 + it * 8);
    tc_fence_before();
    __syncthreads();  // O drained, sMax / sSum / sP reusable
  }
  if (warp == 0) WM_TICK(28);
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------
// backward: 16 warps. Warp w reads TMEM lane quarter w%4 (query rows of tile i) and owns 32 of the 128 key
// columns of tile j (w/4). Per (j, i): S = Q_i K_j^T and dP = dO_i V_j^T land in TMEM, the warps write
// P and dS (bf16) to smem, then THREE threads issue in parallel {dV_j += P^T dO_i, dK_j += dS^T Q_i},
// {dQ_i += dS K_j} and {next pair's S, dP}; each commits to the same 3-arrival mbarrier.
// No validity masks are needed: padded query / key rows are zero in every staged tile, so whatever P and dS
// hold there is multiplied by zero rows or lands in rows that are never stored.
// ------------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 544;  // 16 elementwise warps + 1 MMA-issue warp

template <int DHP>
WM_DEVICE void store_acc_chunk(uint32_t taddr, __nv_bfloat16* dst, int c0, int dh, bool valid) {
  uint32_t v[16];
  tmem_ld16(taddr + c0, v);
  tmem_ld_wait();
  if (valid) {
#pragma unroll
    for (int jj = 0; jj < 16; jj += 4) {
      if (c0 + jj < dh) {
        uint2 pk;
        pk.x = pack_bf16x2(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1]));
        pk.y = pack_bf16x2(__uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
        *reinterpret_cast<uint2*>(dst + c0 + jj) = pk;
      }
    }
  }
}

template <int DHP, bool DROP>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ ctx,
                const __nv_bfloat16* __restrict__ dctx, const float* __restrict__ lse,
                __nv_bfloat16* __restrict__ dqkv, int S, int H, int dh, float scale, uint32_t thresh8,
                float drop_scale, uint64_t seed, uint64_t stream_id) {
  constexpr uint32_t RS = TileGeom<DHP>::RS;
  constexpr uint32_t RS_P = (128 / 8) * 128;  // P / dS tiles are [128 q, 128 keys]
  constexpr int NCH = DHP / 16;               // 16-column chunks per accumulator
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kSP * DHP * 2;
  uint8_t* sV = sK + kSP * DHP * 2;
  uint8_t* sdO = sV + kSP * DHP * 2;
  uint8_t* sP = sdO + kSP * DHP * 2;
  uint8_t* sdS = sP + 128 * 128 * 2;
  float* sLse = reinterpret_cast<float*>(sdS + 128 * 128 * 2);  // [384] -lse * log2(e) (+ log2(drop_scale))
  float* sDelta = sLse + kSP;                                   // [384] delta * scale
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;

  const int D = H * dh;
  const int ld = 3 * D;
  const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = warp >> 2;               // 32-key column slice of the current key tile
  const int row = (warp & 3) * 32 + lane;  // row inside a 128-row tile == TMEM lane
  const __nv_bfloat16* qbase = qkv + static_cast<size_t>(b) * S * ld + h * dh;
  const __nv_bfloat16* obase = ctx + static_cast<size_t>(b) * S * D + h * dh;
  const __nv_bfloat16* dobase = dctx + static_cast<size_t>(b) * S * D + h * dh;
  const int nt = (S + 127) / 128;
  const int grp_per_row = (S + 15) / 16;
  const uint32_t thresh4 = thresh8 * 0x01010101u;

  if (warp == 0) WM_TICK(32);
  {  // two rounds of two tiles keep the in-flight loads within the 544-thread register budget
    constexpr int kIt = (kSP + (32 / (DHP / 4)) * (kBwdThreads / 32) - 1) / ((32 / (DHP / 4)) * (kBwdThreads / 32));
    uint8_t* const t0[2] = {sQ, sK};
    const __nv_bfloat16* const s0[2] = {qbase, qbase + D};
    const int l0[2] = {ld, ld};
    load_head_tiles<DHP, 2, kIt>(t0, s0, l0, S, kSP, dh);
    uint8_t* const t1[2] = {sV, sdO};
    const __nv_bfloat16* const s1[2] = {qbase + 2 * D, dobase};
    const int l1[2] = {ld, D};
    load_head_tiles<DHP, 2, kIt>(t1, s1, l1, S, kSP, dh);
  }
  if (warp == 0) WM_TICK(33);
  if (tid < kSP) {  // per-row statistics: LSE (exp2 domain) and delta = sum_d dO * O (pre-scaled)
    float l = 0.0f, acc = 0.0f;
    if (tid < S) {
      l = -lse[static_cast<size_t>(bh) * S + tid] * 1.4426950408889634f;
      const uint2* po = reinterpret_cast<const uint2*>(obase + static_cast<size_t>(tid) * D);
      const uint2* pd = reinterpret_cast<const uint2*>(dobase + static_cast<size_t>(tid) * D);
      uint2 o[DHP / 4], d[DHP / 4];
#pragma unroll
      for (int p = 0; p < DHP / 4; ++p) {  // all loads in flight at once
        const bool ok = p < dh / 4;
        o[p] = ok ? __ldg(po + p) : make_uint2(0u, 0u);
        d[p] = ok ? __ldg(pd + p) : make_uint2(0u, 0u);
      }
#pragma unroll
      for (int p = 0; p < DHP / 4; ++p) {
        acc = fmaf(bf16_lo(o[p].x), bf16_lo(d[p].x), acc);
        acc = fmaf(bf16_hi(o[p].x), bf16_hi(d[p].x), acc);
        acc = fmaf(bf16_lo(o[p].y), bf16_lo(d[p].y), acc);
        acc = fmaf(bf16_hi(o[p].y), bf16_hi(d[p].y), acc);
      }
    }
    sLse[tid] = l;
    sDelta[tid] = acc * scale;
  }
  if (warp == 0) WM_TICK(51);
  cp_async_wait_all();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) WM_TICK(34);
  const uint32_t tS = tmem, tdP = tmem + 128, tdV = tmem + 256, tdK = tmem + 256 + DHP, tdQ = tmem + 256 + 2 * DHP;
  const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
  const uint32_t idesc_kv = umma_idesc_bf16(128, DHP, 1, 1);
  const uint32_t idesc_q = umma_idesc_bf16(128, DHP, 0, 1);
  const float c2 = scale * 1.4426950408889634f;
  const float ds_scale = drop_scale * scale;  // dS = P * (dP * keep * drop_scale - delta) * scale
  const bool issue_warp = warp == 16;
  // K-major views (contract over head dim): LBO = 128, SBO = RS. MN-major views (contract over rows): LBO = RS, SBO = 128
  const uint64_t kQ = umma_smem_desc(smem_u32(sQ), 128, RS, UMMA_SWZ_NONE), mQ = umma_smem_desc(smem_u32(sQ), RS, 128, UMMA_SWZ_NONE);
  const uint64_t kK = umma_smem_desc(smem_u32(sK), 128, RS, UMMA_SWZ_NONE), mK = umma_smem_desc(smem_u32(sK), RS, 128, UMMA_SWZ_NONE);
  const uint64_t kV = umma_smem_desc(smem_u32(sV), 128, RS, UMMA_SWZ_NONE);
  const uint64_t kdO = umma_smem_desc(smem_u32(sdO), 128, RS, UMMA_SWZ_NONE), mdO = umma_smem_desc(smem_u32(sdO), RS, 128, UMMA_SWZ_NONE);
  // P / dS [128 q, 128 keys]: MN-major (mn = keys, k = q rows) for dV / dK, K-major over keys for dQ
  const uint64_t mP = umma_smem_desc(smem_u32(sP), RS_P, 128, UMMA_SWZ_NONE);
  const uint64_t mdS = umma_smem_desc(smem_u32(sdS), RS_P, 128, UMMA_SWZ_NONE);
  const uint64_t kdS = umma_smem_desc(smem_u32(sdS), 128, RS_P, UMMA_SWZ_NONE);

  auto issue_scores = [&](int i, int j) {  // S = Q_i K_j^T, dP = dO_i V_j^T
#pragma unroll
    for (int k = 0; k < DHP / 16; ++k)
      umma_ss(tS, umma_desc_advance(kQ, (i * 16) * RS + k * 256), umma_desc_advance(kK, (j * 16) * RS + k * 256), idesc_s, k != 0);
#pragma unroll
    for (int k = 0; k < DHP / 16; ++k)
      umma_ss(tdP, umma_desc_advance(kdO, (i * 16) * RS + k * 256), umma_desc_advance(kV, (j * 16) * RS + k * 256), idesc_s, k != 0);
  };
  auto store_kv = [&](int j) {  // thread = key row; the 2*NCH 16-column chunks are dealt round-robin to the 4 groups
    const int kr = j * 128 + row;
    const bool kvalid = kr < S;
    __nv_bfloat16* drow = dqkv + (static_cast<size_t>(b) * S + (kvalid ? kr : 0)) * ld + h * dh;
    for (int c = grp; c < 2 * NCH; c += 4) {
      const int which = c / NCH, cc = c - which * NCH;
      store_acc_chunk<DHP>((which == 0 ? tdK : tdV) + lane_sel, drow + (which == 0 ? D : 2 * D), cc * 16, dh, kvalid);
    }
  };

  if (issue_warp) {
    // ---- MMA-issue warp: one block barrier per (j, i) pair (P / dS complete), then all five products + commit
    if (lane == 0) {
      issue_scores(0, 0);
      umma_commit(&bar);
    }
    for (int j = 0; j < nt; ++j) {
      for (int i = 0; i < nt; ++i) {
        __syncthreads();
        if (lane == 0) {
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 128 / 16; ++k) {  // contraction over the 128 query rows of tile i
            umma_ss(tdV, umma_desc_advance(mP, (k * 2) * RS_P), umma_desc_advance(mdO, (i * 16 + k * 2) * RS), idesc_kv, (i | k) != 0);
            umma_ss(tdK, umma_desc_advance(mdS, (k * 2) * RS_P), umma_desc_advance(mQ, (i * 16 + k * 2) * RS), idesc_kv, (i | k) != 0);
          }
#pragma unroll
          for (int k = 0; k < 128 / 16; ++k)  // dQ_i += dS K_j, contraction over the 128 keys of tile j
            umma_ss(tdQ + i * DHP, umma_desc_advance(kdS, k * 256), umma_desc_advance(mK, (j * 16 + k * 2) * RS), idesc_q, (j | k) != 0);
          const int in = i + 1 < nt ? i + 1 : 0, jn = i + 1 < nt ? j : j + 1;
          if (jn < nt) issue_scores(in, jn);
          umma_commit(&bar);
        }
        __syncwarp();
      }
    }
  } else {
    uint32_t phase = 0;
    for (int j = 0; j < nt; ++j) {
      for (int i = 0; i < nt; ++i) {
        if (warp == 0 && j == 0) WM_TICK(35 + i * 3);
        // this pair's dropout masks (2 Philox blocks per thread) are generated under the MMA wait below
        uint32_t kmA[4], kmB[4];
        if (DROP) {
          const int qq = i * 128 + row;
          const uint64_t rowbase = (static_cast<uint64_t>(bh) * S + (qq < S ? qq : 0)) * grp_per_row;
          const int kk = j * 128 + grp * 32;
          keep_masks16(seed, stream_id, rowbase + (kk >> 4), thresh4, kmA);
          keep_masks16(seed, stream_id, rowbase + (kk >> 4) + 1, thresh4, kmB);
#pragma unroll
          for (int w = 0; w < 4; ++w) { WM_PIN(kmA[w]); WM_PIN(kmB[w]); }
        }
        mbar_wait(&bar, phase, 51);  // S/dP of (j, i) ready; every earlier product has completed as well
        phase ^= 1u;
        tc_fence_after();
        if (warp == 0 && j == 0) WM_TICK(36 + i * 3);
        if (i == 0 && j > 0) store_kv(j - 1);  // dK/dV of the previous key tile are final
        const int q = i * 128 + row;
        const float lneg = sLse[i * 128 + row];
        const float dl = sDelta[i * 128 + row];
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          const int c0 = grp * 32 + hh * 16;   // column inside the key tile
          const int k0 = j * 128 + c0;          // global key index
          uint32_t vs[16], vd[16];
          tmem_ld16(tS + lane_sel + c0, vs);
          tmem_ld16(tdP + lane_sel + c0, vd);
          tmem_ld_wait();
          uint32_t km[4];
          if (DROP) {
#pragma unroll
            for (int w = 0; w < 4; ++w) km[w] = hh ? kmB[w] : kmA[w];
          }
          float pp[16], ds[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            // valid entries have (s - lse) <= 0; the clamp only tames padded keys / rows (inf * 0 would be NaN)
            const float p = fast_exp2(fminf(fmaf(__uint_as_float(vs[jj]), c2, lneg), 0.0f));
            if (DROP) {
              const uint32_t m32 = WM_KEEP32(km, jj);
              pp[jj] = __uint_as_float(__float_as_uint(p * drop_scale) & m32);
              ds[jj] = p * fmaf(__uint_as_float(vd[jj] & m32), ds_scale, -dl);
            } else {
              pp[jj] = p;
              ds[jj] = p * fmaf(__uint_as_float(vd[jj]), scale, -dl);
            }
          }
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            uint4 pk, dk;
            pk.x = pack_bf16x2(pp[g8 * 8 + 0], pp[g8 * 8 + 1]);
            pk.y = pack_bf16x2(pp[g8 * 8 + 2], pp[g8 * 8 + 3]);
            pk.z = pack_bf16x2(pp[g8 * 8 + 4], pp[g8 * 8 + 5]);
            pk.w = pack_bf16x2(pp[g8 * 8 + 6], pp[g8 * 8 + 7]);
            dk.x = pack_bf16x2(ds[g8 * 8 + 0], ds[g8 * 8 + 1]);
            dk.y = pack_bf16x2(ds[g8 * 8 + 2], ds[g8 * 8 + 3]);
            dk.z = pack_bf16x2(ds[g8 * 8 + 4], ds[g8 * 8 + 5]);
            dk.w = pack_bf16x2(ds[g8 * 8 + 6], ds[g8 * 8 + 7]);
            const int kc = c0 + g8 * 8;
            const uint32_t off = (row >> 3) * RS_P + (kc >> 3) * 128 + (row & 7) * 16;
            *reinterpret_cast<uint4*>(sP + off) = pk;
            *reinterpret_cast<uint4*>(sdS + off) = dk;
          }
        }
        if (warp == 0 && j == 0) WM_TICK(37 + i * 3);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
      }
      if (warp == 0) WM_TICK(44 + j);
    }
    mbar_wait(&bar, phase, 52);  // the last gradient products
    tc_fence_after();
    store_kv(nt - 1);
    for (int c = grp; c < nt * NCH; c += 4) {  // dQ: thread = query row
      const int i = c / NCH, cc = c - i * NCH;
      const int q = i * 128 + row;
      const bool qvalid = q < S;
      __nv_bfloat16* dst = dqkv + (static_cast<size_t>(b) * S + (qvalid ? q : 0)) * ld + h * dh;
      store_acc_chunk<DHP>(tdQ + i * DHP + lane_sel, dst, cc * 16, dh, qvalid);
    }
    if (warp == 0) WM_TICK(48);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static int padded_dh(int dh) { return dh <= 16 ? 16 : dh <= 32 ? 32 : dh <= 48 ? 48 : 0; }

template <int DHP>
static int launch_fwd_t(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, float* lse, int B, int S, int H, int dh,
                        float scale, uint32_t thresh8, float dscale, uint64_t seed, uint64_t stream_id,
                        cudaStream_t stream) {
  const int smem = 3 * kSP * DHP * 2 + 128 * kSP * 2 + 6 * 128 * 4 + 128 * DHP * 2 + 256;
  auto kern = thresh8 ? attn_fwd_kernel<DHP, true> : attn_fwd_kernel<DHP, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return WM_ERR_CUDA;
  kern<<<B * H, kFwdThreads, smem, stream>>>(qkv, ctx, lse, S, H, dh, scale, thresh8, dscale, seed, stream_id);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}
template <int DHP>
static int launch_bwd_t(const __nv_bfloat16* qkv, const __nv_bfloat16* ctx, const __nv_bfloat16* dctx,
                        const float* lse, __nv_bfloat16* dqkv, int B, int S, int H, int dh, float scale,
                        uint32_t thresh8, float dscale, uint64_t seed, uint64_t stream_id, cudaStream_t stream) {
  const int smem = 4 * kSP * DHP * 2 + 2 * 128 * 128 * 2 + 2 * kSP * 4 + 256;
  auto kern = thresh8 ? attn_bwd_kernel<DHP, true> : attn_bwd_kernel<DHP, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return WM_ERR_CUDA;
  kern<<<B * H, kBwdThreads, smem, stream>>>(qkv, ctx, dctx, lse, dqkv, S, H, dh, scale, thresh8, dscale, seed,
                                             stream_id);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// drop_thresh is the 16-bit threshold used everywhere else (round(p*65536)); attention rounds it to 8 bits
static void attn_drop_params(uint32_t drop_thresh16, uint32_t* thresh8, float* scale) {
  *thresh8 = (drop_thresh16 + 128u) >> 8;
  *scale = *thresh8 ? 256.0f / static_cast<float>(256u - *thresh8) : 1.0f;
}

int launch_attn_fwd(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, float* lse, int B, int S, int H, int dh,
                    uint32_t drop_thresh, float drop_scale, uint64_t seed, uint64_t stream_id, cudaStream_t stream) {
  (void)drop_scale;
  if (B <= 0 || H <= 0 || S <= 0 || S > kSP || (dh & 3)) return WM_ERR_SHAPE;
  const int dhp = padded_dh(dh);
  if (!dhp) return WM_ERR_SHAPE;
  uint32_t t8;
  float ds;
  attn_drop_params(drop_thresh, &t8, &ds);
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  switch (dhp) {
    case 16: return launch_fwd_t<16>(qkv, ctx, lse, B, S, H, dh, scale, t8, ds, seed, stream_id, stream);
    case 32: return launch_fwd_t<32>(qkv, ctx, lse, B, S, H, dh, scale, t8, ds, seed, stream_id, stream);
    default: return launch_fwd_t<48>(qkv, ctx, lse, B, S, H, dh, scale, t8, ds, seed, stream_id, stream);
  }
}

int launch_attn_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* ctx, const __nv_bfloat16* dctx, const float* lse,
                    __nv_bfloat16* dqkv, int B, int S, int H, int dh, uint32_t drop_thresh, float drop_scale,
                    uint64_t seed, uint64_t stream_id, cudaStream_t stream) {
  (void)drop_scale;
  if (B <= 0 || H <= 0 || S <= 0 || S > kSP || (dh & 3)) return WM_ERR_SHAPE;
  const int dhp = padded_dh(dh);
  if (!dhp) return WM_ERR_SHAPE;
  uint32_t t8;
  float ds;
  attn_drop_params(drop_thresh, &t8, &ds);
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  switch (dhp) {
    case 16: return launch_bwd_t<16>(qkv, ctx, dctx, lse, dqkv, B, S, H, dh, scale, t8, ds, seed, stream_id, stream);
    case 32: return launch_bwd_t<32>(qkv, ctx, dctx, lse, dqkv, B, S, H, dh, scale, t8, ds, seed, stream_id, stream);
    default: return launch_bwd_t<48>(qkv, ctx, dctx, lse, dqkv, B, S, H, dh, scale, t8, ds, seed, stream_id, stream);
  }
}

}  // namespace wm
