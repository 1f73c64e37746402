// wm_attn.cu -- fused multi-head self-attention forward / backward for S <= 384 on tcgen05 (sm_100a).
//
// Replaces F.scaled_dot_product_attention(q, k, v, None, dropout_p, False) as reached from
// nn.TransformerEncoderLayer (torch/nn/functional.py:6666-6696; reference call site
// src/pretraining/models/weatherbert.py:45-54,116-118). Head dims of this model family are 12/20/28/36.
// Both kernels are persistent (one CTA per SM walks over (batch, head) items), warp-specialised (TMA producer,
// MMA issue, elementwise, store) and keep probabilities in tensor memory; see the comments in front of each.
// Operand tiles are UNSWIZZLED canonical core-matrix columns: an 8-column chunk of R rows is R consecutive
// 16-byte rows (elem(r, d) at (d/8)*R*16 + r*16 + (d%8)*2), which one and the same tile can feed to tcgen05.mma
// either K-major (contract over head dim: LBO = R*16, SBO = 128) or MN-major (contract over rows: LBO = 128,
// SBO = R*16).
// Dropout on P: the counter hash of wm_common.cuh keyed by (seed, stream); group (bh*S + q)*ceil(S/16) + k/16 decides
// 16 consecutive keys of a query row, one byte each; the forward kernel also stores the decisions as one 32-bit word
// per (query row, 32-key slice) for the backward kernel.
#include "wm_kernels.h"

namespace wm {

constexpr int kSP = 384;                   // key rows staged per head (S <= 384)
constexpr int kAttnMaskSlices = kSP / 32;  // 32-key slices per query row in the dropout word buffer
// Dropout keep words: one 32-bit word per (query row, 32-key slice), bit k = key 32*slice + k. Stored so that the
// block the backward kernel needs for one half-tile -- key tile j (4 slices) x 64 query rows -- is contiguous:
// [item][key tile j (3)][query block of 64 (6)][slice in tile (4)][query row in block (64)].
__host__ __device__ inline size_t attn_drop_word_index(int item, int j, int qblk, int slice_in_tile, int qrow) {
  return (((static_cast<size_t>(item) * 3 + j) * 6 + qblk) * 4 + slice_in_tile) * 64 + qrow;
}

// Dropout on attention probabilities: the shared counter hash of wm_common.cuh; one counter decides four
// consecutive keys of one query row (two flag words of two keys each), a group of 16 keys is four consecutive counters.
// The forward kernel stores its decisions for the backward kernel as one 32-bit word per (query row, 32-key slice):
// key k of the slice sits at bit attn_keep_bit(k) -- even keys in the low half, odd keys in the high half, the order
// in which the packed bf16x2 pair masks yield them (one LOP3 per pair gathers two bits).
__host__ __device__ constexpr int attn_keep_bit(int k) { return (k >> 1) + 16 * (k & 1); }
// bf16x2 packing: cvt.rn.bf16x2.f32 runs at ~7.5 thread instructions per clock and SM on B200 (tools/mufubench.cu;
// the exp2 of the same two elements costs as much again), the integer form -- round half away from zero by adding
// 0x8000 to the bit patterns, then one PRMT taking the two upper halves -- moves that work to the ALU pipe.
// -DWM_ATTN_INT_PACK=1 selects it for the P / dS tiles of both attention kernels (A/B: profiles/r02_attn_variants.txt).
#ifndef WM_ATTN_INT_PACK
#define WM_ATTN_INT_PACK 0
#endif
#ifndef WM_ATTN_INTERLEAVE
#define WM_ATTN_INTERLEAVE 1  // dropout hash and exponentials of the same elements side by side (-5 % with dropout)
#endif
WM_DEVICE uint32_t attn_pack2(float lo, float hi) {
#if WM_ATTN_INT_PACK
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(d) : "r"(__float_as_uint(lo) + 0x8000u), "r"(__float_as_uint(hi) + 0x8000u));
  return d;
#else
  return pack_bf16x2(lo, hi);
#endif
}

// One warp copies a compact staging tile [nrows, 8 * pv bytes] from shared memory to global rows of pitch ld
// (elements) with row-contiguous 8-byte pieces (head slices are only 8-byte aligned). Piece idx = lane + 32 k lives
// in row idx / pv: row and piece advance incrementally (a single warp runs ~1 dependent instruction per 5-10 cycles,
// the first version's per-piece float divide made a tile cost ~5k cycles), six pieces in flight.
WM_DEVICE void store_staged_tile(const uint8_t* src, __nv_bfloat16* gbase, size_t ld, int nrows, int pv, int lane) {
  const int total = nrows * pv;
  int r = lane / pv, p = lane - r * pv;
  const int dr = 32 / pv, dp = 32 - dr * pv;
  for (int base = 0; base < total; base += 32 * 6) {
    uint2 v[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int idx = base + u * 32 + lane;
      if (idx < total) v[u] = *reinterpret_cast<const uint2*>(src + idx * 8);
    }
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int idx = base + u * 32 + lane;
      if (idx < total) *reinterpret_cast<uint2*>(gbase + static_cast<size_t>(r) * ld + p * 4) = v[u];
      p += dp;
      r += dr;
      if (p >= pv) {
        p -= pv;
        ++r;
      }
    }
  }
}

WM_DEVICE void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
WM_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// (A/B, profiles/r02_attn_variants.txt: taking a quarter or half of the exponentials from the FMA pipe -- Cody-Waite split +
// degree-4 polynomial on packed fp32 -- made the forward kernel 1-5 % SLOWER: the SFU is not its limiter.)
// ------------------------------------------------------------------------------------------------
// forward: persistent, warp-specialised, TMA-fed; probabilities never leave tensor memory.
//   warps 0-11  softmax: warp w reads TMEM lane quarter w%4 (query rows of the current 128-row tile) and owns
//               the 32-column slice w/4 of every 96-key score chunk
//   warp 12     MMA issue (whole warp convergent, one elected lane): O += P_c V_c, then S_{t+1} chunk c
//   warp 13     TMA producer: K/V and Q tiles one head ahead; zeroes the neighbour-head columns of each Q tile
//   warp 14     ctx store: drains the bf16 output staging tile with row-contiguous 8-byte stores
// Q/K/V head slices are NOT 16-byte aligned in the token-major QKV activation (head pitch dh*2 = 24..72 B), so
// the producer loads the enclosing 8-column-aligned window [c0, c0 + 8*NCH) with one TMA box per 8-column chunk
// ({8 cols, 128 rows} -> 128 consecutive 16-byte rows = a column of core matrices). The window carries 4 columns
// of the neighbouring head on one side; they are zeroed in the Q tile (one 8-byte store per row) so they drop
// out of Q K^T, and they only produce unused output columns in P V. When 8*NCH is not a multiple of 16 the last
// k-step takes its second core-matrix column from a shared all-zero chunk (per-descriptor leading byte offset).
// Softmax is exact two-pass over scores that sit in TMEM (384 fp32 columns per tile). In pass 2 a warp turns its
// 32 score columns of chunk c into 16 columns of packed bf16 P written back IN PLACE (tcgen05.st); the issue warp
// feeds them to the tensor core as the TMEM A operand of O_t += P_c V_c and then queues S_{t+1} chunk c into the
// same columns (tcgen05.mma executes in issue order), so the next tile's scores are complete when the current
// tile's softmax ends. No shared-memory round trip, no proxy fence and no extra barrier for P.
// (v5 staged P through shared memory: a fence.proxy.async per warp and chunk cost ~300 cycles each and a single
// thread could not issue ~35 small MMAs per tile fast enough -- profiles/r01_attn_phase_ticks.txt.)
// ------------------------------------------------------------------------------------------------
// WM_ATTN_FWD_SLICES = 32-column slices per score chunk = softmax warps per TMEM lane quarter: 3 (12 softmax warps, 96-key
// chunks, TMEM loads double-buffered) or 4 (16 softmax warps, 128-key chunks, single-buffered loads to stay under 104
// registers): both kernels are bound by dependent-issue latency, so a fourth warp per scheduler is worth more than the
// prefetch (profiles/r02_attn_ncu_summary.txt; A/B in profiles/r02_attn_variants.txt).
#ifndef WM_ATTN_FWD_SLICES
#define WM_ATTN_FWD_SLICES 4
#endif
constexpr int kFwdSlices = WM_ATTN_FWD_SLICES;
constexpr int kFwdSoftmaxWarps = 4 * kFwdSlices;
constexpr int kFwdWarpMma = kFwdSoftmaxWarps, kFwdWarpProd = kFwdSoftmaxWarps + 1, kFwdWarpStore = kFwdSoftmaxWarps + 2;
constexpr int kFwdThreads = 32 * (kFwdSoftmaxWarps + 3);
constexpr int kKC = 32 * kFwdSlices;       // keys per score chunk
constexpr int kFwdMaxChunks = 384 / kKC;   // chunks that cover kSP key columns
constexpr bool kFwdPrefetch = kFwdSlices == 3;

struct AttnFwdBars {
  uint64_t q_full[3], q_ready[3], q_free[3];
  uint64_t k_full[2], k_free[2], v_full[2], v_free[2];
  uint64_t s_full, p_full[4];
  uint64_t o_full[2], o_free[2];
  uint64_t out_full[2], out_free[2];
};

template <int NCH>
struct AttnFwdGeom {
  static constexpr int DHP = (NCH * 8 + 15) / 16 * 16;
  static constexpr int KVB = 2;                        // K / V buffers
  static constexpr uint32_t CSQ = 128 * 16;            // chunk stride inside a 128-row tile
  static constexpr uint32_t CSK = kSP * 16;            // chunk stride inside a 384-row tile
  static constexpr uint32_t QT = NCH * CSQ, KT = NCH * CSK;
  static constexpr uint32_t OUTB = NCH * 8 * 2 * 128;  // output staging tile (>= 128 * dh * 2)
  // + one chunk of slack: an MN-major B descriptor with N = DHP > 8 * NCH reads one chunk past the last V buffer
  static constexpr uint32_t kSmem = 3 * QT + 2 * KVB * KT + CSK + 2 * OUTB + 2048 + 2 * kFwdSlices * 128 * 4 + 128;
};

template <int NCH, bool DROP>
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ ctx,
                float* __restrict__ lse_out, uint32_t* __restrict__ drop_words, int nitems, int S, int H, int dh,
                float scale, uint32_t thresh15, float drop_scale, DropKeys dkeys_in) {
  using G = AttnFwdGeom<NCH>;
  const DropKeys dkeys = DROP ? drop_keys_live(dkeys_in) : dkeys_in;  // (+ the per-replay words of a captured step)
  constexpr int DHP = G::DHP, KVB = G::KVB, KSTEPS = DHP / 16;
  constexpr bool kZeroTail = (NCH & 1) != 0;  // last k-step: second core-matrix column comes from the zero chunk
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align_up(smem_raw, 128);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 3 * G::QT;
  uint8_t* sV = sK + KVB * G::KT;
  uint8_t* sOut = sV + KVB * G::KT + G::CSK;
  uint8_t* sZero = sOut + 2 * G::OUTB;
  float* sMax = reinterpret_cast<float*>(sZero + 2048);  // [3][128]
  float* sSum = sMax + kFwdSlices * 128;                 // [kFwdSlices][128]
  __shared__ AttnFwdBars bars;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int D = H * dh;
  const int ntq = (S + 127) / 128;             // query tiles per head
  const int nkc = (S + kKC - 1) / kKC;         // score chunks per tile
  const int nrb = (nkc * kKC + 127) / 128;     // 128-row K/V blocks that the chunks touch
  const int nk16 = (S + 15) / 16;              // dropout groups per query row
  const int nmine = nitems > static_cast<int>(blockIdx.x) ? (nitems - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bars.q_full[i], 1);
      mbar_init(&bars.q_ready[i], 1);
      mbar_init(&bars.q_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.k_full[i], 1);
      mbar_init(&bars.k_free[i], 1);
      mbar_init(&bars.v_full[i], 1);
      mbar_init(&bars.v_free[i], 1);
      mbar_init(&bars.o_full[i], 1);
      mbar_init(&bars.o_free[i], kFwdSoftmaxWarps);
      mbar_init(&bars.out_full[i], kFwdSoftmaxWarps);
      mbar_init(&bars.out_free[i], 1);
    }
    mbar_init(&bars.s_full, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bars.p_full[i], kFwdSoftmaxWarps);
    fence_barrier_init();
  }
  for (int i = tid; i < 2048 / 16; i += kFwdThreads) reinterpret_cast<uint4*>(sZero)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == kFwdWarpMma) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem, tO = tmem + kSP;  // O buffers at +0 / +64

  if (warp == kFwdWarpProd) {
    // ------------------------------------------------------------------ TMA producer (+ Q fix-up)
    if (lane == 0) tma_prefetch_desc(&tm_qkv);
    const bool need_fix = (dh & 7) != 0;
    for (int n = 0; n < nmine; ++n) {
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / H, h = item - b * H;
      const int col0 = (h * dh) & ~7;  // 8-column aligned window start inside the Q block
      const int kb = n % KVB;
      if (lane == 0) {
        if (n >= KVB) mbar_wait(&bars.k_free[kb], ((n / KVB) - 1) & 1, 60);
        mbar_arrive_expect_tx(&bars.k_full[kb], NCH * nrb * 2048);
        for (int ch = 0; ch < NCH; ++ch)
          for (int rb = 0; rb < nrb; ++rb)
            tma_load_3d(sK + kb * G::KT + ch * G::CSK + rb * 2048, &tm_qkv, &bars.k_full[kb], D + col0 + ch * 8, rb * 128, b);
        if (n >= KVB) mbar_wait(&bars.v_free[kb], ((n / KVB) - 1) & 1, 61);
        mbar_arrive_expect_tx(&bars.v_full[kb], NCH * nrb * 2048);
        for (int ch = 0; ch < NCH; ++ch)
          for (int rb = 0; rb < nrb; ++rb)
            tma_load_3d(sV + kb * G::KT + ch * G::CSK + rb * 2048, &tm_qkv, &bars.v_full[kb], 2 * D + col0 + ch * 8, rb * 128, b);
        for (int i = 0; i < ntq; ++i) {
          if (n >= 1) mbar_wait(&bars.q_free[i], (n - 1) & 1, 62);
          mbar_arrive_expect_tx(&bars.q_full[i], NCH * 2048);
          for (int ch = 0; ch < NCH; ++ch)
            tma_load_3d(sQ + i * G::QT + ch * G::CSQ, &tm_qkv, &bars.q_full[i], col0 + ch * 8, i * 128, b);
        }
      }
      __syncwarp();
      // zero the neighbouring head's 4 columns of every Q tile (window columns [0, 4) or [dh, dh + 4)), then
      // publish the tile to the issue warp
      const int z0 = ((h * dh) & 7) ? 0 : dh;
      for (int i = 0; i < ntq; ++i) {
        mbar_wait(&bars.q_full[i], n & 1, 71);
        if (need_fix) {
          uint8_t* zp = sQ + i * G::QT + (z0 >> 3) * G::CSQ + (z0 & 7) * 2;
#pragma unroll
          for (int r = 0; r < 4; ++r) *reinterpret_cast<uint2*>(zp + (r * 32 + lane) * 16) = make_uint2(0u, 0u);
          fence_proxy_async_smem();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.q_ready[i]);
      }
    }
  } else if (warp == kFwdWarpMma) {
    // ------------------------------------------------------------------ MMA issue
    // (all 32 lanes run this code convergently; umma_*_warp elect the issuing lane)
    const uint32_t idesc_o = umma_idesc_bf16(128, DHP, 0, 1);
    const uint32_t idesc_s = umma_idesc_bf16(128, kKC, 0, 0);
    const uint32_t zero_addr = smem_u32(sZero);
    auto issue_scores = [&](int i, int kb, int c) {  // S[:, 96c : 96c+96] = Q_i K[96c : 96c+96]^T
      const uint32_t qa = smem_u32(sQ + i * G::QT), ka = smem_u32(sK + kb * G::KT) + c * kKC * 16;
      const uint64_t qd = umma_smem_desc(qa, G::CSQ, 128, UMMA_SWZ_NONE), kd = umma_smem_desc(ka, G::CSK, 128, UMMA_SWZ_NONE);
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        if (kZeroTail && ks == KSTEPS - 1) {
          const uint32_t a0 = qa + ks * 2 * G::CSQ, b0 = ka + ks * 2 * G::CSK;
          umma_ss_warp(tS + c * kKC, umma_smem_desc(a0, zero_addr - a0, 128, UMMA_SWZ_NONE),
                       umma_smem_desc(b0, zero_addr - b0, 128, UMMA_SWZ_NONE), idesc_s, ks != 0);
        } else {
          umma_ss_warp(tS + c * kKC, qd + ks * ((2 * G::CSQ) >> 4), kd + ks * ((2 * G::CSK) >> 4), idesc_s, ks != 0);
        }
      }
    };
    if (nmine > 0) {
      // prologue: all score chunks of the first tile
      mbar_wait(&bars.k_full[0], 0, 63);
      mbar_wait(&bars.q_ready[0], 0, 64);
      tc_fence_after();
      for (int c = 0; c < nkc; ++c) issue_scores(0, 0, c);
      umma_commit_warp(&bars.s_full);
      umma_commit_warp(&bars.q_free[0]);
      if (ntq == 1) umma_commit_warp(&bars.k_free[0]);
    }
    uint32_t t = 0;
    for (int n = 0; n < nmine; ++n) {
      const int kb = n % KVB;
      // V as MN-major B: k = key rows (LBO = 128 between 8-row groups), mn = head dim (SBO = chunk stride)
      const uint64_t vbase = umma_smem_desc(smem_u32(sV + kb * G::KT), 128, G::CSK, UMMA_SWZ_NONE);
      for (int i = 0; i < ntq; ++i, ++t) {
        const uint32_t ob = t & 1u;
        if (t >= 2) mbar_wait(&bars.o_free[ob], ((t >> 1) - 1) & 1, 65);
        if (i == 0) mbar_wait(&bars.v_full[kb], (n / KVB) & 1, 66);
        const bool has_next = (i + 1 < ntq) || (n + 1 < nmine);
        const int i1 = i + 1 < ntq ? i + 1 : 0, n1 = i + 1 < ntq ? n : n + 1;
        const int kb1 = n1 % KVB;
        if (has_next) {
          if (i1 == 0) mbar_wait(&bars.k_full[kb1], (n1 / KVB) & 1, 68);
          mbar_wait(&bars.q_ready[i1], n1 & 1, 69);
        }
        const uint32_t td = tO + ob * 64;
        for (int c = 0; c < nkc; ++c) {
          mbar_wait(&bars.p_full[c], t & 1, 67);  // P chunk c sits in TMEM, every warp is done with S chunk c
          tc_fence_after();
          if (n == 1) WM_TICK(36 + (i * 4 + c) * 2);
          const uint64_t vd = vbase + ((c * kKC * 16) >> 4);
          const int ksteps = min(kKC / 16, (S - c * kKC + 15) / 16);
#pragma unroll
          for (int ks = 0; ks < kKC / 16; ++ks)  // 16 keys = 8 packed columns; slice ks/2 keeps its P in its own first 16 columns
            if (ks < ksteps)
              umma_ts_warp(td, tS + c * kKC + (ks >> 1) * 32 + (ks & 1) * 8, vd + ks * (256 >> 4), idesc_o, (c | ks) != 0);
          if (c == nkc - 1) {
            umma_commit_warp(&bars.o_full[ob]);
            if (i == ntq - 1) umma_commit_warp(&bars.v_free[kb]);
          }
          if (has_next) issue_scores(i1, kb1, c);  // executes after the P V products above (issue order)
          if (n == 1) WM_TICK(37 + (i * 4 + c) * 2);
        }
        if (has_next) {
          umma_commit_warp(&bars.s_full);
          umma_commit_warp(&bars.q_free[i1]);
          if (i1 == ntq - 1) umma_commit_warp(&bars.k_free[kb1]);
        }
      }
    }
  } else if (warp == kFwdWarpStore) {
    // ------------------------------------------------------------------ ctx store
    const int pv = dh >> 2;  // 8-byte pieces per row
    uint32_t t = 0;
    for (int n = 0; n < nmine; ++n) {
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / H, h = item - b * H;
      for (int i = 0; i < ntq; ++i, ++t) {
        const uint32_t ob = t & 1u;
        mbar_wait(&bars.out_full[ob], (t >> 1) & 1, 70);
        const int nrows = min(128, S - i * 128);
        const uint8_t* src = sOut + ob * G::OUTB;
        __nv_bfloat16* obase = ctx + (static_cast<size_t>(b) * S + i * 128) * D + h * dh;
        store_staged_tile(src, obase, D, nrows, pv, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.out_free[ob]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int lq = warp & 3, sl = warp >> 2;   // TMEM lane quarter, 32-column slice of each chunk
    const int row = lq * 32 + lane;            // query row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(lq * 32) << 16;
    const uint32_t add2 = (0x8000u - thresh15) * 0x00010001u;  // per-half addend of the 15-bit keep test (drop_add2)
    const float c2 = scale * 1.4426950408889634f;
    uint32_t t = 0;
    for (int n = 0; n < nmine; ++n) {
      const int item = blockIdx.x + n * gridDim.x;
      const int h = item % H;
      const int front = (h * dh) & 7;
      for (int i = 0; i < ntq; ++i, ++t) {
#define WM_FTICK(k) do { if (warp == 0 && n == 1) WM_TICK(i * 12 + (k)); } while (0)
        WM_FTICK(0);
        const int q = i * 128 + row;
        // dropout group (16 keys) index of this row's key 0: identical numbering in forward and backward
        const uint32_t rowbase = (static_cast<uint32_t>(item) * S + (q < S ? q : 0)) * nk16;  // host-checked to fit
        mbar_wait(&bars.s_full, t & 1, 72);
        tc_fence_after();
        WM_FTICK(2);
        // ---- pass 1: row max over this warp's slices (next chunk's TMEM load in flight under the reduction);
        // three-input maxima: 16 instead of 32 issue slots per 32 columns
        float mloc = -INFINITY;
        {
          uint32_t va[32], vb[kFwdPrefetch ? 32 : 1];
          tmem_ld32(tS + lane_sel + sl * 32, va);
#pragma unroll
          for (int c = 0; c < kFwdMaxChunks; ++c) {
            if (c < nkc) {
              uint32_t(&cur)[32] = (kFwdPrefetch && (c & 1)) ? reinterpret_cast<uint32_t(&)[32]>(vb) : va;
              tmem_ld_wait();
              if (kFwdPrefetch) {
                uint32_t(&nxt)[32] = (c & 1) ? va : reinterpret_cast<uint32_t(&)[32]>(vb);
                if (c + 1 < nkc) tmem_ld32(tS + lane_sel + (c + 1) * kKC + sl * 32, nxt);
              }
              const int k0 = c * kKC + sl * 32;
              if (k0 + 32 <= S) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) mloc = fmax3(mloc, __uint_as_float(cur[j]), __uint_as_float(cur[j + 1]));
              } else {  // the slice that holds key S - 1: only its valid columns (warp-uniform bounds)
                const int nv = S - k0;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  if (j + 2 <= nv) mloc = fmax3(mloc, __uint_as_float(cur[j]), __uint_as_float(cur[j + 1]));
                  else if (j < nv) mloc = fmaxf(mloc, __uint_as_float(cur[j]));
                }
              }
              if (!kFwdPrefetch && c + 1 < nkc) tmem_ld32(tS + lane_sel + (c + 1) * kKC + sl * 32, va);
            }
          }
        }
        sMax[sl * 128 + row] = mloc;
        WM_FTICK(3);
        named_bar_sync(1 + lq, 32 * kFwdSlices);
        WM_FTICK(4);
        float mrow = sMax[row];
#pragma unroll
        for (int k = 1; k < kFwdSlices; ++k) mrow = fmaxf(mrow, sMax[k * 128 + row]);
        const float mneg = -mrow * c2;
        const uint64_t c2p = f2_pack(c2, c2), mnegp = f2_pack(mneg, mneg);
        // ---- pass 2: exp2, row sum, dropout; P (bf16 pairs) replaces the first 16 columns of the slice in TMEM.
        // Packed fp32 pairs: one FFMA2 scales and shifts two scores, one FADD2 adds two probabilities to the row sum.
        uint64_t lsum2 = f2_pack(0.0f, 0.0f);
        {
          uint32_t va[32], vb[kFwdPrefetch ? 32 : 1];
          tmem_ld32(tS + lane_sel + sl * 32, va);
#pragma unroll
          for (int c = 0; c < kFwdMaxChunks; ++c) {
            if (c < nkc) {
              uint32_t(&cur)[32] = (kFwdPrefetch && (c & 1)) ? reinterpret_cast<uint32_t(&)[32]>(vb) : va;
              const int k0 = c * kKC + sl * 32;
              tmem_ld_wait();
              if (kFwdPrefetch) {
                uint32_t(&nxt)[32] = (c & 1) ? va : reinterpret_cast<uint32_t(&)[32]>(vb);
                if (c + 1 < nkc) tmem_ld32(tS + lane_sel + (c + 1) * kKC + sl * 32, nxt);
              }
              uint32_t pk[16];
#if WM_ATTN_INTERLEAVE
              // Dropout and softmax of the SAME four elements side by side: the keep-bit hash is integer work, the
              // exponentials sit on the SFU queue (8 cycles per warp instruction and scheduler). Written as two
              // separate loops the three warps of a scheduler -- which run in step, chunk by chunk -- all hashed and
              // then all queued on the SFU; interleaved, each pipe's work hides under the other's.
              // (The slice that holds key S - 1 takes the same loop with warp-uniform bounds: its invalid columns are
              // skipped, not computed under per-element predicates -- the four warps that own it are otherwise the
              // stragglers every other warp of the tile waits for at the next barrier.)
              if (DROP && k0 + 32 <= S) {  // whole slice: no bounds at all in the loop
                const uint32_t x0 = (rowbase + static_cast<uint32_t>(k0 >> 4)) * 4u;
                uint32_t bits = 0u;
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                  const DropWords f = drop_flags4(x0 + w, dkeys, add2);
                  float xa, xb, xc, xd;
                  f2_unpack(f2_fma(f2_pack_u(cur[4 * w], cur[4 * w + 1]), c2p, mnegp), xa, xb);
                  f2_unpack(f2_fma(f2_pack_u(cur[4 * w + 2], cur[4 * w + 3]), c2p, mnegp), xc, xd);
                  const float ea = fast_exp2(xa), eb = fast_exp2(xb), ec = fast_exp2(xc), ed = fast_exp2(xd);
                  lsum2 = f2_add(lsum2, f2_add(f2_pack(ea, eb), f2_pack(ec, ed)));
                  const uint32_t ma = drop_pair_mask(f.a), mb = drop_pair_mask(f.b);
                  pk[2 * w] = attn_pack2(ea, eb) & ma;
                  pk[2 * w + 1] = attn_pack2(ec, ed) & mb;
                  bits |= (ma & (0x00010001u << (2 * w))) | (mb & (0x00010001u << (2 * w + 1)));
                }
                if (drop_words) drop_words[attn_drop_word_index(item, k0 >> 7, q >> 6, (k0 >> 5) & 3, q & 63)] = bits;
              } else if (DROP) {  // the slice that holds key S - 1 (or lies behind it): warp-uniform bounds, invalid columns skipped
                const uint32_t x0 = (rowbase + static_cast<uint32_t>(k0 >> 4)) * 4u;
                const int nv = S - k0;
                uint32_t bits = 0u;
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                  if (4 * w >= nv) {  // warp-uniform
                    pk[2 * w] = 0u;
                    pk[2 * w + 1] = 0u;
                    continue;
                  }
                  const DropWords f = drop_flags4(x0 + w, dkeys, add2);
                  float xa, xb, xc, xd;
                  f2_unpack(f2_fma(f2_pack_u(cur[4 * w], cur[4 * w + 1]), c2p, mnegp), xa, xb);
                  f2_unpack(f2_fma(f2_pack_u(cur[4 * w + 2], cur[4 * w + 3]), c2p, mnegp), xc, xd);
                  float ea = fast_exp2(xa), eb = fast_exp2(xb), ec = fast_exp2(xc), ed = fast_exp2(xd);
                  if (4 * w + 4 > nv) {  // the one word that straddles S
                    if (4 * w + 1 >= nv) eb = 0.0f;
                    if (4 * w + 2 >= nv) ec = 0.0f;
                    if (4 * w + 3 >= nv) ed = 0.0f;
                  }
                  lsum2 = f2_add(lsum2, f2_add(f2_pack(ea, eb), f2_pack(ec, ed)));
                  const uint32_t ma = drop_pair_mask(f.a), mb = drop_pair_mask(f.b);
                  pk[2 * w] = attn_pack2(ea, eb) & ma;
                  pk[2 * w + 1] = attn_pack2(ec, ed) & mb;
                  bits |= (ma & (0x00010001u << (2 * w))) | (mb & (0x00010001u << (2 * w + 1)));
                }
                if (drop_words) drop_words[attn_drop_word_index(item, k0 >> 7, q >> 6, (k0 >> 5) & 3, q & 63)] = bits;
              } else
#endif
              {
              if (k0 + 32 <= S) {
#pragma unroll
                for (int w = 0; w < 16; ++w) {
                  float x0f, x1f;
                  f2_unpack(f2_fma(f2_pack_u(cur[2 * w], cur[2 * w + 1]), c2p, mnegp), x0f, x1f);
                  const float e0 = fast_exp2(x0f), e1 = fast_exp2(x1f);
                  lsum2 = f2_add(lsum2, f2_pack(e0, e1));
                  pk[w] = attn_pack2(e0, e1);
                }
              } else {
                const int nv = S - k0;  // warp-uniform bounds: invalid columns are skipped, not predicated
#pragma unroll
                for (int w = 0; w < 16; ++w) {
                  if (2 * w >= nv) {
                    pk[w] = 0u;
                    continue;
                  }
                  float x0f, x1f;
                  f2_unpack(f2_fma(f2_pack_u(cur[2 * w], cur[2 * w + 1]), c2p, mnegp), x0f, x1f);
                  const float e0 = fast_exp2(x0f), e1 = (2 * w + 1 < nv) ? fast_exp2(x1f) : 0.0f;
                  lsum2 = f2_add(lsum2, f2_pack(e0, e1));
                  pk[w] = attn_pack2(e0, e1);
                }
              }
              if (DROP) {  // keep flags of the 32 keys: 8 counters, two flag words (= two bf16x2 pairs) each
                const uint32_t x0 = (rowbase + static_cast<uint32_t>(k0 >> 4)) * 4u;
                uint32_t bits = 0u;
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                  const DropWords f = drop_flags4(x0 + w, dkeys, add2);
                  const uint32_t ma = drop_pair_mask(f.a), mb = drop_pair_mask(f.b);
                  pk[2 * w] &= ma;
                  pk[2 * w + 1] &= mb;
                  // key 2i -> bit i, key 2i + 1 -> bit 16 + i (attn_keep_bit)
                  bits |= (ma & (0x00010001u << (2 * w))) | (mb & (0x00010001u << (2 * w + 1)));
                }
                // the backward kernel reads these instead of re-deriving the hash in its transposed order
                if (drop_words) drop_words[attn_drop_word_index(item, k0 >> 7, q >> 6, (k0 >> 5) & 3, q & 63)] = bits;
              }
              }
              if (!kFwdPrefetch && c + 1 < nkc) tmem_ld32(tS + lane_sel + (c + 1) * kKC + sl * 32, va);
              tmem_st16(tS + lane_sel + k0, pk);
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&bars.p_full[c]);
              WM_FTICK(5 + c);
            }
          }
        }
        float lsum;
        {
          float l0, l1;
          f2_unpack(lsum2, l0, l1);
          lsum = l0 + l1;
        }
        sSum[sl * 128 + row] = lsum;
        named_bar_sync(1 + lq, 32 * kFwdSlices);
        WM_FTICK(9);
        float tot = sSum[row];
#pragma unroll
        for (int k = 1; k < kFwdSlices; ++k) tot += sSum[k * 128 + row];
        // ---- epilogue: O / (row sum) -> bf16 staging tile (compact [128, dh]); slice sl takes columns [16 sl, 16 sl + 16)
        const uint32_t ob = t & 1u;
        mbar_wait(&bars.o_full[ob], (t >> 1) & 1, 74);
        tc_fence_after();
        WM_FTICK(10);
        if (t >= 2) mbar_wait(&bars.out_free[ob], ((t >> 1) - 1) & 1, 75);
        if (sl * 16 < DHP) {
          const float inv = drop_scale / tot;
          uint32_t v[16];
          tmem_ld16(tO + ob * 64 + lane_sel + sl * 16, v);
          tmem_ld_wait();
          uint8_t* srow = sOut + ob * G::OUTB + row * (dh * 2);
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const int d = sl * 16 + j - front;  // head-dim index of TMEM column sl*16 + j
            if (d >= 0 && d < dh) {
              uint2 o2;
              o2.x = pack_bf16x2(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
              o2.y = pack_bf16x2(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
              *reinterpret_cast<uint2*>(srow + d * 2) = o2;
            }
          }
        }
        if (sl == 0 && q < S && lse_out) lse_out[static_cast<size_t>(item) * S + q] = mrow * scale + logf(tot);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars.o_free[ob]);
          mbar_arrive(&bars.out_full[ob]);
        }
        WM_FTICK(11);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kFwdWarpMma) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------
// backward: persistent, warp-specialised, TMA-fed; TRANSPOSED score tiles so that P^T and dS^T can feed the
// tensor core straight from tensor memory.
//   warps 0-15  elementwise, two groups of 8 that ping-pong over the stream of half-tiles: group g&1 owns TMEM
//               region g&1. A half-tile is (key tile j: 128 keys = TMEM lanes) x (64 query rows = columns);
//               warp w of a group reads lane quarter w%4 (32 keys) and the 32-query column slice (w/4)%2.
//   warp 16     MMA issue: gradient products (whole warp convergent, one elected lane)
//   warp 20     MMA issue: score tiles of the next-but-one half-tile (a second issue warp: every wait / commit costs
//               a few hundred cycles, one warp doing both was the critical path)
//   warp 17     TMA producer: K_j/V_j tiles (double buffered, fixed up), Q/dO half-tiles + dropout words (ring)
//   warps 18-19 dQ/dK/dV store: drain five bf16 staging tiles with row-contiguous 8-byte stores
// Per half-tile (j, ih):   S^T = K_j Q_ih^T and dP^T = V_j dO_ih^T land in region r (2 x 64 fp32 columns);
//   the group turns them into P^T and dS^T (packed bf16, written back IN PLACE with tcgen05.st) and also parks dS^T
//   in shared memory; then dV_j += P^T dO_ih and dK_j += dS^T Q_ih take their A operand from TMEM, and once both
//   halves of query tile i are there dQ_i += dS_i K_j reads dS from shared memory (MN-major A). The next-but-one
//   half-tile's S^T / dP^T are queued into the same region right behind (tcgen05.mma executes in issue order).
// All five accumulators (dV_j, dK_j, dQ_0..2) stay in TMEM: 2*128 + 5*48 = 496 columns.
// Padded query / key rows are zero in every staged tile (TMA zero fill), so no validity masks are needed: whatever
// P and dS hold there is multiplied by zero rows or lands in rows that are never stored; P is clamped to <= 1
// so that nothing becomes inf * 0.
// Dropout: the forward kernel stores its keep decisions as one 32-bit word per (query row, 32-key slice); a thread
// here owns one key (bit = lane) and walks over query rows, so it shifts its bit into the sign position
// (regenerating Philox in this transposed order would cost one Philox block per element).
// ------------------------------------------------------------------------------------------------
constexpr int kBwdEwWarps = 16;
constexpr int kBwdThreads = 32 * 21;

struct AttnBwdBars {
  uint64_t kv_full[2], kv_ready[2], kv_free[2];
  uint64_t qd_full[6], qd_free[6];
  uint64_t st_full[3];
  uint64_t sdp_full[2], pds_full[2], ts_done[2];
  uint64_t acc_full, acc_free;
  uint64_t out_full[5], out_free[5];
};

template <int NCH>
struct AttnBwdGeom {
  static constexpr int DHP = (NCH * 8 + 15) / 16 * 16;
  static constexpr int RQ = NCH <= 5 ? 5 : 3;          // Q/dO half-tile ring depth (loads run RQ - 2 half-tiles ahead)
  static constexpr uint32_t CS128 = 128 * 16, CS64 = 64 * 16;
  static constexpr uint32_t T128 = NCH * CS128, T64 = NCH * CS64;
  static constexpr uint32_t SLOT = 2 * T64 + 1024;     // Q half-tile, dO half-tile, 4 x 64 dropout words
  static constexpr uint32_t DSB = 128 * 128 * 2;       // dS tile [16 q chunks][128 keys][16 B]
  static constexpr uint32_t OUTB = NCH * 8 * 2 * 128;  // staging tile (>= 128 * dh * 2)
  static constexpr uint32_t STB = kSP * 8;             // per-row statistics of one head (two planes of kSP floats)
  static constexpr uint32_t kSmem = 4 * T128 + CS128 + RQ * SLOT + 2 * DSB + 5 * OUTB + 3 * STB + 2048 + 128;
};

// per (batch, head): two planes of kSP floats, x[q] = -lse * log2(e) + log2(drop_scale) and
// sum over 4 * kN consecutive bf16 elements of o * d, same summation order as the run-time loop in the kernel below
template <int kN>
WM_DEVICE float attn_row_dot(const uint2* __restrict__ po, const uint2* __restrict__ pd) {
  uint2 o[kN], d[kN];
#pragma unroll
  for (int p = 0; p < kN; ++p) {
    o[p] = __ldg(po + p);
    d[p] = __ldg(pd + p);
  }
  float acc = 0.0f;
#pragma unroll
  for (int p = 0; p < kN; ++p) {
    acc = fmaf(bf16_lo(o[p].x), bf16_lo(d[p].x), acc);
    acc = fmaf(bf16_hi(o[p].x), bf16_hi(d[p].x), acc);
    acc = fmaf(bf16_lo(o[p].y), bf16_lo(d[p].y), acc);
    acc = fmaf(bf16_hi(o[p].y), bf16_hi(d[p].y), acc);
  }
  return acc;
}
// y[q] = -(sum_d dO * O) * scale / drop_scale; rows >= S are zero. Planes (not interleaved pairs) so that one 16-byte
// shared-memory load in the backward kernel yields four consecutive x (or y): aligned register pairs for the packed
// fp32 instructions. One thread per (b, q, h); h runs fastest so a warp reads whole token rows.
__global__ void attn_bwd_stats_kernel(const __nv_bfloat16* __restrict__ ctx, const __nv_bfloat16* __restrict__ dctx,
                                      const float* __restrict__ lse, float* __restrict__ stats, int B, int S, int H,
                                      int dh, float scale, float drop_scale) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * kSP * H;
  if (idx >= total) return;
  const int h = static_cast<int>(idx % H);
  const int q = static_cast<int>((idx / H) % kSP);
  const int b = static_cast<int>(idx / (static_cast<long long>(H) * kSP));
  float ox = 0.0f, oy = 0.0f;
  if (q < S) {
    const size_t off = (static_cast<size_t>(b) * S + q) * (static_cast<size_t>(H) * dh) + static_cast<size_t>(h) * dh;
    const uint2* po = reinterpret_cast<const uint2*>(ctx + off);
    const uint2* pd = reinterpret_cast<const uint2*>(dctx + off);
    float acc;
    switch (dh) {  // the head widths of the model table get all their loads in flight at once (a run-time trip count
                   // serialises load -> fma: the kernel ran at 3.7 TB/s)
      case 12: acc = attn_row_dot<3>(po, pd); break;
      case 20: acc = attn_row_dot<5>(po, pd); break;
      case 28: acc = attn_row_dot<7>(po, pd); break;
      case 36: acc = attn_row_dot<9>(po, pd); break;
      default:
        acc = 0.0f;
        for (int p = 0; p < dh / 4; ++p) {
          const uint2 o = __ldg(po + p), d = __ldg(pd + p);
          acc = fmaf(bf16_lo(o.x), bf16_lo(d.x), acc);
          acc = fmaf(bf16_hi(o.x), bf16_hi(d.x), acc);
          acc = fmaf(bf16_lo(o.y), bf16_lo(d.y), acc);
          acc = fmaf(bf16_hi(o.y), bf16_hi(d.y), acc);
        }
    }
    ox = -lse[(static_cast<size_t>(b) * H + h) * S + q] * 1.4426950408889634f + log2f(drop_scale);
    oy = -acc * scale / drop_scale;
  }
  float* base = stats + (static_cast<size_t>(b) * H + h) * (2 * kSP);
  base[q] = ox;
  base[kSP + q] = oy;
}

// The same statistics with whole-row accesses (dh % 4 == 0, D % 8 == 0, D <= 1024): a CTA takes 32 consecutive query rows
// of one sequence; a warp reads a token row of ctx and dctx with coalesced 16-byte loads (the per-(b, q, h) kernel
// above reads each row as 8-byte pieces 72 bytes apart: every load instruction touches 18 cache lines, nine times
// over), forms the products, and leaves one partial sum per FOUR elements in shared memory -- a head is a whole number
// of those; one thread per (query, head) then adds up its dh / 4 partials and writes x and y, 32 queries = 128 bytes
// per head and plane.
__global__ void __launch_bounds__(256)
attn_bwd_stats_rows_kernel(const __nv_bfloat16* __restrict__ ctx, const __nv_bfloat16* __restrict__ dctx,
                           const float* __restrict__ lse, float* __restrict__ stats, int S, int H, int dh, float scale,
                           float drop_scale) {
  extern __shared__ float hs[];  // [32 rows][D / 4 + 1]
  const int D = H * dh, nq = D >> 2, pitch = nq + 1, nchunks = D >> 3;
  const int b = blockIdx.y, q0 = blockIdx.x * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int ql = warp * 4 + r, q = q0 + ql;
    if (q < S) {  // warp-uniform
      const size_t off = (static_cast<size_t>(b) * S + q) * static_cast<size_t>(D);
      const uint4* po = reinterpret_cast<const uint4*>(ctx + off);
      const uint4* pd = reinterpret_cast<const uint4*>(dctx + off);
      uint4 o[4], d[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunks) { o[i] = __ldg(po + c); d[i] = __ldg(pd + c); }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunks) {
          float lo = bf16_lo(o[i].x) * bf16_lo(d[i].x);
          lo = fmaf(bf16_hi(o[i].x), bf16_hi(d[i].x), lo);
          lo = fmaf(bf16_lo(o[i].y), bf16_lo(d[i].y), lo);
          lo = fmaf(bf16_hi(o[i].y), bf16_hi(d[i].y), lo);
          float hi = bf16_lo(o[i].z) * bf16_lo(d[i].z);
          hi = fmaf(bf16_hi(o[i].z), bf16_hi(d[i].z), hi);
          hi = fmaf(bf16_lo(o[i].w), bf16_lo(d[i].w), hi);
          hi = fmaf(bf16_hi(o[i].w), bf16_hi(d[i].w), hi);
          hs[ql * pitch + 2 * c] = lo;
          hs[ql * pitch + 2 * c + 1] = hi;
        }
      }
    }
  }
  __syncthreads();
  const int per = dh >> 2;
  for (int t = threadIdx.x; t < 32 * H; t += blockDim.x) {
    const int ql = t & 31, h = t >> 5, q = q0 + ql;
    float ox = 0.0f, oy = 0.0f;
    if (q < S) {
      float acc = 0.0f;
      for (int j = 0; j < per; ++j) acc += hs[ql * pitch + h * per + j];
      ox = -lse[(static_cast<size_t>(b) * H + h) * S + q] * 1.4426950408889634f + log2f(drop_scale);
      oy = -acc * scale / drop_scale;
    }
    float* base = stats + (static_cast<size_t>(b) * H + h) * (2 * kSP);
    base[q] = ox;
    base[kSP + q] = oy;
  }
}

template <int NCH, bool DROP>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_q,
                const __grid_constant__ CUtensorMap tm_do, const float* __restrict__ stats,
                const uint32_t* __restrict__ drop_words, __nv_bfloat16* __restrict__ dqkv, int nitems, int S, int H,
                int dh, float scale, float clampv) {
  using G = AttnBwdGeom<NCH>;
  constexpr int DHP = G::DHP, KSTEPS = DHP / 16, RQ = G::RQ, NCG = DHP / 16;
  constexpr bool kZeroTail = (NCH & 1) != 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align_up(smem_raw, 128);
  uint8_t* sK = smem;                       // [2][T128]
  uint8_t* sV = sK + 2 * G::T128;           // [2][T128] (+ one chunk of slack for N = DHP > 8 NCH reads)
  uint8_t* sRing = sV + 2 * G::T128 + G::CS128;
  uint8_t* sdS = sRing + RQ * G::SLOT;      // [2][DSB]
  uint8_t* sOut = sdS + 2 * G::DSB;         // [5][OUTB]: dK_j, dV_j, dQ_0..2
  uint8_t* sStat = sOut + 5 * G::OUTB;      // [3][STB]: the producer runs up to two heads ahead when S <= 128
  uint8_t* sZero = sStat + 3 * G::STB;
  __shared__ AttnBwdBars bars;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int D = H * dh;
  const int nt = (S + 127) / 128;   // key tiles == query tiles per head
  const int nh = 2 * nt;            // 64-row query half-tiles per key tile
  const int nmine = nitems > static_cast<int>(blockIdx.x) ? (nitems - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int nJ = nmine * nt;        // key tiles this CTA walks through
  const int nG = nJ * nh;           // half-tiles this CTA walks through

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.kv_full[i], 1);
      mbar_init(&bars.kv_ready[i], 1);
      mbar_init(&bars.kv_free[i], 1);
      mbar_init(&bars.sdp_full[i], 1);
      mbar_init(&bars.pds_full[i], 8);
      mbar_init(&bars.ts_done[i], 1);
    }
    for (int i = 0; i < 3; ++i) mbar_init(&bars.st_full[i], 1);
    for (int i = 0; i < 6; ++i) {
      mbar_init(&bars.qd_full[i], 1);
      mbar_init(&bars.qd_free[i], 1);
    }
    mbar_init(&bars.acc_full, 1);
    mbar_init(&bars.acc_free, 8);
    for (int i = 0; i < 5; ++i) {
      mbar_init(&bars.out_full[i], 4);
      mbar_init(&bars.out_free[i], 1);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 2048 / 16; i += kBwdThreads) reinterpret_cast<uint4*>(sZero)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 16) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // region r: S^T at r*128 + [0, 64), dP^T at r*128 + [64, 128); accumulators behind
  const uint32_t tdV = tmem + 256, tdK = tmem + 256 + DHP, tdQ = tmem + 256 + 2 * DHP;

  if (warp == 17) {
    // ------------------------------------------------------------------ TMA producer (+ K/V fix-up)
    if (lane == 0) {
      tma_prefetch_desc(&tm_kv);
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_do);
    }
    const bool need_fix = (dh & 7) != 0;
    // All 32 lanes wait (a lone polling lane wakes up late -- see the note in wm_common.cuh), lane 0 issues.
    auto load_kv = [&](int J) {  // K_j, V_j of key tile J (and the head's statistics with its first tile)
      const int n = J / nt, j = J - n * nt;
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / H, h = item - b * H;
      const int col0 = (h * dh) & ~7;
      const int kb = J & 1;
      if (J >= 2) mbar_wait(&bars.kv_free[kb], ((J >> 1) - 1) & 1, 80);
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars.kv_full[kb], 2 * NCH * 2048);
        tma_load_4d(sK + kb * G::T128, &tm_kv, &bars.kv_full[kb], 0, j * 128, (D + col0) >> 3, b);
        tma_load_4d(sV + kb * G::T128, &tm_kv, &bars.kv_full[kb], 0, j * 128, (2 * D + col0) >> 3, b);
        if (j == 0) {
          mbar_arrive_expect_tx(&bars.st_full[n % 3], G::STB);
          bulk_load_1d(sStat + (n % 3) * G::STB, stats + static_cast<size_t>(item) * (2 * kSP), G::STB, &bars.st_full[n % 3]);
        }
      }
      __syncwarp();
    };
    auto fix_kv = [&](int J) {  // whole warp: zero the neighbouring head's 4 columns of K_j and V_j, publish
      const int n = J / nt;
      const int item = blockIdx.x + n * gridDim.x;
      const int h = item % H;
      const int kb = J & 1;
      mbar_wait(&bars.kv_full[kb], (J >> 1) & 1, 81);
      if (need_fix) {
        const int z0 = ((h * dh) & 7) ? 0 : dh;
        const uint32_t zoff = (z0 >> 3) * G::CS128 + (z0 & 7) * 2;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          *reinterpret_cast<uint2*>(sK + kb * G::T128 + zoff + (r * 32 + lane) * 16) = make_uint2(0u, 0u);
          *reinterpret_cast<uint2*>(sV + kb * G::T128 + zoff + (r * 32 + lane) * 16) = make_uint2(0u, 0u);
        }
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.kv_ready[kb]);
    };
    if (nJ > 0) {
      load_kv(0);
      fix_kv(0);
    }
    const int ih_load = RQ < nh - 1 ? RQ : nh - 1;  // late enough that kv_free of key tile J - 1 has been committed
    int g = 0;
    for (int J = 0; J < nJ; ++J) {
      const int n = J / nt;
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / H, h = item - b * H;
      const int col0 = (h * dh) & ~7;
      for (int ih = 0; ih < nh; ++ih, ++g) {
        if (ih == ih_load && J + 1 < nJ) load_kv(J + 1);
        const int slot = g % RQ;
        if (g >= RQ) mbar_wait(&bars.qd_free[slot], ((g / RQ) - 1) & 1, 82);
        if (lane == 0) {
          uint8_t* dst = sRing + slot * G::SLOT;
          mbar_arrive_expect_tx(&bars.qd_full[slot], 2 * NCH * 1024 + (DROP ? 1024 : 0));
          tma_load_4d(dst, &tm_q, &bars.qd_full[slot], 0, ih * 64, col0 >> 3, b);
          tma_load_4d(dst + G::T64, &tm_do, &bars.qd_full[slot], 0, ih * 64, col0 >> 3, b);
          if (DROP) {  // [4 key slices][64 query rows] keep words of this half-tile: 1 KB, contiguous
            const int j = J - n * nt;
            bulk_load_1d(dst + 2 * G::T64, drop_words + attn_drop_word_index(item, j, ih, 0, 0), 1024, &bars.qd_full[slot]);
          }
        }
        __syncwarp();
        if (ih == nh - 1 && J + 1 < nJ) fix_kv(J + 1);
      }
    }
  } else if (warp == 20) {
    // ------------------------------------------------------------------ MMA issue: S^T = K_j Q_ih^T, dP^T = V_j dO_ih^T
    const uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
    const uint32_t zero_addr = smem_u32(sZero);
    for (int g2 = 0; g2 < nG; ++g2) {
      const int r = g2 & 1;
      const int J2 = g2 / nh;
      const int kb = J2 & 1, slot = g2 % RQ;
      // Region r is free once the products that read half-tile g2 - 2's P^T / dS^T have run. The second region
      // starts half a period late (after the first elementwise pass): started together, the two elementwise groups
      // stay in phase -- both fight for issue slots, then both wait for the tensor pipe (2x slower, measured).
      if (g2 >= 2) mbar_wait(&bars.ts_done[r], ((g2 - 2) >> 1) & 1, 96);
      if (g2 == 1) mbar_wait(&bars.pds_full[0], 0, 97);
      if (g2 - J2 * nh == 0) mbar_wait(&bars.kv_ready[kb], (J2 >> 1) & 1, 83);  // first half-tile of a key tile
      mbar_wait(&bars.qd_full[slot], (g2 / RQ) & 1, 84);
      tc_fence_after();
      const uint32_t ka = smem_u32(sK + kb * G::T128), va = smem_u32(sV + kb * G::T128);
      const uint32_t qa = smem_u32(sRing + slot * G::SLOT), da = qa + G::T64;
      const uint32_t tS = tmem + r * 128;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const uint32_t a = which ? va : ka, bq = which ? da : qa;
        const uint64_t ad = umma_smem_desc(a, G::CS128, 128, UMMA_SWZ_NONE), bd = umma_smem_desc(bq, G::CS64, 128, UMMA_SWZ_NONE);
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          if (kZeroTail && ks == KSTEPS - 1) {
            const uint32_t a0 = a + ks * 2 * G::CS128, b0 = bq + ks * 2 * G::CS64;
            umma_ss_warp(tS + which * 64, umma_smem_desc(a0, zero_addr - a0, 128, UMMA_SWZ_NONE),
                         umma_smem_desc(b0, zero_addr - b0, 128, UMMA_SWZ_NONE), idesc_s, ks != 0);
          } else {
            umma_ss_warp(tS + which * 64, ad + ks * ((2 * G::CS128) >> 4), bd + ks * ((2 * G::CS64) >> 4), idesc_s, ks != 0);
          }
        }
      }
      umma_commit_warp(&bars.sdp_full[r]);
    }
  } else if (warp == 16) {
    // ------------------------------------------------------------------ MMA issue: gradient products
    const uint32_t idesc_kv = umma_idesc_bf16(128, DHP, 0, 1);  // A = P^T / dS^T from TMEM, B MN-major
    const uint32_t idesc_q = umma_idesc_bf16(128, DHP, 1, 1);   // A = dS (MN-major, smem), B = K_j MN-major
    int g = 0;
    for (int J = 0; J < nJ; ++J) {
      const int j = J % nt;
      const int kb = J & 1;
      const uint32_t ka = smem_u32(sK + kb * G::T128);
      for (int ih = 0; ih < nh; ++ih, ++g) {
        const int r = g & 1, slot = g % RQ;
        const uint32_t tS = tmem + r * 128;
        const uint32_t qa = smem_u32(sRing + slot * G::SLOT), da = qa + G::T64;
        mbar_wait(&bars.pds_full[r], (g >> 1) & 1, 85);  // P^T, dS^T in TMEM; dS half in shared memory
        if (ih == 0 && J >= 1) mbar_wait(&bars.acc_free, (J - 1) & 1, 86);  // previous dK/dV (dQ) drained
        tc_fence_after();
        // dV_j += P^T dO_ih, dK_j += dS^T Q_ih: k = 64 query rows, A k-step = 8 packed columns of the slice's first 16
        {
          const uint64_t dod = umma_smem_desc(da, 128, G::CS64, UMMA_SWZ_NONE), qd = umma_smem_desc(qa, 128, G::CS64, UMMA_SWZ_NONE);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t acol = (ks >> 1) * 32 + (ks & 1) * 8;
            umma_ts_warp(tdV, tS + acol, dod + ks * (256 >> 4), idesc_kv, (ih | ks) != 0);
            umma_ts_warp(tdK, tS + 64 + acol, qd + ks * (256 >> 4), idesc_kv, (ih | ks) != 0);
          }
        }
        umma_commit_warp(&bars.ts_done[r]);     // region r may take the next-but-one half-tile's scores
        umma_commit_warp(&bars.qd_free[slot]);  // last readers of this half-tile's Q / dO
        if (ih & 1) {  // both halves of query tile i are in the dS buffer: dQ_i += dS_i K_j (k = 128 keys)
          const int i = ih >> 1, tb = (g >> 1) & 1;
          const uint64_t sd = umma_smem_desc(smem_u32(sdS + tb * G::DSB), 128, 2048, UMMA_SWZ_NONE);
          const uint64_t kd = umma_smem_desc(ka, 128, G::CS128, UMMA_SWZ_NONE);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_ss_warp(tdQ + i * DHP, sd + ks * (256 >> 4), kd + ks * (256 >> 4), idesc_q, (j | ks) != 0);
        }
        if (ih == nh - 1) {
          umma_commit_warp(&bars.kv_free[kb]);
          umma_commit_warp(&bars.acc_full);
        }
      }
    }
  } else if (warp >= 18) {  // (warps 18, 19)
    // ------------------------------------------------------------------ gradient store (two warps)
    // warp 18: dK_j (slot 0), dQ_0, dQ_2 (slots 2, 4); warp 19: dV_j (slot 1), dQ_1 (slot 3)
    const int pv = dh >> 2, sw = warp - 18;
    auto drain = [&](int slot, uint32_t parity, __nv_bfloat16* gbase, int nrows) {
      mbar_wait(&bars.out_full[slot], parity, 87);
      store_staged_tile(sOut + slot * G::OUTB, gbase, 3 * static_cast<size_t>(D), nrows, pv, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.out_free[slot]);
    };
    for (int J = 0; J < nJ; ++J) {
      const int n = J / nt, j = J - n * nt;
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / H, h = item - b * H;
      __nv_bfloat16* hb = dqkv + static_cast<size_t>(b) * S * (3 * D) + h * dh;
      drain(sw, J & 1, hb + static_cast<size_t>(j) * 128 * (3 * D) + (1 + sw) * D, min(128, S - j * 128));
      if (j == nt - 1)
        for (int i = sw; i < nt; i += 2) drain(2 + i, n & 1, hb + static_cast<size_t>(i) * 128 * (3 * D), min(128, S - i * 128));
    }
  } else {
    // ------------------------------------------------------------------ elementwise warps
    const int grp = warp >> 3;                 // TMEM region / half-tile parity owned by this group
    const int lq = warp & 3, h2 = (warp >> 2) & 1;
    const int krow = lq * 32 + lane;           // key row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(lq * 32) << 16;
    const float c2 = scale * 1.4426950408889634f;
    const uint64_t c2p = f2_pack(c2, c2), scalep = f2_pack(scale, scale);
    const uint32_t lanebit = 1u << attn_keep_bit(lane);  // this thread's key inside a keep word of the forward kernel
    // clampv = log2(drop_scale), computed on the host: log2f() in here put its zero / denormal special case (an FSEL
    // per element) into the inner loop
    const uint32_t tS = tmem + grp * 128 + lane_sel;
    // TMEM accumulator columns [16 cg, 16 cg + 16) of this thread's row -> bf16 staging row (compact [128, dh])
    auto stage_acc = [&](uint32_t tacc, uint8_t* srow, int front) {
#pragma unroll
      for (int cg = 0; cg < NCG; ++cg) {
        uint32_t v[16];
        tmem_ld16(tacc + lane_sel + cg * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 16; jj += 4) {
          const int d = cg * 16 + jj - front;
          if (d >= 0 && d < dh) {
            uint2 o2;
            o2.x = pack_bf16x2(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1]));
            o2.y = pack_bf16x2(__uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
            *reinterpret_cast<uint2*>(srow + d * 2) = o2;
          }
        }
      }
    };
    for (int g = grp; g < nG; g += 2) {
      const int J = g / nh, ih = g - J * nh;
      const int n = J / nt, j = J - n * nt;
      const int slot = g % RQ;
      const int tb = (g >> 1) & 1;  // dS buffer of this query tile
      const float* stx = reinterpret_cast<const float*>(sStat + (n % 3) * G::STB) + ih * 64 + h2 * 32;  // -lse log2e + log2(drop_scale)
      const float* sty = stx + kSP;                                                                      // -delta scale / drop_scale
      const uint32_t* mw = reinterpret_cast<const uint32_t*>(sRing + slot * G::SLOT + 2 * G::T64) + lq * 64 + h2 * 32;
      if (j == 0 && ih < 2) mbar_wait(&bars.st_full[n % 3], (n / 3) & 1, 88);
      mbar_wait(&bars.sdp_full[grp], (g >> 1) & 1, 89);
      tc_fence_after();
      if (warp == 0 && n == 1) WM_TICK(g - nt * nh);
      // No wait for the dropout words (same transaction barrier as Q/dO, which the score-issue warp observed before
      // it queued this half-tile's scores) nor for the dS buffer: its previous reader, the dQ product of tile
      // (g >> 1) - 2, was queued by the product-issue warp before the products whose completion (ts_done) released
      // this region to the score-issue warp, and tcgen05.mma of one thread completes in order.
      // dS tile layout: [q chunk of 8][128 keys][16 B]; this thread fills key row krow of chunks (ih&1)*8 + h2*4 + 0..3
      uint8_t* dsrow = sdS + tb * G::DSB + ((ih & 1) * 8 + h2 * 4) * 2048 + krow * 16;
#pragma unroll
      for (int bt = 0; bt < 2; ++bt) {
        uint32_t vs[16], vd[16];
        tmem_ld16(tS + h2 * 32 + bt * 16, vs);
        tmem_ld16(tS + 64 + h2 * 32 + bt * 16, vd);
        tmem_ld_wait();
        uint32_t pk[8], dk[8];
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) {  // four query rows at a time: two 16-byte statistics loads, one of dropout words
          const int e0 = bt * 16 + w4 * 4;
          const float4 lx4 = *reinterpret_cast<const float4*>(stx + e0);
          const float4 nd4 = *reinterpret_cast<const float4*>(sty + e0);
          const float lx[4] = {lx4.x, lx4.y, lx4.z, lx4.w};
          const float nd[4] = {nd4.x, nd4.y, nd4.z, nd4.w};
          uint32_t mword[4] = {0u, 0u, 0u, 0u};
          if (DROP) {
            const uint4 m4 = *reinterpret_cast<const uint4*>(mw + e0);
            mword[0] = m4.x; mword[1] = m4.y; mword[2] = m4.z; mword[3] = m4.w;
          }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {  // packed fp32 pairs: query rows e0 + 2hh, e0 + 2hh + 1
            const int i0 = w4 * 4 + 2 * hh;
            float xa, xb;
            f2_unpack(f2_fma(f2_pack_u(vs[i0], vs[i0 + 1]), c2p, f2_pack(lx[2 * hh], lx[2 * hh + 1])), xa, xb);
            const float pa = fast_exp2(fminf(xa, clampv)), pb = fast_exp2(fminf(xb, clampv));
            const uint64_t p2 = f2_pack(pa, pb);
            const uint64_t nd2 = f2_pack(nd[2 * hh], nd[2 * hh + 1]);
            const uint64_t vd2 = f2_pack_u(vd[i0], vd[i0 + 1]);
            float da, db;
            if (DROP) {  // dS = P (keep dP scale - delta) = (keep P) (dP scale) + P (-delta): one select per element
              const float qa = (mword[2 * hh] & lanebit) ? pa : 0.0f, qb = (mword[2 * hh + 1] & lanebit) ? pb : 0.0f;
              f2_unpack(f2_fma(f2_pack(qa, qb), f2_mul(vd2, scalep), f2_mul(p2, nd2)), da, db);
              pk[w4 * 2 + hh] = attn_pack2(qa, qb);
            } else {
              f2_unpack(f2_mul(p2, f2_fma(vd2, scalep, nd2)), da, db);
              pk[w4 * 2 + hh] = attn_pack2(pa, pb);
            }
            dk[w4 * 2 + hh] = attn_pack2(da, db);
          }
        }
        tmem_st8(tS + h2 * 32 + bt * 8, pk);        // P^T: 16 query rows = 8 packed columns
        tmem_st8(tS + 64 + h2 * 32 + bt * 8, dk);   // dS^T
        *reinterpret_cast<uint4*>(dsrow + (bt * 2) * 2048) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
        *reinterpret_cast<uint4*>(dsrow + (bt * 2 + 1) * 2048) = make_uint4(dk[4], dk[5], dk[6], dk[7]);
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.pds_full[grp]);
      if (warp == 0 && n == 1) WM_TICK(g - nt * nh + 1);
      // ---- drains (group 1 handles the last half-tile of every key tile)
      if (ih == nh - 1) {
        const int item = blockIdx.x + n * gridDim.x;
        const int front = ((item % H) * dh) & 7;
        mbar_wait(&bars.acc_full, J & 1, 92);
        tc_fence_after();
        // dK_j -> slot 0 (warps with h2 == 0), dV_j -> slot 1 (h2 == 1); thread = key row
        if (J >= 1) mbar_wait(&bars.out_free[h2], (J - 1) & 1, 93);
        stage_acc(h2 ? tdV : tdK, sOut + h2 * G::OUTB + krow * (dh * 2), front);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.out_full[h2]);
        if (j == nt - 1) {  // dQ tiles of the head: thread = query row; h2 == 0 takes tiles 0 and 2, h2 == 1 tile 1
          for (int i = h2; i < nt; i += 2) {
            if (n >= 1) mbar_wait(&bars.out_free[2 + i], (n - 1) & 1, 94);
            stage_acc(tdQ + i * DHP, sOut + (2 + i) * G::OUTB + krow * (dh * 2), front);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.out_full[2 + i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.acc_free);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static int attn_sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  return sms;
}
// 16-byte chunks that cover one head's columns from the 8-column-aligned window start (worst-case offset 4)
static int attn_chunks(int dh) { return (dh & 7) ? (dh + 4) / 8 : dh / 8; }

template <int NCH>
static int launch_fwd_t(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, float* lse, uint32_t* drop_words, int B, int S,
                        int H, int dh, float scale, uint32_t thresh15, float dscale, uint64_t seed, uint64_t stream_id,
                        cudaStream_t stream) {
  CUtensorMap tm;
  const int D = H * dh;
  int rc = make_tmap_bf16_rows3d(&tm, qkv, 3 * D, S, B, 3 * D, 128);
  if (rc != WM_OK) return rc;
  const int smem = AttnFwdGeom<NCH>::kSmem;
  auto kern = thresh15 ? attn_fwd_kernel<NCH, true> : attn_fwd_kernel<NCH, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return WM_ERR_CUDA;
  const int nitems = B * H;
  const int grid = nitems < attn_sm_count() ? nitems : attn_sm_count();
  if (thresh15 && static_cast<uint64_t>(nitems) * S * ((S + 15) / 16) * 4ull > 0xFFFFFFFFull) return WM_ERR_SHAPE;  // 32-bit mask counters
  kern<<<grid, kFwdThreads, smem, stream>>>(tm, ctx, lse, drop_words, nitems, S, H, dh, scale, thresh15, dscale,
                                            drop_keys(seed, stream_id));
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}
template <int NCH>
static int launch_bwd_t(const __nv_bfloat16* qkv, const __nv_bfloat16* ctx, const __nv_bfloat16* dctx,
                        const float* lse, __nv_bfloat16* dqkv, const uint32_t* drop_words, float* stats, int B, int S,
                        int H, int dh, float scale, bool drop, float dscale, cudaStream_t stream) {
  const int D = H * dh;
  CUtensorMap tm_kv, tm_q, tm_do;
  int rc = make_tmap_bf16_chunked4d(&tm_kv, qkv, 3 * D, S, B, 3 * D, 128, NCH);
  if (rc == WM_OK) rc = make_tmap_bf16_chunked4d(&tm_q, qkv, 3 * D, S, B, 3 * D, 64, NCH);
  if (rc == WM_OK) rc = make_tmap_bf16_chunked4d(&tm_do, dctx, D, S, B, D, 64, NCH);
  if (rc != WM_OK) return rc;
  {
    const long long total = static_cast<long long>(B) * kSP * H;
    if ((dh & 3) == 0 && (D & 7) == 0 && D <= 1024 && (reinterpret_cast<uintptr_t>(ctx) & 15u) == 0 &&
        (reinterpret_cast<uintptr_t>(dctx) & 15u) == 0 && B <= 65535) {
      const int smem_stats = 32 * (D / 4 + 1) * static_cast<int>(sizeof(float));
      attn_bwd_stats_rows_kernel<<<dim3(kSP / 32, B), 256, smem_stats, stream>>>(ctx, dctx, lse, stats, S, H, dh, scale, dscale);
    } else {
      attn_bwd_stats_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(ctx, dctx, lse, stats, B, S, H, dh,
                                                                                           scale, dscale);
    }
    WM_COUNT_LAUNCH();
  }
  const int smem = AttnBwdGeom<NCH>::kSmem;
  auto kern = drop ? attn_bwd_kernel<NCH, true> : attn_bwd_kernel<NCH, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return WM_ERR_CUDA;
  const int nitems = B * H;
  const int grid = nitems < attn_sm_count() ? nitems : attn_sm_count();
  kern<<<grid, kBwdThreads, smem, stream>>>(tm_kv, tm_q, tm_do, stats, drop_words, dqkv, nitems, S, H, dh, scale,
                                            log2f(dscale));
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// drop_thresh is the 16-bit threshold of the C ABI (round(p*65536)); every kernel compares 15 bits (wm_common.cuh)
static void attn_drop_params(uint32_t drop_thresh16, uint32_t* thresh15, float* scale) {
  *thresh15 = drop_thresh15(drop_thresh16);
  *scale = drop_keep_scale(drop_thresh16);
}
// TMA needs 16-byte aligned row pitches and head-block starts: D = H * dh a multiple of 8
static bool attn_shape_ok(int B, int S, int H, int dh) {
  return B > 0 && H > 0 && S > 0 && S <= kSP && !(dh & 3) && dh >= 12 && dh <= 48 && !((H * dh) & 7);
}

size_t attn_dropout_words_bytes(int B, int S, int H) {
  (void)S;
  return static_cast<size_t>(B) * H * kAttnMaskSlices * kSP * sizeof(uint32_t);
}
size_t attn_bwd_workspace_bytes(int B, int S, int H) {
  (void)S;
  return static_cast<size_t>(B) * H * kSP * sizeof(float2);
}

int launch_attn_fwd(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, float* lse, uint32_t* drop_words, int B, int S, int H,
                    int dh, uint32_t drop_thresh, uint64_t seed, uint64_t stream_id, cudaStream_t stream) {
  if (!attn_shape_ok(B, S, H, dh)) return WM_ERR_SHAPE;
  uint32_t t15;
  float ds;
  attn_drop_params(drop_thresh, &t15, &ds);
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  switch (attn_chunks(dh)) {
    case 2: return launch_fwd_t<2>(qkv, ctx, lse, drop_words, B, S, H, dh, scale, t15, ds, seed, stream_id, stream);
    case 3: return launch_fwd_t<3>(qkv, ctx, lse, drop_words, B, S, H, dh, scale, t15, ds, seed, stream_id, stream);
    case 4: return launch_fwd_t<4>(qkv, ctx, lse, drop_words, B, S, H, dh, scale, t15, ds, seed, stream_id, stream);
    case 5: return launch_fwd_t<5>(qkv, ctx, lse, drop_words, B, S, H, dh, scale, t15, ds, seed, stream_id, stream);
    case 6: return launch_fwd_t<6>(qkv, ctx, lse, drop_words, B, S, H, dh, scale, t15, ds, seed, stream_id, stream);
    default: return WM_ERR_SHAPE;
  }
}

// drop_words: the buffer the forward call filled (required when drop_thresh rounds to a non-zero probability);
// workspace: attn_bwd_workspace_bytes(B, S, H) bytes, 16-byte aligned.
int launch_attn_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* ctx, const __nv_bfloat16* dctx, const float* lse,
                    __nv_bfloat16* dqkv, const uint32_t* drop_words, void* workspace, int B, int S, int H, int dh,
                    uint32_t drop_thresh, cudaStream_t stream) {
  if (!attn_shape_ok(B, S, H, dh)) return WM_ERR_SHAPE;
  uint32_t t15;
  float ds;
  attn_drop_params(drop_thresh, &t15, &ds);
  if (!workspace || (t15 && !drop_words)) return WM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15u) || (reinterpret_cast<uintptr_t>(drop_words) & 15u)) return WM_ERR_ALIGN;
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  float* st = static_cast<float*>(workspace);
  switch (attn_chunks(dh)) {
    case 2: return launch_bwd_t<2>(qkv, ctx, dctx, lse, dqkv, drop_words, st, B, S, H, dh, scale, t15 != 0, ds, stream);
    case 3: return launch_bwd_t<3>(qkv, ctx, dctx, lse, dqkv, drop_words, st, B, S, H, dh, scale, t15 != 0, ds, stream);
    case 4: return launch_bwd_t<4>(qkv, ctx, dctx, lse, dqkv, drop_words, st, B, S, H, dh, scale, t15 != 0, ds, stream);
    case 5: return launch_bwd_t<5>(qkv, ctx, dctx, lse, dqkv, drop_words, st, B, S, H, dh, scale, t15 != 0, ds, stream);
    case 6: return launch_bwd_t<6>(qkv, ctx, dctx, lse, dqkv, drop_words, st, B, S, H, dh, scale, t15 != 0, ds, stream);
    default: return WM_ERR_SHAPE;
  }
}

}  // namespace wm
