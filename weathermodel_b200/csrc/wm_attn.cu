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
// (keep iff (byte & 0x7F) >= thresh7), regenerated identically in backward.
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
// query row, one byte each (7 bits used).
constexpr int kAttnPhiloxRounds = 7;
// keep iff (byte & 0x7F) >= thresh7: adding (128 - thresh7) to the 7-bit value carries into bit 7 of the byte exactly
// then (no carry crosses a byte). add4 = (128 - thresh7) * 0x01010101; f[w] carries the keep flag of element
// 4w + b in bit 7 of byte b (the other bits are noise) -- two integer ops per four elements.
WM_DEVICE void keep_flags16(uint64_t seed, uint64_t stream, uint64_t grp, uint32_t add4, uint32_t (&f)[4]) {
  const Philox4 r = philox4x32<kAttnPhiloxRounds>(seed, stream, grp);
  f[0] = (r.x & 0x7F7F7F7Fu) + add4;
  f[1] = (r.y & 0x7F7F7F7Fu) + add4;
  f[2] = (r.z & 0x7F7F7F7Fu) + add4;
  f[3] = (r.w & 0x7F7F7F7Fu) + add4;
}
// 0xFFFF / 0x0000 in each half of the result: keep flags of elements j (low half) and j + 1 (high half), j even,
// for masking a packed bf16x2 pair -- one PRMT in sign-replicate mode (selector nibble bit 3)
WM_DEVICE uint32_t keep_pair_mask(uint32_t fword, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(fword), "r"(0u), "r"(sel));
  return d;
}
#define WM_PAIR_SEL(j) ((((j) & 3) | 8u) * 0x11u | ((((j) & 3) + 1u) | 8u) * 0x1100u)
#define WM_KEEP_PAIR(f, j) keep_pair_mask((f)[(j) >> 2], WM_PAIR_SEL(j))
// byte-mask form (0xFF in every kept byte) used by the v4 backward kernel
WM_DEVICE void keep_masks16(uint64_t seed, uint64_t stream, uint64_t grp, uint32_t add4, uint32_t (&m)[4]) {
  uint32_t f[4];
  keep_flags16(seed, stream, grp, add4, f);
#pragma unroll
  for (int w = 0; w < 4; ++w) m[w] = ((f[w] >> 7) & 0x01010101u) * 0xFFu;
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
constexpr int kFwdSoftmaxWarps = 12;
constexpr int kFwdThreads = 32 * 15;
constexpr int kKC = 96;  // keys per score chunk (3 slices of 32 columns)

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
  static constexpr uint32_t kSmem = 3 * QT + 2 * KVB * KT + CSK + 2 * OUTB + 2048 + 2 * 3 * 128 * 4 + 128;
};

template <int NCH, bool DROP>
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ ctx,
                float* __restrict__ lse_out, int nitems, int S, int H, int dh, float scale, uint32_t thresh7,
                float drop_scale, uint64_t seed, uint64_t stream_id) {
  using G = AttnFwdGeom<NCH>;
  constexpr int DHP = G::DHP, KVB = G::KVB, KSTEPS = DHP / 16;
  constexpr bool kZeroTail = (NCH & 1) != 0;  // last k-step: second core-matrix column comes from the zero chunk
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 3 * G::QT;
  uint8_t* sV = sK + KVB * G::KT;
  uint8_t* sOut = sV + KVB * G::KT + G::CSK;
  uint8_t* sZero = sOut + 2 * G::OUTB;
  float* sMax = reinterpret_cast<float*>(sZero + 2048);  // [3][128]
  float* sSum = sMax + 3 * 128;                          // [3][128]
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
  if (warp == 12) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem, tO = tmem + kSP;  // O buffers at +0 / +64

  if (warp == 13) {
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
  } else if (warp == 12) {
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
  } else if (warp == 14) {
    // ------------------------------------------------------------------ ctx store
    const int pv = dh >> 2;  // 8-byte pieces per row
    const float inv_pv = 1.0f / static_cast<float>(pv);
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
        for (int idx = lane; idx < nrows * pv; idx += 32) {
          const int r = static_cast<int>((static_cast<float>(idx) + 0.5f) * inv_pv);
          const int pp = idx - r * pv;
          *reinterpret_cast<uint2*>(obase + static_cast<size_t>(r) * D + pp * 4) = *reinterpret_cast<const uint2*>(src + idx * 8);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.out_free[ob]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int lq = warp & 3, sl = warp >> 2;   // TMEM lane quarter, 32-column slice of each chunk
    const int row = lq * 32 + lane;            // query row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(lq * 32) << 16;
    const uint32_t thresh4 = (128u - thresh7) * 0x01010101u;  // per-byte addend of the 7-bit keep test
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
        const uint64_t rowbase = (static_cast<uint64_t>(item) * S + (q < S ? q : 0)) * nk16;
        mbar_wait(&bars.s_full, t & 1, 72);
        tc_fence_after();
        WM_FTICK(2);
        // ---- pass 1: row max over this warp's slices (next chunk's TMEM load in flight under the reduction)
        float mloc = -INFINITY;
        {
          uint32_t va[32], vb[32];
          tmem_ld32(tS + lane_sel + sl * 32, va);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < nkc) {
              uint32_t(&cur)[32] = (c & 1) ? vb : va;
              uint32_t(&nxt)[32] = (c & 1) ? va : vb;
              tmem_ld_wait();
              if (c + 1 < nkc) tmem_ld32(tS + lane_sel + (c + 1) * kKC + sl * 32, nxt);
              const int k0 = c * kKC + sl * 32;
              if (k0 + 32 <= S) {
#pragma unroll
                for (int j = 0; j < 32; ++j) mloc = fmaxf(mloc, __uint_as_float(cur[j]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (k0 + j < S) mloc = fmaxf(mloc, __uint_as_float(cur[j]));
              }
            }
          }
        }
        sMax[sl * 128 + row] = mloc;
        WM_FTICK(3);
        named_bar_sync(1 + lq, 96);
        WM_FTICK(4);
        const float mrow = fmaxf(fmaxf(sMax[row], sMax[128 + row]), sMax[256 + row]);
        const float mneg = -mrow * c2;
        // ---- pass 2: exp2, row sum, dropout; P (bf16 pairs) replaces the first 16 columns of the slice in TMEM
        float lsum = 0.0f;
        {
          uint32_t va[32], vb[32];
          tmem_ld32(tS + lane_sel + sl * 32, va);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < nkc) {
              uint32_t(&cur)[32] = (c & 1) ? vb : va;
              uint32_t(&nxt)[32] = (c & 1) ? va : vb;
              tmem_ld_wait();
              if (c + 1 < nkc) tmem_ld32(tS + lane_sel + (c + 1) * kKC + sl * 32, nxt);
              const int k0 = c * kKC + sl * 32;
              uint32_t pk[16];
              if (k0 + 32 <= S) {
#pragma unroll
                for (int w = 0; w < 16; ++w) {
                  const float e0 = fast_exp2(fmaf(__uint_as_float(cur[2 * w]), c2, mneg));
                  const float e1 = fast_exp2(fmaf(__uint_as_float(cur[2 * w + 1]), c2, mneg));
                  lsum += e0 + e1;
                  pk[w] = pack_bf16x2(e0, e1);
                }
              } else {
#pragma unroll
                for (int w = 0; w < 16; ++w) {
                  float e0 = 0.0f, e1 = 0.0f;
                  if (k0 + 2 * w < S) e0 = fast_exp2(fmaf(__uint_as_float(cur[2 * w]), c2, mneg));
                  if (k0 + 2 * w + 1 < S) e1 = fast_exp2(fmaf(__uint_as_float(cur[2 * w + 1]), c2, mneg));
                  lsum += e0 + e1;
                  pk[w] = pack_bf16x2(e0, e1);
                }
              }
              if (DROP) {
                uint32_t fa[4], fb[4];
                keep_flags16(seed, stream_id, rowbase + (k0 >> 4), thresh4, fa);
                keep_flags16(seed, stream_id, rowbase + (k0 >> 4) + 1, thresh4, fb);
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                  pk[w] &= WM_KEEP_PAIR(fa, 2 * w);
                  pk[w + 8] &= WM_KEEP_PAIR(fb, 2 * w);
                }
              }
              tmem_st16(tS + lane_sel + k0, pk);
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&bars.p_full[c]);
              WM_FTICK(5 + c);
            }
          }
        }
        sSum[sl * 128 + row] = lsum;
        named_bar_sync(1 + lq, 96);
        WM_FTICK(9);
        const float tot = sSum[row] + sSum[128 + row] + sSum[256 + row];
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
  if (warp == 12) {
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
                __nv_bfloat16* __restrict__ dqkv, int S, int H, int dh, float scale, uint32_t thresh7,
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
  const uint32_t thresh4 = (128u - thresh7) * 0x01010101u;  // per-byte addend of the 7-bit keep test

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
static int launch_fwd_t(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, float* lse, int B, int S, int H, int dh,
                        float scale, uint32_t thresh7, float dscale, uint64_t seed, uint64_t stream_id,
                        cudaStream_t stream) {
  CUtensorMap tm;
  const int D = H * dh;
  int rc = make_tmap_bf16_rows3d(&tm, qkv, 3 * D, S, B, 3 * D, 128);
  if (rc != WM_OK) return rc;
  const int smem = AttnFwdGeom<NCH>::kSmem;
  auto kern = thresh7 ? attn_fwd_kernel<NCH, true> : attn_fwd_kernel<NCH, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return WM_ERR_CUDA;
  const int nitems = B * H;
  const int grid = nitems < attn_sm_count() ? nitems : attn_sm_count();
  kern<<<grid, kFwdThreads, smem, stream>>>(tm, ctx, lse, nitems, S, H, dh, scale, thresh7, dscale, seed, stream_id);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}
template <int DHP>
static int launch_bwd_t(const __nv_bfloat16* qkv, const __nv_bfloat16* ctx, const __nv_bfloat16* dctx,
                        const float* lse, __nv_bfloat16* dqkv, int B, int S, int H, int dh, float scale,
                        uint32_t thresh7, float dscale, uint64_t seed, uint64_t stream_id, cudaStream_t stream) {
  const int smem = 4 * kSP * DHP * 2 + 2 * 128 * 128 * 2 + 2 * kSP * 4 + 256;
  auto kern = thresh7 ? attn_bwd_kernel<DHP, true> : attn_bwd_kernel<DHP, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return WM_ERR_CUDA;
  kern<<<B * H, kBwdThreads, smem, stream>>>(qkv, ctx, dctx, lse, dqkv, S, H, dh, scale, thresh7, dscale, seed,
                                             stream_id);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// drop_thresh is the 16-bit threshold used everywhere else (round(p*65536)); attention rounds it to 7 bits
static void attn_drop_params(uint32_t drop_thresh16, uint32_t* thresh7, float* scale) {
  *thresh7 = (drop_thresh16 + 256u) >> 9;
  *scale = *thresh7 ? 128.0f / static_cast<float>(128u - *thresh7) : 1.0f;
}

int launch_attn_fwd(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, float* lse, int B, int S, int H, int dh,
                    uint32_t drop_thresh, float drop_scale, uint64_t seed, uint64_t stream_id, cudaStream_t stream) {
  (void)drop_scale;
  // TMA needs 16-byte aligned row pitches and head-block starts: D = H * dh a multiple of 8
  if (B <= 0 || H <= 0 || S <= 0 || S > kSP || (dh & 3) || dh < 8 || dh > 48 || ((H * dh) & 7)) return WM_ERR_SHAPE;
  uint32_t t8;
  float ds;
  attn_drop_params(drop_thresh, &t8, &ds);
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  switch (attn_chunks(dh)) {
    case 1: return WM_ERR_SHAPE;
    case 2: return launch_fwd_t<2>(qkv, ctx, lse, B, S, H, dh, scale, t8, ds, seed, stream_id, stream);
    case 3: return launch_fwd_t<3>(qkv, ctx, lse, B, S, H, dh, scale, t8, ds, seed, stream_id, stream);
    case 4: return launch_fwd_t<4>(qkv, ctx, lse, B, S, H, dh, scale, t8, ds, seed, stream_id, stream);
    case 5: return launch_fwd_t<5>(qkv, ctx, lse, B, S, H, dh, scale, t8, ds, seed, stream_id, stream);
    default: return launch_fwd_t<6>(qkv, ctx, lse, B, S, H, dh, scale, t8, ds, seed, stream_id, stream);
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
