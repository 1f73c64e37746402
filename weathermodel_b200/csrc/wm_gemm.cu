// wm_gemm.cu -- tcgen05 / TMEM / TMA GEMMs for the encoder's dense layers (sm_100a only).
//
// Replaces the cuBLAS SGEMMs the reference reaches through nn.Linear / F.multi_head_attention_forward
// (reference: src/pretraining/models/weatherbert.py:34,45-56; torch/nn/modules/transformer.py:944-982)
//
//   gemm_tn   : C[M,N] = epilogue( A[M,K] . B[N,K]^T )   A,B bf16 row-major (K contiguous), fp32 acc.
//               Persistent, warp-specialised: warp0 = TMA producer, warp1 = MMA issuer (whole warp, elected lane),
//               8 / 12 / 16 epilogue warps (TMEM -> regs -> fused bias/ReLU/dropout/gate/residual -> global, shared-memory
//               staging or TMA-store boxes). TMEM holds two accumulator stages so the epilogue of tile i overlaps the
//               MMAs of i+1. Used for forward linears and for dgrad (with a transposed bf16 weight copy as B).
//   gemm_tn2  : the same on CTA pairs (cta_group::2, 256 x BN tiles, each CTA stages half of B).
//   gemm_wgrad: P[s][Nout,Kout] = A[Mtok,Nout]^T . B[Mtok,Kout] over token slice s (split-K),
//               both operands MN-major straight from the token-major activations (no transposes), one or two
//               128-row A^T tiles per CTA against the same B tile, bias gradient from an all-ones B chunk,
//               fp32 partials reduced deterministically by wgrad_reduce.
#include "wm_kernels.h"

#include <type_traits>

namespace wm {

// ------------------------------------------------------------------------------------------------
// host: TMA descriptor encode through the driver entry point (no -lcuda link dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2D bf16 row-major tensor [rows, cols] with row pitch ld (elements); box = {box_cols, box_rows};
// 128B swizzle (box_cols must be 64 bf16 = 128 B). Out-of-bounds elements read as zero.
static int make_tmap_bf16_swz(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                              uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swizzle) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return WM_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || ((ld * 2) & 15u)) return WM_ERR_ALIGN;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? WM_OK : WM_ERR_DRIVER;
}
int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                   uint32_t box_cols, uint32_t box_rows) {
  return make_tmap_bf16_swz(map, base, rows, cols, ld, box_cols, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

// 3D view {cols, S rows, B batches} of a token-major bf16 activation [B*S, ld] for the attention kernels:
// box = {8 columns (one 16-byte core-matrix row), box_rows, 1}, no swizzle -- every box lands as box_rows
// consecutive 16-byte rows, i.e. a column of canonical 8x16B core matrices. Rows >= S read as zero, so a tile
// never sees the next sequence of the batch.
int make_tmap_bf16_rows3d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t S, uint64_t B, uint64_t ld,
                          uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return WM_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || ((ld * 2) & 15u)) return WM_ERR_ALIGN;
  cuuint64_t gdim[3] = {cols, S, B};
  cuuint64_t gstr[2] = {ld * 2, S * ld * 2};
  cuuint32_t box[3] = {8, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? WM_OK : WM_ERR_DRIVER;
}

// 4D view {8 columns, S rows, cols/8 chunks, B batches} of the same activation: one box {8, box_rows, box_chunks, 1}
// lands as box_chunks consecutive core-matrix columns ([chunk][row][16 B]) -- a whole operand tile per TMA instruction
// instead of one per chunk. Coordinates: (0, row, first column / 8, batch).
int make_tmap_bf16_chunked4d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t S, uint64_t B, uint64_t ld,
                             uint32_t box_rows, uint32_t box_chunks) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return WM_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || ((ld * 2) & 15u) || (cols & 7u)) return WM_ERR_ALIGN;
  cuuint64_t gdim[4] = {8, S, cols / 8, B};
  cuuint64_t gstr[3] = {ld * 2, 16, S * ld * 2};
  cuuint32_t box[4] = {8, box_rows, box_chunks, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? WM_OK : WM_ERR_DRIVER;
}

// ------------------------------------------------------------------------------------------------
// gemm_tn
// ------------------------------------------------------------------------------------------------
constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kMaxStages = 8;
// warp0 TMA, warp1 MMA, then kEW epilogue warps: kEW / 4 per TMEM lane quarter, each owning BN / (kEW / 4) columns.
// With K = 576 the epilogue, not the MMAs (4,608 cycles per 128 x 256 tile), sets the pace. Round 1 measured ~1,070
// cycles per 16-column chunk at ~130 issued instructions with the run-time-flag body (hence 16 warps, 4 per
// scheduler); the straight-line bodies below (epi_fast16: 62-165 instructions per chunk) made 8 and 16 warps
// equivalent on most sites. 8 remain the only choice for tiles whose width is not a multiple of 64.
constexpr int gemm_threads(int ew) { return 64 + 32 * ew; }
constexpr int kWgradThreads = 192;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccStride = 256;
constexpr int kBiasResident = 2560;

struct GemmSmemTail {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  uint64_t res_full[16];           // per epilogue warp: the residual box of the current tile has landed (TMA-store epilogue)
  // the WHOLE bias vector (pre-scaled, zero-padded to whole tiles), staged once per CTA, when it fits (every layer of
  // the model does: N <= 4 x 576); otherwise per epilogue warp (kEW x 1024 / kEW floats) the slice of the current tile,
  // re-staged per tile -- a global-load latency per warp and tile with nothing to hide behind (ncu source view:
  // long-scoreboard on that store was 26 % of the per-tile stall samples of the K = 576 epilogues)
  alignas(16) float bias[kBiasResident];
};

// Epilogue of one 16-column group pair for one accumulator row. Auxiliary operands (residual / gate rows)
// arrive already loaded in registers: they are prefetched one chunk ahead so their L2 latency overlaps the
// TMEM load and the math of the previous chunk (the first version of this epilogue stalled ~25k cycles per
// tile on dependent global loads: profiles/r01_gemm_v1_stalls.txt).
struct EpiAux {
  uint4 a[2];  // the prefetched operand of a 16-column chunk: the gate rows if a gate is given, else the residual rows
  uint32_t bits;  // sign bits of the chunk (ep.gate_bits): bit j = element n0 + j of this row passes the gate
};
// bit of element j (0..15) of a chunk inside its uint16: even elements in the low byte, odd ones in the high byte
// (the order in which the packed bf16x2 output words yield them)
__host__ __device__ constexpr int sign_bit_pos(int j) { return (j & 1) * 8 + (j >> 1); }
WM_DEVICE size_t sign_bits_index(int row, int n0, int N) {
  return (static_cast<size_t>(row >> 5) * static_cast<size_t>((N + 15) >> 4) + static_cast<size_t>(n0 >> 4)) * 32 + (row & 31);
}

WM_DEVICE void epi_load_aux(EpiAux& x, const GemmEpilogue& ep, int row, int n0, int M, int N, bool wide) {
  const __nv_bfloat16* base = ep.gate ? ep.gate : ep.residual;
  const int ldx = ep.gate ? ep.ld_gate : ep.ld_res;
  x.a[0] = make_uint4(0u, 0u, 0u, 0u);
  x.a[1] = make_uint4(0u, 0u, 0u, 0u);
  x.bits = 0u;
  // (the buffer covers whole 32-row blocks of the M rows, not whole 128-row tiles)
  if (ep.gate_bits && n0 < N && row < ((M + 31) & ~31)) x.bits = __ldg(ep.gate_bits + sign_bits_index(row, n0, N));
  if (!base || row >= M) return;
  const __nv_bfloat16* p = base + static_cast<size_t>(row) * ldx + n0;
  if (wide && n0 + 16 <= N) {
    ldg256(p, x.a[0], x.a[1]);
  } else {
    if (n0 < N) x.a[0] = __ldg(reinterpret_cast<const uint4*>(p));
    if (n0 + 8 < N) x.a[1] = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  }
}

#ifdef WM_DIAG
// Diagnostic build only (tools/gemm_diag.py; never part of libwm_b200.so): bit 0 skips the output stores, bit 1 the
// arithmetic as well, bit 2 the TMEM loads, bit 3 stores the tile transposed-free into a compact per-CTA scratch.
int g_gemm_diag = 0;  // copied into GemmEpilogue::diag by the launcher
#endif

// ---- straight-line epilogues for the flag sets of the training schedule -------------------------------------
// The K = 576 GEMMs are bound by the ISSUE rate of their epilogue warps (ncu source view of the generic code below:
// ~17 issued instructions per output element, stalls "not selected" / "math pipe throttle"): every option is a
// run-time test there, and the compiler guards each with predicated register copies. For outputs in bf16 without a
// bf16 gate operand the flag set becomes a template argument and the arithmetic runs on register PAIRS: one FFMA2 /
// FADD2 / FMUL2 per two elements. Without a residual, ReLU and both kinds of mask act on the packed bf16x2 word AFTER
// the conversion (max(., 0), a positive scale and an all-or-nothing mask commute with the rounding); with one, the
// dropout mask is applied per fp32 element so that the sum is rounded once.
constexpr int kEpiBias = 1, kEpiRelu = 2, kEpiDrop = 4, kEpiGateBits = 8, kEpiSignOut = 16, kEpiResidual = 32;
WM_DEVICE int epi_fast_flags(const GemmEpilogue& ep, int out_bytes) {
  if (out_bytes != 2 || ep.gate) return -1;
#ifdef WM_DIAG
  if (ep.diag) return -1;
#endif
  const int f = (ep.bias ? kEpiBias : 0) | (ep.relu ? kEpiRelu : 0) | (ep.drop_thresh ? kEpiDrop : 0) |
                (ep.gate_bits ? kEpiGateBits : 0) | (ep.sign_bits_out ? kEpiSignOut : 0) | (ep.residual ? kEpiResidual : 0);
  switch (f) {  // qkv / plain dgrad / eval linear1 / training linear1 with and without dropout / linear2 dgrad /
                // out-proj and linear2 with and without dropout / the dgrads that add the gradient of the residual branch
    case 0: case kEpiBias: case kEpiBias | kEpiRelu: case kEpiBias | kEpiRelu | kEpiSignOut:
    case kEpiBias | kEpiRelu | kEpiDrop | kEpiSignOut: case kEpiGateBits:
    case kEpiBias | kEpiDrop | kEpiResidual: case kEpiBias | kEpiResidual: case kEpiResidual:
      return f;
  }
  return -1;
}
// With dropout the bias slice in shared memory is PRE-SCALED by 1 / (1 - p): (acc + b) * s = fma(acc, s, b * s).
WM_DEVICE float epi_bias_prescale(const GemmEpilogue& ep, int fast) { return (fast >= 0 && (fast & kEpiDrop)) ? ep.drop_scale : 1.0f; }
WM_DEVICE void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int kF>
WM_DEVICE void epi_fast16(const uint32_t (&v)[16], const EpiAux& aux, const float4 (&bias4)[4], const GemmEpilogue& ep,
                          const DropKeys& dk, int row, int n0, int M, int N, bool wide, uint32_t sdst, int swz_chunk,
                          bool res_in_box) {
  // sdst: 32-bit shared address of this lane's row of the TMA box (swz_chunk >= 0) or of its 32 bytes of the staging
  // tile (swz_chunk < 0); 0 = direct global stores. Same contract as epi_process16 otherwise.
  if (n0 >= N) return;  // warp-uniform
  const int m32 = (M + 31) & ~31;
  if (!sdst && row >= M && (!(kF & kEpiSignOut) || row >= m32)) return;
  uint32_t o[8];
  const float sc = (kF & kEpiDrop) ? ep.drop_scale : ep.gate_scale;
  const uint64_t sc2 = f2_pack(sc, sc);
  const uint32_t add2 = drop_add2(ep.drop_thresh);
  const uint32_t x0 = (static_cast<uint32_t>(row) * static_cast<uint32_t>((N + 15) >> 4) + static_cast<uint32_t>(n0 >> 4)) * 4u;
  const uint32_t r7 = static_cast<uint32_t>(threadIdx.x) & 7u;  // row inside a TMA box = lane
  uint32_t rw[8];
  if constexpr ((kF & kEpiResidual) != 0) {
    uint4 r0 = aux.a[0], r1 = aux.a[1];
    if (res_in_box) {  // the residual tile was TMA-loaded into this warp's output box: same swizzled positions as the output
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w)
                   : "r"(sdst + ((static_cast<uint32_t>(swz_chunk) ^ r7) << 4)) : "memory");
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w)
                   : "r"(sdst + ((static_cast<uint32_t>(swz_chunk + 1) ^ r7) << 4)) : "memory");
    }
    rw[0] = r0.x; rw[1] = r0.y; rw[2] = r0.z; rw[3] = r0.w;
    rw[4] = r1.x; rw[5] = r1.y; rw[6] = r1.z; rw[7] = r1.w;
  }
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    uint64_t p0 = f2_pack_u(v[4 * j4], v[4 * j4 + 1]), p1 = f2_pack_u(v[4 * j4 + 2], v[4 * j4 + 3]);
    if constexpr ((kF & kEpiBias) != 0) {
      const float4 bv = bias4[j4];
      if constexpr ((kF & kEpiDrop) != 0) {
        p0 = f2_fma(p0, sc2, f2_pack(bv.x, bv.y));
        p1 = f2_fma(p1, sc2, f2_pack(bv.z, bv.w));
      } else {
        p0 = f2_add(p0, f2_pack(bv.x, bv.y));
        p1 = f2_add(p1, f2_pack(bv.z, bv.w));
      }
    } else if constexpr ((kF & (kEpiDrop | kEpiGateBits)) != 0) {
      p0 = f2_mul(p0, sc2);
      p1 = f2_mul(p1, sc2);
    }
    float a0, a1, a2, a3;
    f2_unpack(p0, a0, a1);
    f2_unpack(p1, a2, a3);
    if constexpr ((kF & kEpiResidual) != 0) {  // the sum with the residual is formed in fp32: masks per element, then FADD2
      if constexpr ((kF & kEpiDrop) != 0) {
        const DropWords fl = drop_flags4(x0 + j4, dk, add2);
        a0 = __uint_as_float(__float_as_uint(a0) & drop_mask32<0>(fl));
        a1 = __uint_as_float(__float_as_uint(a1) & drop_mask32<1>(fl));
        a2 = __uint_as_float(__float_as_uint(a2) & drop_mask32<2>(fl));
        a3 = __uint_as_float(__float_as_uint(a3) & drop_mask32<3>(fl));
      }
      p0 = f2_add(f2_pack(a0, a1), f2_pack(bf16_lo(rw[2 * j4]), bf16_hi(rw[2 * j4])));
      p1 = f2_add(f2_pack(a2, a3), f2_pack(bf16_lo(rw[2 * j4 + 1]), bf16_hi(rw[2 * j4 + 1])));
      f2_unpack(p0, a0, a1);
      f2_unpack(p1, a2, a3);
    }
    o[2 * j4] = pack_bf16x2(a0, a1);
    o[2 * j4 + 1] = pack_bf16x2(a2, a3);
  }
  if constexpr ((kF & kEpiRelu) != 0) {
    const __nv_bfloat162 z = __floats2bfloat162_rn(0.0f, 0.0f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&o[j]), z);
      o[j] = *reinterpret_cast<const uint32_t*>(&r);
    }
  }
  if constexpr ((kF & kEpiDrop) != 0 && (kF & kEpiResidual) == 0) {
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const DropWords fl = drop_flags4(x0 + w, dk, add2);
      o[2 * w] &= drop_pair_mask(fl.a);
      o[2 * w + 1] &= drop_pair_mask(fl.b);
    }
  }
  if constexpr ((kF & kEpiGateBits) != 0) {  // bit j of the chunk's uint16 = even element of pair j, bit 8 + j = its odd one
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t m;  // the two bits go to the sign positions of bytes 0 and 1; PRMT replicates each over one half
      asm("prmt.b32 %0, %1, %2, 0x9988;" : "=r"(m) : "r"(aux.bits << (7 - j)), "r"(0u));
      o[j] &= m;
    }
  }
  const bool second = n0 + 8 < N;  // N % 8 == 0 (host-checked)
  if constexpr ((kF & kEpiSignOut) != 0) {  // (outputs are >= 0 here: "positive" is "non-zero halfword")
    uint32_t acc = 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += __vminu2(o[i], 0x00010001u) << i;  // even elements at bit i, odd at 16 + i
    uint32_t m = (acc & 0xFFu) | ((acc >> 8) & 0xFF00u);
    if (!second) m &= 0x0F0Fu;
    if (row < m32) ep.sign_bits_out[sign_bits_index(row, n0, N)] = static_cast<uint16_t>(row < M ? m : 0u);
    if (!sdst && row >= M) return;
  }
  const uint4 o0 = make_uint4(o[0], o[1], o[2], o[3]), o1 = make_uint4(o[4], o[5], o[6], o[7]);
  if (sdst) {
    if (swz_chunk >= 0) {
      sts128(sdst + ((static_cast<uint32_t>(swz_chunk) ^ r7) << 4), o0);
      sts128(sdst + ((static_cast<uint32_t>(swz_chunk + 1) ^ r7) << 4), o1);
    } else {
      sts128(sdst, o0);
      sts128(sdst + 16, o1);
    }
    return;
  }
  __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(ep.out) + static_cast<size_t>(row) * ep.ld_out + n0;
  if (wide && second) {
    stg256(op, o0, o1);
  } else {
    *reinterpret_cast<uint4*>(op) = o0;
    if (second) *reinterpret_cast<uint4*>(op + 8) = o1;
  }
}

template <typename OutT>
WM_DEVICE void epi_process16(const uint32_t (&v)[16], const EpiAux& aux, const float* sbias, const GemmEpilogue& ep,
                             const DropKeys& dk, int row, int n0, int M, int N, bool wide, uint8_t* sdst, int swz_chunk = -1,
                             bool res_in_box = false) {
  // sdst (bf16 outputs only): this lane's 32 bytes of the chunk inside the warp's shared-memory staging tile; the
  // tile leaves through coalesced stores once all chunks are in (gemm_epilogue_tile). nullptr: direct stores.
  // swz_chunk >= 0: sdst is the lane's 128-byte ROW of a 32 x 64 SWIZZLE_128B box (TMA-store epilogue) and the two
  // 16-byte pieces go to chunks (swz_chunk ^ (row & 7)) and ((swz_chunk + 1) ^ (row & 7)) of it -- the layout
  // cp.async.bulk.tensor expects, and conflict-free for lane = row (8 consecutive rows cover all 32 banks once).
  if (n0 >= N) return;  // warp-uniform
#ifdef WM_DIAG
  const int diag = ep.diag;
  if (diag & 2) {
    if (v[0] == 0x7fc12345u) reinterpret_cast<uint32_t*>(ep.out)[0] = v[3];  // keep the loads alive
    return;
  }
#endif
  const int m32 = (M + 31) & ~31;
  if (!sdst && row >= M && (!ep.sign_bits_out || row >= m32)) return;
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
  if (ep.bias) {  // four 16-byte broadcast loads (the scalar form cost ~2 L1 wavefronts per column: 38 % of the
                  // kernel's L1 data-pipe traffic in profiles/r01_gemm_lin1_full.txt)
    const float4* b4 = reinterpret_cast<const float4*>(sbias);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      const float4 bv = b4[j4];
      f[4 * j4] += bv.x;
      f[4 * j4 + 1] += bv.y;
      f[4 * j4 + 2] += bv.z;
      f[4 * j4 + 3] += bv.w;
    }
  }
  if (ep.relu) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
  }
  if (ep.drop_thresh) {
    const uint32_t add2 = drop_add2(ep.drop_thresh);
    const uint32_t x0 = (static_cast<uint32_t>(row) * static_cast<uint32_t>((N + 15) >> 4) + static_cast<uint32_t>(n0 >> 4)) * 4u;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const DropWords fl = drop_flags4(x0 + w, dk, add2);
      f[4 * w] = __uint_as_float(__float_as_uint(f[4 * w] * ep.drop_scale) & drop_mask32<0>(fl));
      f[4 * w + 1] = __uint_as_float(__float_as_uint(f[4 * w + 1] * ep.drop_scale) & drop_mask32<1>(fl));
      f[4 * w + 2] = __uint_as_float(__float_as_uint(f[4 * w + 2] * ep.drop_scale) & drop_mask32<2>(fl));
      f[4 * w + 3] = __uint_as_float(__float_as_uint(f[4 * w + 3] * ep.drop_scale) & drop_mask32<3>(fl));
    }
  }
  if (ep.gate_bits) {  // dgrad through dropout(relu(.)): pass where the producing GEMM recorded a positive output
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = ((aux.bits >> sign_bit_pos(j)) & 1u) ? f[j] * ep.gate_scale : 0.0f;
  } else if (ep.gate) {  // the same from the saved bf16 activation itself
    const uint32_t aw[8] = {aux.a[0].x, aux.a[0].y, aux.a[0].z, aux.a[0].w, aux.a[1].x, aux.a[1].y, aux.a[1].z, aux.a[1].w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[2 * j] = bf16_lo(aw[j]) > 0.0f ? f[2 * j] * ep.gate_scale : 0.0f;
      f[2 * j + 1] = bf16_hi(aw[j]) > 0.0f ? f[2 * j + 1] * ep.gate_scale : 0.0f;
    }
  }
  if (ep.residual) {
    uint4 r0 = aux.a[0], r1 = aux.a[1];
    if (res_in_box) {  // the residual tile was TMA-loaded into this warp's output box: same swizzled positions as the output
      const uint32_t r7 = static_cast<uint32_t>(threadIdx.x) & 7u;
      r0 = *reinterpret_cast<const uint4*>(sdst + ((static_cast<uint32_t>(swz_chunk) ^ r7) << 4));
      r1 = *reinterpret_cast<const uint4*>(sdst + ((static_cast<uint32_t>(swz_chunk + 1) ^ r7) << 4));
    } else if (ep.gate) {  // both operands given (not on the training schedule): the residual is loaded in place
      const uint4* rp = reinterpret_cast<const uint4*>(ep.residual + static_cast<size_t>(row) * ep.ld_res + n0);
      r0 = row < M ? __ldg(rp) : make_uint4(0u, 0u, 0u, 0u);
      r1 = (row < M && n0 + 8 < N) ? __ldg(rp + 1) : make_uint4(0u, 0u, 0u, 0u);
    }
    const uint32_t aw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[2 * j] += bf16_lo(aw[j]);
      f[2 * j + 1] += bf16_hi(aw[j]);
    }
  }
  const bool second = n0 + 8 < N;  // N % 8 == 0 (host-checked)
  if constexpr (sizeof(OutT) != 2) {
    if (ep.sign_bits_out) {
      uint32_t m = 0u;
#pragma unroll
      for (int j = 0; j < 16; ++j) m |= (f[j] > 0.0f && row < M && (j < 8 || second)) ? (1u << sign_bit_pos(j)) : 0u;
      ep.sign_bits_out[sign_bits_index(row, n0, N)] = static_cast<uint16_t>(m);
      if (row >= M) return;
    }
  }
  if constexpr (sizeof(OutT) == 2) {
    uint4 o0, o1;
    o0.x = pack_bf16x2(f[0], f[1]);   o0.y = pack_bf16x2(f[2], f[3]);
    o0.z = pack_bf16x2(f[4], f[5]);   o0.w = pack_bf16x2(f[6], f[7]);
    o1.x = pack_bf16x2(f[8], f[9]);   o1.y = pack_bf16x2(f[10], f[11]);
    o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
    if (ep.sign_bits_out) {
      // 32 lanes x 2 bytes, contiguous: the rows of a whole 32-row block (rows >= M write zeros). The outputs are
      // >= 0 here (post ReLU), so "positive" is "non-zero halfword": one packed min per word, then shift-adds.
      const uint32_t w8[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
      uint32_t acc = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += __vminu2(w8[i], 0x00010001u) << i;  // even elements at bit i, odd at 16 + i
      uint32_t m = (acc & 0xFFu) | ((acc >> 8) & 0xFF00u);
      if (!second) m &= 0x0F0Fu;
      if (row < m32) ep.sign_bits_out[sign_bits_index(row, n0, N)] = static_cast<uint16_t>(row < M ? m : 0u);
      if (!sdst && row >= M) return;
    }
    if (sdst) {
      if (swz_chunk >= 0) {
        const uint32_t r7 = static_cast<uint32_t>(threadIdx.x) & 7u;  // row inside the box = lane
        *reinterpret_cast<uint4*>(sdst + ((static_cast<uint32_t>(swz_chunk) ^ r7) << 4)) = o0;
        *reinterpret_cast<uint4*>(sdst + ((static_cast<uint32_t>(swz_chunk + 1) ^ r7) << 4)) = o1;
      } else {
        *reinterpret_cast<uint4*>(sdst) = o0;
        *reinterpret_cast<uint4*>(sdst + 16) = o1;
      }
      return;
    }
#ifdef WM_DIAG
    if (diag & 1) {
      if (o0.x == 0x7fc17fc1u && o1.w == 0x12345678u) reinterpret_cast<uint32_t*>(ep.out)[0] = o0.y;
      return;
    }
    if (diag & 8) {  // same bytes, but every row of the tile lands in one contiguous 64 KB block per CTA
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + (static_cast<size_t>(blockIdx.x) * 128 + (row & 127)) * 256 + (n0 & 255);
      stg256(o, o0, o1);
      return;
    }
#endif
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + static_cast<size_t>(row) * ep.ld_out + n0;
    if (wide && second) {
      stg256(o, o0, o1);
    } else {
      *reinterpret_cast<uint4*>(o) = o0;
      if (second) *reinterpret_cast<uint4*>(o + 8) = o1;
    }
  } else {
    float* o = reinterpret_cast<float*>(ep.out) + static_cast<size_t>(row) * ep.ld_out + n0;
    const int lim = second ? 16 : 8;
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      if (j < lim) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
  }
}

// The chunk loop of one epilogue warp and tile: TMEM loads run one 16-column chunk ahead of the arithmetic.
// kF >= 0: one of the straight-line epilogues (epi_fast16<kF>); kF < 0: the generic one.
template <typename OutT, int kAuxDepth, int kF>
WM_DEVICE void epi_chunks(const GemmEpilogue& ep, const DropKeys& dk, EpiAux (&aux)[kAuxDepth], const float* sbias, uint32_t tbase,
                          int row, int n_base, int cols_per, int M, int N, bool wide, int lane, uint8_t* stage, uint32_t stage_s,
                          uint32_t pitch, bool box, bool res_box, bool box_live) {
  uint32_t va[16], vb[16];
  tmem_ld16(tbase, va);
  for (int cb = 0; cb < cols_per; cb += 16 * kAuxDepth) {
#pragma unroll
    for (int d = 0; d < kAuxDepth; ++d) {
      const int c0 = cb + d * 16;
      if (c0 < cols_per) {
#ifdef WM_DIAG
        if (ep.diag & 4) continue;
#endif
        if constexpr (kF >= 0) {
          // (the bias values of the chunk are fetched BEFORE the wait: tcgen05.wait::ld is a compiler barrier for memory
          // operations, and behind it the first FADD2 / FFMA2 stalled on this load)
          float4 b4[4];
          if constexpr ((kF & kEpiBias) != 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) b4[j] = reinterpret_cast<const float4*>(sbias + c0)[j];
          }
          tmem_ld_wait();
          const uint32_t dst = box ? stage_s : (stage ? stage_s + static_cast<uint32_t>(c0) * 2u : 0u);
          if (d & 1) {
            if (c0 + 16 < cols_per) tmem_ld16(tbase + c0 + 16, va);
            epi_fast16<kF>(vb, aux[d], b4, ep, dk, row, n_base + c0, M, N, wide, dst, box ? (c0 >> 3) : -1, box_live);
          } else {
            if (c0 + 16 < cols_per) tmem_ld16(tbase + c0 + 16, vb);
            epi_fast16<kF>(va, aux[d], b4, ep, dk, row, n_base + c0, M, N, wide, dst, box ? (c0 >> 3) : -1, box_live);
          }
          if constexpr ((kF & (kEpiGateBits | kEpiResidual)) != 0) {
            if (!res_box && c0 + 16 * kAuxDepth < cols_per) epi_load_aux(aux[d], ep, row, n_base + c0 + 16 * kAuxDepth, M, N, wide);
          }
        } else {
          tmem_ld_wait();
          if (d & 1) {
            if (c0 + 16 < cols_per) tmem_ld16(tbase + c0 + 16, va);
            if (box) epi_process16<OutT>(vb, aux[d], sbias + c0, ep, dk, row, n_base + c0, M, N, wide, stage + lane * 128, c0 >> 3, box_live);
            else epi_process16<OutT>(vb, aux[d], sbias + c0, ep, dk, row, n_base + c0, M, N, wide,
                                     stage ? stage + lane * pitch + c0 * 2 : nullptr);
          } else {
            if (c0 + 16 < cols_per) tmem_ld16(tbase + c0 + 16, vb);
            if (box) epi_process16<OutT>(va, aux[d], sbias + c0, ep, dk, row, n_base + c0, M, N, wide, stage + lane * 128, c0 >> 3, box_live);
            else epi_process16<OutT>(va, aux[d], sbias + c0, ep, dk, row, n_base + c0, M, N, wide,
                                     stage ? stage + lane * pitch + c0 * 2 : nullptr);
          }
          if (!res_box && c0 + 16 * kAuxDepth < cols_per) epi_load_aux(aux[d], ep, row, n_base + c0 + 16 * kAuxDepth, M, N, wide);
        }
      }
    }
  }
}

// One epilogue warp's share of one accumulator tile: its 32 TMEM lanes (rows) x cols_per columns from n_base on.
// Everything that does not depend on the accumulator (bias slice, first residual / gate chunks) is issued before
// waiting for the MMAs; TMEM loads run one 16-column chunk ahead of the arithmetic.
template <typename OutT, int kAuxDepth, int kF>
WM_DEVICE void gemm_epilogue_tile(const GemmEpilogue& ep, float* sbias, uint64_t* acc_full, uint32_t aph,
                                  uint32_t wait_code, uint32_t tbase, int row, int n_base, int cols_per, int M, int N,
                                  bool wide, int lane, uint8_t* stage, uint64_t* acc_empty, uint32_t acc_empty_cluster,
                                  const CUtensorMap* tmC = nullptr, const CUtensorMap* tmR = nullptr,
                                  uint64_t* res_full = nullptr, uint32_t* res_count = nullptr, bool bias_resident = false,
                                  const DropKeys* live_keys = nullptr) {
  // kF: epi_fast_flags() of this launch (-1: the generic epilogue). bias_resident: sbias already points at this
  // tile's columns of the whole (pre-scaled) bias vector, staged once per CTA.
  // tmR != nullptr (with tmC): the residual operand arrives by TMA, too -- one box load per warp and tile into the output
  // box itself (read-modify-write in place) instead of 32 row-strided 32-byte loads per lane and chunk.
  // tmC != nullptr (staged == 2, cols_per == 64): TMA-store epilogue. `stage` is this warp's 1024-byte aligned 4 KB
  // box (32 rows x 64 bf16, SWIZZLE_128B); the lanes write their packed rows straight from the tcgen05.ld registers,
  // one elected lane hands the box to cp.async.bulk.tensor (rows >= M / columns >= N are clipped by the tensor map)
  // and the warp moves on: no read-back of the staging tile, no per-lane address arithmetic, no st.global.
  // acc_empty / acc_empty_cluster: the barrier that hands the accumulator stage back to the MMA warp (a local
  // barrier, or the leader CTA's as a shared::cluster address when acc_empty is nullptr). It is released as soon as
  // the last TMEM load has landed -- with staged stores that is before the tile is written out.
  // stage: this warp's 32 x (cols_per * 2 + 16)-byte staging tile or nullptr. Thread-per-row stores (32 bytes per
  // lane, rows a leading dimension apart) cost the LSU / L1 one line per lane: ~8 B/clk/SM, and with K = 576 they,
  // not the MMAs, set the pace (tools/gemm_diag.py: linear1 0.505 ms with, 0.359 ms without the stores). Staged, a
  // warp-wide 16-byte store covers whole row segments: 4 - 5 lines per instruction instead of 32.
  const uint32_t pitch = static_cast<uint32_t>(cols_per) * 2u + 16u;  // odd number of 16-byte units: conflict-free
  // (+ the per-replay words of a captured step: device globals, read once per warp by the caller where it can)
  const DropKeys dk = live_keys ? *live_keys : (ep.drop_thresh ? drop_keys_live(ep.dkeys) : ep.dkeys);
  if (ep.bias && !bias_resident) {
    const float bs = epi_bias_prescale(ep, kF);
    __syncwarp();
    for (int j = lane; j < cols_per; j += 32) sbias[j] = (n_base + j < N) ? __ldg(ep.bias + n_base + j) * bs : 0.0f;
    __syncwarp();
  }
  // ring of kAuxDepth prefetched 16-column chunks per thread (4 with 8 epilogue warps, 2 with 16): ~32 KB of
  // residual / gate rows in flight per SM
  static_assert(kAuxDepth == 2 || kAuxDepth == 4, "the TMEM double buffer alternates on the chunk parity");
  EpiAux aux[kAuxDepth];
  const bool res_box = tmR != nullptr;
#pragma unroll
  for (int d = 0; d < kAuxDepth; ++d) {
    if (res_box) { aux[d].a[0] = aux[d].a[1] = make_uint4(0u, 0u, 0u, 0u); aux[d].bits = 0u; }
    else if (d * 16 < cols_per) epi_load_aux(aux[d], ep, row, n_base + d * 16, M, N, wide);
  }
  if (tmC) {  // the previous tile's box must have been read by the TMA engine before it is overwritten
    if (lane == 0) {
      bulk_wait_group_read<0>();
      if (res_box && n_base < N && row < M) {  // (lane 0's row is the first row of the box; skipped boxes are never waited for)
        mbar_arrive_expect_tx(res_full, 4096u);
        tma_load_2d(stage, tmR, res_full, n_base, row);
      }
    }
    __syncwarp();
  }
  mbar_wait(acc_full, aph, wait_code);
  tc_fence_after();
  const bool box_live = res_box && n_base < N && (row - lane) < M;
  if (box_live) {  // (warp-uniform) one completion of this warp's barrier per box actually loaded
    mbar_wait(res_full, *res_count & 1u, wait_code + 100);
    ++*res_count;
  }
  const uint32_t stage_s = stage ? smem_u32(stage) + static_cast<uint32_t>(lane) * (tmC ? 128u : pitch) : 0u;
epi_chunks<OutT, kAuxDepth, (sizeof(OutT) == 2 ? kF : -1)>(ep, dk, aux, sbias, tbase, row, n_base, cols_per, M, N, wide, lane, stage, stage_s, pitch,
                                                              tmC != nullptr, res_box, box_live);
#ifdef WM_DIAG
  tmem_ld_wait();
#endif
  tc_fence_before();
  __syncwarp();
  if (lane == 0) {
    if (acc_empty) mbar_arrive(acc_empty);
    else mbar_arrive_cluster(acc_empty_cluster);
  }
  if constexpr (sizeof(OutT) == 2) {
    if (tmC) {
      fence_proxy_async_smem();  // the lanes' generic-proxy writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0 && n_base < N && row < M) {  // (row == first row of the box for lane 0)
        tma_store_2d(tmC, stage, n_base, row);
        bulk_commit_group();
      }
    } else if (stage) {
      __syncwarp();
      // 16-byte units of the tile in row-major order, 32 per instruction: unit i = lane + 32 k is piece i % u of
      // row i / u (u = units per row); consecutive lanes write consecutive global addresses within a row
      const int u = cols_per >> 3;
      const int q32 = 32 / u, m32u = 32 - q32 * u;
      int r = lane / u, piece = lane - r * u;
      const int row0 = row - lane;
      __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(ep.out);
      for (int k = 0; k < u; ++k) {
        const int col = n_base + piece * 8;
        const uint4 val = *reinterpret_cast<const uint4*>(stage + r * pitch + piece * 16);
#ifdef WM_DIAG
        if ((ep.diag & 1) && !(val.x == 0x7fc17fc1u && val.w == 0x12345678u)) continue;
#endif
        if (row0 + r < M && col < N)
          *reinterpret_cast<uint4*>(outp + static_cast<size_t>(row0 + r) * ep.ld_out + col) = val;
        r += q32;
        piece += m32u;
        if (piece >= u) { piece -= u; ++r; }
      }
      __syncwarp();  // the next tile's chunks overwrite the staging tile
    }
  }
}

// The persistent loop of one epilogue warp over its CTA's (or CTA pair's) tiles, for one epilogue flag set kF. The
// dispatch over kF happens ONCE per kernel, outside this loop: with it inside, ptxas hoisted the loop invariants of all
// seven bodies out of the tile loop together and spilled ~10 KB.
struct EpiRole {
  int t0, tstride, total_tiles, n_tiles, BN, M, N;
  int row_tile, row_off;      // rows per tile index (128, or 256 for a CTA pair) and this CTA's offset inside it
  uint32_t tmem_base, epi_stage_warp, wait_code;
  uint8_t* epi_stage;
  int staged, res_tma;
  bool bias_resident, pair;
};
template <typename OutT, int kEW, int kF>
WM_DEVICE void gemm_epilogue_role(const GemmEpilogue& ep, const EpiRole& r, GemmSmemTail* tail, const CUtensorMap* tmC,
                                  const CUtensorMap* tmR, int warp, int lane) {
  const int q = warp & 3;                 // TMEM lane quarter this warp may read
  const int ew = warp - 2;                // 0..kEW-1
  const int half = ew >> 2;               // which slice of the tile's columns this warp owns
  const int cols_per = r.BN / (kEW / 4);  // a multiple of 16 (host-checked)
  float* sbias = tail->bias + ew * (kEW > 8 ? 64 : 128);  // (not resident:) 16-byte aligned slices, >= the warp's column count
  // 32-byte accesses need 32-byte aligned rows: every leading dimension a multiple of 16 bf16 elements
  const bool wide = ((ep.ld_out | (ep.residual ? ep.ld_res : 0) | (ep.gate ? ep.ld_gate : 0)) & 15) == 0 &&
                    ((reinterpret_cast<uintptr_t>(ep.out) | reinterpret_cast<uintptr_t>(ep.residual) |
                      reinterpret_cast<uintptr_t>(ep.gate)) & 31) == 0;
  uint32_t acc_empty_leader[2] = {0u, 0u};
  if (r.pair) {
    acc_empty_leader[0] = mapa_u32(&tail->acc_empty[0], 0);
    acc_empty_leader[1] = mapa_u32(&tail->acc_empty[1], 0);
  }
  int it = 0;
  uint32_t res_count = 0u;
  const DropKeys dk = ep.drop_thresh ? drop_keys_live(ep.dkeys) : ep.dkeys;
  for (int t = r.t0; t < r.total_tiles; t += r.tstride, ++it) {
    const int m_blk = t / r.n_tiles, n_blk = t % r.n_tiles;
    const int as = it & 1;
    const uint32_t aph = (it >> 1) & 1u;
    const uint32_t tbase = r.tmem_base + as * kAccStride + half * cols_per + (static_cast<uint32_t>(q * 32) << 16);
    gemm_epilogue_tile<OutT, (kEW == 16 ? 2 : 4), kF>(
        ep, r.bias_resident ? tail->bias + n_blk * r.BN + half * cols_per : sbias, &tail->acc_full[as], aph, r.wait_code, tbase,
        m_blk * r.row_tile + r.row_off + q * 32 + lane, n_blk * r.BN + half * cols_per, cols_per, r.M, r.N, wide, lane,
        r.staged ? r.epi_stage + static_cast<uint32_t>(ew) * r.epi_stage_warp : nullptr, r.pair ? nullptr : &tail->acc_empty[as],
        acc_empty_leader[as], r.staged == 2 ? tmC : nullptr, (r.staged == 2 && r.res_tma) ? tmR : nullptr, &tail->res_full[ew],
        &res_count, r.bias_resident, &dk);
  }
  if (r.staged == 2 && lane == 0) bulk_wait_group_read<0>();  // the last box must be read before the CTA's smem goes away
}
template <typename OutT, int kEW>
WM_DEVICE void gemm_epilogue_dispatch(int fast, const GemmEpilogue& ep, const EpiRole& r, GemmSmemTail* tail,
                                      const CUtensorMap* tmC, const CUtensorMap* tmR, int warp, int lane) {
#define WM_EPI_ROLE(F) gemm_epilogue_role<OutT, kEW, F>(ep, r, tail, tmC, tmR, warp, lane)
  if constexpr (sizeof(OutT) == 2) {
    if (fast == (kEpiBias | kEpiRelu | kEpiDrop | kEpiSignOut)) WM_EPI_ROLE(kEpiBias | kEpiRelu | kEpiDrop | kEpiSignOut);
    else if (fast == kEpiGateBits) WM_EPI_ROLE(kEpiGateBits);
    else if (fast == kEpiBias) WM_EPI_ROLE(kEpiBias);
    else if (fast == 0) WM_EPI_ROLE(0);
    else if (fast == (kEpiBias | kEpiRelu | kEpiSignOut)) WM_EPI_ROLE(kEpiBias | kEpiRelu | kEpiSignOut);
    else if (fast == (kEpiBias | kEpiRelu)) WM_EPI_ROLE(kEpiBias | kEpiRelu);
    else if (fast == (kEpiBias | kEpiDrop | kEpiResidual)) WM_EPI_ROLE(kEpiBias | kEpiDrop | kEpiResidual);
    else if (fast == kEpiResidual) WM_EPI_ROLE(kEpiResidual);
    else if (fast == (kEpiBias | kEpiResidual)) WM_EPI_ROLE(kEpiBias | kEpiResidual);
    else WM_EPI_ROLE(-1);
  } else {
    WM_EPI_ROLE(-1);
  }
#undef WM_EPI_ROLE
}

template <typename OutT, int kEW>
__global__ void __launch_bounds__(gemm_threads(kEW), 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, int M, int N, int K, int BN,
               int stages, int staged, int res_tma, GemmEpilogue ep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024B-align the tile ring (SWIZZLE_128B atoms)
  uint8_t* smem = smem_align_up(smem_raw, 1024);
  const uint32_t a_bytes = kBM * kBK * 2;
  const uint32_t b_bytes = static_cast<uint32_t>(BN) * kBK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  // [tile ring][epilogue staging: kEW x 32 rows x (BN / (kEW / 4) * 2 + 16) bytes if staged == 1, kEW boxes of 4 KB
  // (32 rows x 64 columns, SWIZZLE_128B, for the TMA-store epilogue) if staged == 2][tail]
  uint8_t* epi_stage = smem + static_cast<size_t>(stages) * stage_bytes;
  const uint32_t epi_stage_warp = staged == 2 ? 4096u : 32u * (static_cast<uint32_t>(BN / (kEW / 4)) * 2u + 16u);
  GemmSmemTail* tail = reinterpret_cast<GemmSmemTail*>(epi_stage + (staged ? kEW * epi_stage_warp : 0u));

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + kBM - 1) / kBM;
  const int n_tiles = (N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int num_kb = (K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (staged == 2) tma_prefetch_desc(&tmC);
    if (res_tma) tma_prefetch_desc(&tmR);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tail->acc_full[s], 1);
      mbar_init(&tail->acc_empty[s], kEW);  // one arrive per epilogue warp
    }
    for (int s = 0; s < 16; ++s) mbar_init(&tail->res_full[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(&tail->tmem_base);
  const int fast = epi_fast_flags(ep, static_cast<int>(sizeof(OutT)));
  const bool bias_resident = ep.bias && n_tiles * BN <= kBiasResident;
  if (bias_resident) {
    const float bs = epi_bias_prescale(ep, fast);
    for (int j = threadIdx.x; j < n_tiles * BN; j += blockDim.x) tail->bias[j] = j < N ? __ldg(ep.bias + j) * bs : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m_blk = t / n_tiles, n_blk = t % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&tail->empty[s], ph ^ 1u, 11);
          uint8_t* sa = smem + static_cast<size_t>(s) * stage_bytes;
          mbar_arrive_expect_tx(&tail->full[s], stage_bytes);
          tma_load_2d(sa, &tmA, &tail->full[s], kb * kBK, m_blk * kBM);
          tma_load_2d(sa + a_bytes, &tmB, &tail->full[s], kb * kBK, n_blk * BN);
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issue: the WHOLE warp runs this loop convergently and umma_*_warp elect the issuing lane. One thread inside
    // `if (lane == 0)` makes the compiler wrap every tcgen05 instruction in a per-active-lane loop: ~100 cycles per
    // MMA (tools/mmabench.cu), i.e. 400 of the 512 cycles a k-block's four N = 256 MMAs take, plus the barrier wait.
    const uint32_t idesc = umma_idesc_bf16(kBM, static_cast<uint32_t>(BN), 0, 0);
    const uint64_t desc_a0 = umma_smem_desc(smem_u32(smem), 16, 1024, UMMA_SWZ_128B);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1u;
      mbar_wait(&tail->acc_empty[as], aph ^ 1u, 12);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * kAccStride;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&tail->full[s], ph, 13);
        tc_fence_after();
        // descriptors differ from the stage-0 ones only in the start-address field: one add per MMA
        const uint64_t da_s = umma_desc_advance(desc_a0, static_cast<uint32_t>(s) * stage_bytes);
        const uint64_t db_s = umma_desc_advance(da_s, a_bytes);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)
          umma_ss_warp(d_tmem, umma_desc_advance(da_s, k * 32), umma_desc_advance(db_s, k * 32), idesc,
                       (kb | k) != 0 ? 1u : 0u);
        umma_commit_warp(&tail->empty[s]);
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
      umma_commit_warp(&tail->acc_full[as]);
    }
  } else {
    const EpiRole r{static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x), total_tiles, n_tiles, BN, M, N, kBM, 0,
                    tmem_base, epi_stage_warp, 14u, epi_stage, staged, res_tma, bias_resident, false};
    gemm_epilogue_dispatch<OutT, kEW>(fast, ep, r, tail, &tmC, &tmR, warp, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// gemm_tn2: the same GEMM on CTA PAIRS (cta_group::2). One 256 x BN tile per pair: each CTA stages its own 128
// rows of A and HALF of the B tile, so the B operand crosses L2 -> smem once per 256 rows instead of once per 128
// (the single-CTA kernel is L2-feed bound at ~1.0-1.1 PFLOP/s). The leader CTA issues M = 256 MMAs and commits to
// the barriers of both CTAs; each CTA's 8 epilogue warps drain their own 128 TMEM lanes.
// ------------------------------------------------------------------------------------------------
template <typename OutT, int kEW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(gemm_threads(kEW), 1)
gemm_tn2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, int M, int N, int K, int BN,
                int stages, int staged, int res_tma, GemmEpilogue ep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align_up(smem_raw, 1024);
  const int BNH = BN >> 1;  // B rows staged by each CTA
  const uint32_t a_bytes = kBM * kBK * 2;
  const uint32_t b_bytes = static_cast<uint32_t>(BNH) * kBK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  // [tile ring][epilogue staging (see gemm_tn_kernel)][tail]
  uint8_t* epi_stage = smem + static_cast<size_t>(stages) * stage_bytes;
  const uint32_t epi_stage_warp = staged == 2 ? 4096u : 32u * (static_cast<uint32_t>(BN / (kEW / 4)) * 2u + 16u);
  GemmSmemTail* tail = reinterpret_cast<GemmSmemTail*>(epi_stage + (staged ? kEW * epi_stage_warp : 0u));

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_tiles = (M + 2 * kBM - 1) / (2 * kBM);
  const int n_tiles = (N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int num_kb = (K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tail->acc_full[s], 1);
      mbar_init(&tail->acc_empty[s], 2 * kEW);  // the epilogue warps of both CTAs (the leader's copy is used)
    }
    if (res_tma) tma_prefetch_desc(&tmR);
    for (int s = 0; s < 16; ++s) mbar_init(&tail->res_full[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta<kTmemCols>(&tail->tmem_base);
  const int fast = epi_fast_flags(ep, static_cast<int>(sizeof(OutT)));
  const bool bias_resident = ep.bias && n_tiles * BN <= kBiasResident;
  if (bias_resident) {
    const float bs = epi_bias_prescale(ep, fast);
    for (int j = threadIdx.x; j < n_tiles * BN; j += blockDim.x) tail->bias[j] = j < N ? __ldg(ep.bias + j) * bs : 0.0f;
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = pair; t < total_tiles; t += npairs) {
        const int m_blk = t / n_tiles, n_blk = t % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&tail->empty[s], ph ^ 1u, 61);
          uint8_t* sa = smem + static_cast<size_t>(s) * stage_bytes;
          if (leader) mbar_arrive_expect_tx(&tail->full[s], 2 * stage_bytes);  // bytes of BOTH CTAs land here
          tma_load_2d_2cta(sa, &tmA, &tail->full[s], kb * kBK, m_blk * 2 * kBM + static_cast<int>(rank) * kBM);
          tma_load_2d_2cta(sa + a_bytes, &tmB, &tail->full[s], kb * kBK, n_blk * BN + static_cast<int>(rank) * BNH);
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {  // whole warp, convergent issue (see gemm_tn_kernel)
      const uint32_t idesc = umma_idesc_bf16(2 * kBM, static_cast<uint32_t>(BN), 0, 0);
      const uint64_t desc_a0 = umma_smem_desc(smem_u32(smem), 16, 1024, UMMA_SWZ_128B);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = pair; t < total_tiles; t += npairs, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1u;
        mbar_wait(&tail->acc_empty[as], aph ^ 1u, 62);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&tail->full[s], ph, 63);
          tc_fence_after();
          const uint64_t da_s = umma_desc_advance(desc_a0, static_cast<uint32_t>(s) * stage_bytes);
          const uint64_t db_s = umma_desc_advance(da_s, a_bytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_ss_2cta_warp(d_tmem, umma_desc_advance(da_s, k * 32), umma_desc_advance(db_s, k * 32), idesc,
                              (kb | k) != 0 ? 1u : 0u);
          umma_commit_2cta_warp(&tail->empty[s]);
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
        umma_commit_2cta_warp(&tail->acc_full[as]);
      }
    }
  } else {
    const EpiRole r{pair, npairs, total_tiles, n_tiles, BN, M, N, 2 * kBM, static_cast<int>(rank) * kBM,
                    tmem_base, epi_stage_warp, 64u, epi_stage, staged, res_tma, bias_resident, true};
    gemm_epilogue_dispatch<OutT, kEW>(fast, ep, r, tail, &tmC, &tmR, warp, lane);
  }
  __syncwarp();  // reconverge the single-lane role warps: barrier.cluster.*.aligned needs whole warps
  tc_fence_before();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while its partner can still touch its smem / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<kTmemCols>(tmem_base);
  }
}

// tile width: multiples of 32 up to 256; minimise tiles * (bn + ~32 columns of per-tile overhead), then the
// padded column count, then prefer the wider tile
static int pick_bn(int N) {
  int best = 32;
  long best_cost = -1, best_pad = 0;
  for (int bn = 32; bn <= 256; bn += 32) {
    const long tiles = (N + bn - 1) / bn;
    const long cost = tiles * (bn + 32), pad = tiles * bn;
    if (best_cost < 0 || cost < best_cost || (cost == best_cost && pad < best_pad) ||
        (cost == best_cost && pad == best_pad && bn > best)) {
      best = bn;
      best_cost = cost;
      best_pad = pad;
    }
  }
  return best;
}

// Kernel variant per call site. All variants are bit-identical (tests/test_gpu_kernels.py), so the choice is purely a
// matter of speed: CTA-pair tiles halve the shared-memory operand traffic per SM and win when the main loop dominates
// (K >= 1024: +6-19 %), 16 epilogue warps win when the epilogue does (K = 576: +1-9 %); tools/gemm_sites.py.
// Order of precedence: forced option (wm_set_option) > tuned table entry (wm_gemm_set_variant) > heuristic.
int g_gemm_two_cta = -1;   // "gemm_two_cta": -1 auto, 0 single-CTA tiles, 1 CTA-pair tiles (M >= 1024)
int g_gemm_epi_warps = 0;  // "gemm_epi_warps": 0 auto, 8 or 16 (16 needs a tile width that is a multiple of 64)
int g_gemm_staged = -1;    // "gemm_staged": -1 auto, 0 thread-per-row stores, 1 stores staged through shared memory, 2 TMA-store boxes

struct GemmVariant { int M, N, K; uint32_t sig; int two_cta, epi_warps, staged; };
constexpr int kMaxGemmVariants = 128;
static GemmVariant g_gemm_variants[kMaxGemmVariants];
static int g_num_gemm_variants = 0;

#ifdef WM_DIAG
int gemm_set_diag(int v) { g_gemm_diag = v; return WM_OK; }
#endif
uint32_t gemm_signature(const GemmEpilogue& ep, int out_fp32) {
  return (ep.bias ? 1u : 0u) | (ep.relu ? 2u : 0u) | (ep.drop_thresh ? 4u : 0u) | (ep.gate ? 8u : 0u) |
         (ep.gate_bits ? 16u : 0u) | (ep.residual ? 32u : 0u) | (ep.sign_bits_out ? 64u : 0u) | (out_fp32 ? 128u : 0u);
}
int gemm_set_variant(int M, int N, int K, uint32_t sig, int two_cta, int epi_warps, int staged) {
  if ((two_cta != 0 && two_cta != 1) || (epi_warps != 8 && epi_warps != 16) || staged < 0 || staged > 2) return WM_ERR_ARG;
  for (int i = 0; i < g_num_gemm_variants; ++i) {
    GemmVariant& v = g_gemm_variants[i];
    if (v.M == M && v.N == N && v.K == K && v.sig == sig) {
      v.two_cta = two_cta;
      v.epi_warps = epi_warps;
      v.staged = staged;
      return WM_OK;
    }
  }
  if (g_num_gemm_variants == kMaxGemmVariants) g_num_gemm_variants = 0;  // start over rather than fail
  g_gemm_variants[g_num_gemm_variants++] = GemmVariant{M, N, K, sig, two_cta, epi_warps, staged};
  return WM_OK;
}
static void gemm_pick_variant(int M, int N, int K, uint32_t sig, bool residual, int* two_cta, int* epi_warps, int* staged) {
  // untuned sites of wide models: CTA pairs, 8 epilogue warps, direct stores -- except the short-K products that add a
  // residual, where the TMA-store epilogue also TMA-loads the residual into the box (out-proj forward: 0.149 against
  // 0.214 ms). Measured on the D = 576 shapes (profiles/r02_gemm_sites_variants_v3.txt); narrower models gain nothing
  // from either and keep the plain variant.
  const bool wide_model = N >= 512 && K >= 512;
  const bool box = wide_model && residual && K < 1024;
  int two = wide_model ? 1 : 0, ew = box ? 16 : 8, stg = box ? 2 : 0;
  for (int i = 0; i < g_num_gemm_variants; ++i) {
    const GemmVariant& v = g_gemm_variants[i];
    if (v.M == M && v.N == N && v.K == K && v.sig == sig) {
      two = v.two_cta;
      ew = v.epi_warps;
      stg = v.staged;
      break;
    }
  }
  if (g_gemm_two_cta >= 0) two = g_gemm_two_cta;
  if (g_gemm_epi_warps > 0) ew = g_gemm_epi_warps;
  if (g_gemm_staged >= 0) stg = g_gemm_staged;
  *two_cta = two;
  *epi_warps = ew;
  *staged = stg;
}
static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

static int launch_gemm_tn_impl(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int b_rows,
                               const GemmEpilogue& ep_in, int out_fp32, int bn_override, cudaStream_t stream) {
  GemmEpilogue ep = ep_in;
#ifdef WM_DIAG
  ep.diag = g_gemm_diag;
#endif
  if (ep.drop_thresh) {
    if (static_cast<uint64_t>(M) * static_cast<uint64_t>((N + 15) / 16) * 4ull > 0xFFFFFFFFull) return WM_ERR_SHAPE;  // 32-bit mask counters
    ep.dkeys = drop_keys(ep.seed, ep.stream);
  }
  if (M <= 0 || N <= 0 || K <= 0 || b_rows <= 0 || b_rows > N) return WM_ERR_SHAPE;
  if ((N & 7) || (K & 7) || (lda & 7) || (ldb & 7) || (ep.ld_out & 7)) return WM_ERR_ALIGN;
  const int BN = bn_override > 0 ? bn_override : pick_bn(N);
  if (BN & 31 || BN > 256 || BN < 32) return WM_ERR_SHAPE;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16(&tmA, A, M, K, lda, kBK, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmB, B, b_rows, K, ldb, kBK, BN);
  if (rc) return rc;
  int want_two, want_ew, want_staged;
  gemm_pick_variant(M, N, K, gemm_signature(ep, out_fp32), ep.residual != nullptr && !ep.gate && !ep.gate_bits, &want_two, &want_ew,
                    &want_staged);
  const int fixed_smem = 2048 + static_cast<int>(sizeof(GemmSmemTail));
  const bool aligned_out = !out_fp32 && (ep.ld_out & 7) == 0 && (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0;
  // staged == 2: TMA-store epilogue. Every epilogue warp owns one 32 x 64 box, so the warp count follows from the
  // tile width: BN = 256 -> 16 warps, 192 -> 12, 128 -> 8.
  const bool tma_out = want_staged == 2 && aligned_out && (BN & 63) == 0 && BN >= 128;
  const int ew = tma_out ? BN / 16 : ((want_ew >= 16 && (BN & 63) == 0) ? 16 : 8);
  const int threads = gemm_threads(ew);
  // staged stores: 16-byte aligned rows and at least three pipeline stages left next to the staging tiles
  const int staging = tma_out ? ew * 4096 : ew * 32 * (BN / (ew / 4) * 2 + 16);
  const bool can_stage = want_staged && aligned_out;
  CUtensorMap tmC = tmA;  // (unused unless tma_out; a valid map keeps the __grid_constant__ copy well-defined)
  CUtensorMap tmR = tmA;
  // the residual rides the same boxes when it is a plain 16-byte aligned bf16 matrix (no gate operand in the way)
  const bool res_tma = tma_out && ep.residual && !ep.gate && !ep.gate_bits && (ep.ld_res & 7) == 0 && (reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0;
  if (tma_out) {
    rc = make_tmap_bf16(&tmC, ep.out, M, N, ep.ld_out, 64, 32);
    if (rc) return rc;
    if (res_tma) {
      rc = make_tmap_bf16(&tmR, ep.residual, M, N, ep.ld_res, 64, 32);
      if (rc) return rc;
    }
  }
  if (want_two && !out_fp32 && M >= 1024 && (BN & 31) == 0 && bn_override >= 0) {
    // CTA-pair path: 256 x BN tiles, each CTA stages 128 rows of A and BN/2 rows of B per k-block
    rc = make_tmap_bf16(&tmB, B, b_rows, K, ldb, kBK, BN / 2);
    if (rc) return rc;
    const int stage2 = (kBM + BN / 2) * kBK * 2;
    const int staged2 = (can_stage && (227 * 1024 - fixed_smem - staging) / stage2 >= 3) ? (tma_out ? 2 : 1) : 0;
    int st2 = (227 * 1024 - fixed_smem - (staged2 ? staging : 0)) / stage2;
    if (st2 > kMaxStages) st2 = kMaxStages;
    const int smem2 = st2 * stage2 + fixed_smem - 1024 + (staged2 ? staging : 0);
    const int tiles2 = ((M + 2 * kBM - 1) / (2 * kBM)) * ((N + BN - 1) / BN);
    const int pairs = min(tiles2, num_sms() / 2);
    if (!tma_out || staged2 == 2) {
      auto kern2 = ew == 16 ? gemm_tn2_kernel<__nv_bfloat16, 16> : ew == 12 ? gemm_tn2_kernel<__nv_bfloat16, 12> : gemm_tn2_kernel<__nv_bfloat16, 8>;
      if (cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2) != cudaSuccess) return WM_ERR_CUDA;
      kern2<<<2 * pairs, threads, smem2, stream>>>(tmA, tmB, tmC, tmR, M, N, K, BN, st2, staged2, (staged2 == 2 && res_tma) ? 1 : 0, ep);
      WM_COUNT_LAUNCH();
      return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
    }
    rc = make_tmap_bf16(&tmB, B, b_rows, K, ldb, kBK, BN);  // (no room for the boxes: fall through to single-CTA tiles)
    if (rc) return rc;
  }
  const int stage_bytes = (kBM + BN) * kBK * 2;
  const int staged1 = (can_stage && (227 * 1024 - fixed_smem - staging) / stage_bytes >= 3) ? (tma_out ? 2 : 1) : 0;
  if (tma_out && staged1 != 2) return WM_ERR_SHAPE;  // cannot happen for BN <= 256 (3 x 48 KB + 64 KB + tail < 227 KB)
  int stages = (227 * 1024 - fixed_smem - (staged1 ? staging : 0)) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  const int smem = stages * stage_bytes + fixed_smem - 1024 + (staged1 ? staging : 0);
  const int m_tiles = (M + kBM - 1) / kBM, n_tiles = (N + BN - 1) / BN;
  const int grid = min(m_tiles * n_tiles, num_sms());
  auto kern = out_fp32 ? (ew == 16 ? gemm_tn_kernel<float, 16> : gemm_tn_kernel<float, 8>)
                       : (ew == 16 ? gemm_tn_kernel<__nv_bfloat16, 16>
                                   : ew == 12 ? gemm_tn_kernel<__nv_bfloat16, 12> : gemm_tn_kernel<__nv_bfloat16, 8>);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return WM_ERR_CUDA;
  kern<<<grid, threads, smem, stream>>>(tmA, tmB, tmC, tmR, M, N, K, BN, stages, staged1, (staged1 == 2 && res_tma) ? 1 : 0, ep);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

size_t gemm_sign_bits_bytes(int M, int N) {
  return static_cast<size_t>((M + 31) / 32) * static_cast<size_t>((N + 15) / 16) * 32 * sizeof(uint16_t);
}

int launch_gemm_tn(const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                   const GemmEpilogue& ep, int out_fp32, int bn_override, cudaStream_t stream) {
  return launch_gemm_tn_impl(A, lda, B, ldb, M, N, K, N, ep, out_fp32, bn_override, stream);
}
// B has only b_rows (< N) real rows; the rest of the N tile is TMA zero-fill (padded output heads)
int launch_gemm_tn_rows(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int b_rows,
                        const GemmEpilogue& ep, int out_fp32, cudaStream_t stream) {
  return launch_gemm_tn_impl(A, lda, B, ldb, M, N, K, b_rows, ep, out_fp32, 0, stream);
}

// ------------------------------------------------------------------------------------------------
// gemm_wgrad: P[split][Nout, Kout] = sum_{t in slice} A[t, Nout]^T B[t, Kout]   (both MN-major)
// ------------------------------------------------------------------------------------------------
constexpr int kWgStages = 4;
struct WgSmemTail {
  uint64_t full[kWgStages];
  uint64_t empty[kWgStages];
  uint64_t acc_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kWgradThreads, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  int Mtok, int Nout, int Kout, int BN, int tok_per_split, float* __restrict__ partial,
                  float* __restrict__ bias_partial, int mh, int stages) {
  // mh = 2: the CTA multiplies TWO 128-row tiles of A^T (rows 256 x and 256 x + 128 of Nout) with the same B tile, each into
  // its own accumulator (TMEM columns [0, 256) and [256, 512)). What holds this kernel back is how much operand data an
  // SM can take in per clock: one 128 x 192 tile needs 40 KB per 384 MMA cycles (104 B/clk; ncu: ~69 B/clk achieved, tensor
  // pipe 71 % active, the MMA warp waiting for operands a third of its time) -- two tiles share the B half of that: 56 KB
  // per 768 cycles = 73 B/clk. (A CTA-pair version saved less and an L2 prefetch made it worse: profiles/r02_wgrad_bn_ab.txt.)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align_up(smem_raw, 1024);
  const uint32_t a_half = kBM * kBK * 2;                              // 2 boxes of 64 tok x 64 n
  const uint32_t a_bytes = static_cast<uint32_t>(mh) * a_half;
  const uint32_t b_bytes = static_cast<uint32_t>(BN) * kBK * 2;       // BN/64 boxes
  // Fused bias gradient (column sums of A over the tokens): an all-ones 64 x 64 chunk sits right behind the B
  // tile of every stage, so the k_blk == 0 CTAs simply run their MMAs 16 columns wider (N = BN + 16) and find
  // sum_t A[t, n] in accumulator column BN -- no separate pass over the activation gradient. With BN = 256 (the
  // instruction's widest N) the ones chunk gets its own N = 16 MMA per k-step into columns [256, 272) instead.
  const bool fuse_bias = bias_partial != nullptr;
  const bool wide_bias = fuse_bias && BN == 256;  // (host: mh == 1 then -- two 272-column accumulators do not fit)
  const uint32_t ones_bytes = fuse_bias ? 8192u : 0u;
  const uint32_t stage_bytes = a_bytes + b_bytes + ones_bytes;
  WgSmemTail* tail = reinterpret_cast<WgSmemTail*>(smem + static_cast<size_t>(stages) * stage_bytes);

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const int n_blk0 = blockIdx.x * mh;  // first 128-row tile over Nout (UMMA M side) of this CTA
  const int nh = min(mh, (Nout + kBM - 1) / kBM - n_blk0);  // tiles it really has (the last CTA of an odd count: one)
  const int k_blk = blockIdx.y;  // tile over Kout (UMMA N side)
  const int split = blockIdx.z;
  const int tok0 = split * tok_per_split;
  const int tok1 = min(Mtok, tok0 + tok_per_split);
  const int num_kb = (tok1 - tok0 + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    mbar_init(&tail->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tail->tmem_base);
  if (fuse_bias && k_blk == 0) {
    const uint4 ones = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);  // bf16 1.0 x 8
    for (int i = threadIdx.x; i < stages * 512; i += blockDim.x)
      *reinterpret_cast<uint4*>(smem + static_cast<size_t>(i >> 9) * stage_bytes + a_bytes + b_bytes + (i & 511) * 16) = ones;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&tail->empty[s], ph ^ 1u, 21);
        uint8_t* sa = smem + static_cast<size_t>(s) * stage_bytes;
        mbar_arrive_expect_tx(&tail->full[s], static_cast<uint32_t>(nh) * a_half + b_bytes);
        const int t = tok0 + kb * kBK;
        // NOTE: token rows past tok1 but < Mtok would belong to the next split; tok_per_split is a
        // multiple of 64 so a box never straddles a split boundary; rows >= Mtok are zero-filled.
        for (int c = 0; c < nh * (kBM / 64); ++c)
          tma_load_2d(sa + c * 8192, &tmA, &tail->full[s], n_blk0 * kBM + c * 64, t);
        for (int c = 0; c < BN / 64; ++c)
          tma_load_2d(sa + a_bytes + c * 8192, &tmB, &tail->full[s], k_blk * BN + c * 64, t);
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // MMA issue, whole warp convergent (see gemm_tn_kernel)
    const bool ones = fuse_bias && k_blk == 0;
    const uint32_t idesc = umma_idesc_bf16(kBM, static_cast<uint32_t>(BN + (ones && !wide_bias ? 16 : 0)), 1, 1);
    const uint32_t idesc_ones = umma_idesc_bf16(kBM, 16, 1, 1);
    // MN-major SW128: 64-wide MN chunks LBO = 8192 B apart, 8-token groups SBO = 1024 B apart
    const uint64_t desc_a0 = umma_smem_desc(smem_u32(smem), 8192, 1024, UMMA_SWZ_128B);
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&tail->full[s], ph, 23);
      tc_fence_after();
      const uint64_t da_s = umma_desc_advance(desc_a0, static_cast<uint32_t>(s) * stage_bytes);
      const uint64_t db_s = umma_desc_advance(da_s, a_bytes);
#pragma unroll
      for (int k = 0; k < kBK / 16; ++k) {
        umma_ss_warp(tmem_base, umma_desc_advance(da_s, k * 2048), umma_desc_advance(db_s, k * 2048), idesc,
                     (kb | k) != 0 ? 1u : 0u);
        if (nh == 2)  // (warp-uniform) the second A^T tile against the same B tile
          umma_ss_warp(tmem_base + 256, umma_desc_advance(da_s, a_half + k * 2048), umma_desc_advance(db_s, k * 2048), idesc,
                       (kb | k) != 0 ? 1u : 0u);
        if (ones && wide_bias)  // (warp-uniform) the all-ones chunk as a second, 16-column product
          umma_ss_warp(tmem_base + 256, umma_desc_advance(da_s, k * 2048), umma_desc_advance(db_s, b_bytes + k * 2048), idesc_ones,
                       (kb | k) != 0 ? 1u : 0u);
      }
      umma_commit_warp(&tail->empty[s]);
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
    umma_commit_warp(&tail->acc_full);
  } else {
    const int q = warp & 3;
    float* out = partial + static_cast<size_t>(split) * Nout * Kout;
    if (num_kb > 0) {
      mbar_wait(&tail->acc_full, 0, 24);
      tc_fence_after();
    }
    // All MMAs are done, so the tile ring is free: the warp's 32 x BN fp32 block goes through it (row pitch
    // BN * 4 + 16 bytes: an odd number of 16-byte units, conflict-free for lane = row) and leaves as coalesced
    // 512-byte row segments. Thread-per-row float4 stores (32 lines per instruction) cost ~8 B/clk/SM -- 12k cycles
    // for a 128 x 192 tile, against ~100k cycles of MMAs per work item.
    const uint32_t pitch = static_cast<uint32_t>(BN) * 4u + 16u;
    uint8_t* stage = smem + static_cast<uint32_t>(q) * 32u * pitch;
    for (int h = 0; h < nh; ++h) {
      const int row0 = (n_blk0 + h) * kBM + q * 32;
      const uint32_t tcol = tmem_base + static_cast<uint32_t>(h) * 256u + (static_cast<uint32_t>(q * 32) << 16);
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        if (num_kb > 0) {
          tmem_ld32(tcol + c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<uint4*>(stage + lane * pitch + (c0 + g * 4) * 4) = make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
      }
      __syncwarp();
      {
        const int u = BN >> 2;  // 16-byte units per row
        const int q32 = 32 / u, m32u = 32 - q32 * u;
        int r = lane / u, piece = lane - r * u;
        for (int k = 0; k < u; ++k) {
          const int n = k_blk * BN + piece * 4;
          const uint4 val = *reinterpret_cast<const uint4*>(stage + r * pitch + piece * 16);
          if (row0 + r < Nout && n < Kout)  // Kout % 4 == 0 (host-checked)
            *reinterpret_cast<uint4*>(out + static_cast<size_t>(row0 + r) * Kout + n) = val;
          r += q32;
          piece += m32u;
          if (piece >= u) { piece -= u; ++r; }
        }
      }
      __syncwarp();  // the next half overwrites the staging block
      const int row = row0 + lane;
      if (fuse_bias && k_blk == 0) {
        uint32_t v[16];
        if (num_kb > 0) {
          tmem_ld16(tcol + BN, v);
          tmem_ld_wait();
        } else {
          v[0] = 0u;
        }
        if (row < Nout) bias_partial[static_cast<size_t>(split) * Nout + row] = __uint_as_float(v[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// out[r, c] = (accumulate ? out[r, c] : 0) + sum_s partial[s][r, c] for r < rows_valid, c < cols_valid
// (fixed summation order -> deterministic); out row pitch ld_out, partial tiles are [Nout, Kout].
// The fused bias gradient's partials are folded by extra blocks of the SAME launch (blockIdx.x >= main_blocks): one
// launch per weight gradient instead of two.
WM_DEVICE void wgrad_bias_fold(const float* __restrict__ partial, float* __restrict__ out, int Nout, int rows_valid, int splits,
                               int block) {
  const int r = block * blockDim.x + threadIdx.x;
  if (r >= rows_valid) return;
  float acc = 0.0f;
  for (int s = 0; s < splits; ++s) acc += partial[static_cast<size_t>(s) * Nout + r];
  out[r] = acc;
}
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int Nout, int Kout,
                                    int rows_valid, int cols_valid, int ld_out, int splits, int accumulate, int main_blocks,
                                    const float* __restrict__ bias_partial, float* __restrict__ dbias) {
  if (static_cast<int>(blockIdx.x) >= main_blocks) {
    wgrad_bias_fold(bias_partial, dbias, Nout, rows_valid, splits, blockIdx.x - main_blocks);
    return;
  }
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(rows_valid) * cols_valid) return;
  const int r = static_cast<int>(i / cols_valid), c = static_cast<int>(i - static_cast<int64_t>(r) * cols_valid);
  const size_t n = static_cast<size_t>(Nout) * Kout;
  const size_t src = static_cast<size_t>(r) * Kout + c;
  float acc = accumulate ? out[static_cast<size_t>(r) * ld_out + c] : 0.0f;
  for (int s = 0; s < splits; ++s) acc += __ldg(partial + s * n + src);
  out[static_cast<size_t>(r) * ld_out + c] = acc;
}

// the same for full-width outputs whose rows are 16-byte aligned: one float4 per thread and split
__global__ void wgrad_reduce4_kernel(const float4* __restrict__ partial, float4* __restrict__ out, int64_t n4_per_split,
                                     int64_t n4_valid, int splits, int accumulate, int main_blocks,
                                     const float* __restrict__ bias_partial, float* __restrict__ dbias, int Nout, int rows_valid) {
  if (static_cast<int>(blockIdx.x) >= main_blocks) {
    wgrad_bias_fold(bias_partial, dbias, Nout, rows_valid, splits, blockIdx.x - main_blocks);
    return;
  }
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4_valid) return;
  float4 acc = accumulate ? out[i] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  for (int s = 0; s < splits; ++s) {
    const float4 v = __ldg(partial + s * n4_per_split + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  out[i] = acc;
}

// Work items = output tiles x token splits, one per CTA, all the same size: pick the split count that fills whole
// waves of SMs (the first plan used ceil(2 * SMs / tiles) and lost up to a third of the machine to a nearly empty
// last wave: 300 items on 148 SMs = 3 waves at 68 %), with a mild preference for fewer partial tiles to reduce.
int g_wgrad_bn = 0;  // "wgrad_bn": 0 = the rule below, 64..256 = forced tile width (A/B runs)
int g_wgrad_mh = 0;  // "wgrad_mh": 0 = the rule below, 1 / 2 = A^T tiles per CTA forced
// A^T tiles (128 rows of Nout) per CTA: two where both accumulators fit in TMEM (tile width <= 240 with the bias column)
// and the problem is big enough to fill the machine with half as many CTAs per token split
static int wgrad_pick_mh(int Mtok, int Nout, int bn) {
  int mh = (bn <= 192 && Nout > kBM && Mtok >= 16384) ? 2 : 1;
  if (g_wgrad_mh == 1 || g_wgrad_mh == 2) mh = (bn <= 192 && Nout > kBM) ? g_wgrad_mh : 1;
  return mh;
}
int wgrad_plan(int Mtok, int Nout, int Kout, int* BN, int* splits, int* tok_per_split, int* mh_out = nullptr) {
  // tile width over Kout (64-column TMA boxes): 256 for wide outputs it divides, 192 where that divides Kout (the D = 576
  // shapes were tuned on it: 256-wide tiles there lose 15-25 % to the padded third tile), otherwise
  // the width that minimises tiles x (width + ~32 columns of per-tile overhead), the wider one on a tie
  int bn = 64;
  if (Kout % 256 == 0 && Kout >= 1024) {
    bn = 256;  // (linear2's 576 x 2304 gradient: 0.484 ms against 0.510-0.516 with 192, tools/wgrad_bn_ab.py)
  } else if (Kout % 192 == 0) {
    bn = 192;
  } else {
    long best = -1;
    for (int cand = 64; cand <= 256; cand += 64) {
      const long cost = static_cast<long>((Kout + cand - 1) / cand) * (cand + 32);
      if (best < 0 || cost <= best) {
        best = cost;
        bn = cand;
      }
    }
  }
  if (g_wgrad_bn >= 64 && g_wgrad_bn <= 256 && (g_wgrad_bn & 63) == 0) bn = g_wgrad_bn;
  const int mh = wgrad_pick_mh(Mtok, Nout, bn);
  if (mh_out) *mh_out = mh;
  const int tiles = (((Nout + kBM - 1) / kBM + mh - 1) / mh) * ((Kout + bn - 1) / bn);
  const int kb_total = (Mtok + kBK - 1) / kBK;
  const int sms = num_sms();
  int best_sp = 1;
  double best_score = -1.0;
  for (int sp = 1; sp <= 48 && sp <= kb_total; ++sp) {
    const int kb_per = (kb_total + sp - 1) / sp;
    if (sp > 1 && kb_per < 16) break;  // keep the per-item prologue / epilogue amortised
    const int sp_eff = (kb_total + kb_per - 1) / kb_per;
    const long items = static_cast<long>(tiles) * sp_eff;
    const long waves = (items + sms - 1) / sms;
    // time ~ waves * k-blocks per item; normalise by the ideal tiles * kb_total / sms
    const double t = static_cast<double>(waves) * kb_per;
    const double ideal = static_cast<double>(tiles) * kb_total / sms;
    const double score = ideal / t - 0.003 * sp_eff;
    if (score > best_score) {
      best_score = score;
      best_sp = sp_eff;
    }
  }
  const int kb_per = (kb_total + best_sp - 1) / best_sp;
  *BN = bn;
  *splits = (kb_total + kb_per - 1) / kb_per;
  *tok_per_split = kb_per * kBK;
  return tiles;
}

// true if launch_gemm_wgrad produces the bias gradient inside the GEMM (room for the all-ones chunk next to the tile)
bool wgrad_fuses_bias(int Mtok, int Nout, int Kout) {
  int bn, sp, tps;
  wgrad_plan(Mtok, Nout, Kout, &bn, &sp, &tps);
  (void)bn;
  return true;  // every tile width fuses it now (256-wide tiles through a second, 16-column MMA)
}

size_t wgrad_workspace_bytes(int Mtok, int Nout, int Kout) {
  int bn, sp, tps;
  wgrad_plan(Mtok, Nout, Kout, &bn, &sp, &tps);
  size_t bytes = static_cast<size_t>(sp) * Nout * Kout * sizeof(float) + static_cast<size_t>(sp) * Nout * sizeof(float);
  // fallback path of the fused bias gradient (BN == 256): column-sum workspace + one row of results
  const size_t cs = colsum_workspace_bytes(Mtok, Nout) + static_cast<size_t>(Nout) * sizeof(float);
  return bytes > cs ? bytes : cs;
}

static int launch_gemm_wgrad_impl(const void* A, int lda, const void* B, int ldb, int Mtok, int Nout, int Kout,
                                  float* dW, int rows_valid, int cols_valid, int ld_dw, int accumulate,
                                  float* workspace, float* dbias, cudaStream_t stream) {
  if (Mtok <= 0 || Nout <= 0 || Kout <= 0) return WM_ERR_SHAPE;
  if (rows_valid > Nout || cols_valid > Kout || ld_dw < cols_valid) return WM_ERR_SHAPE;
  if ((Nout & 7) || (Kout & 7) || (lda & 7) || (ldb & 7)) return WM_ERR_ALIGN;
  int BN, splits, tps, mh;
  wgrad_plan(Mtok, Nout, Kout, &BN, &splits, &tps, &mh);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16(&tmA, A, Mtok, Nout, lda, 64, kBK);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmB, B, Mtok, Kout, ldb, 64, kBK);
  if (rc) return rc;
  const bool fuse = dbias != nullptr;
  float* bias_partial = fuse ? workspace + static_cast<size_t>(splits) * Nout * Kout : nullptr;
  const int stage_bytes = (mh * kBM + BN) * kBK * 2 + (fuse ? 8192 : 0);
  int stages = (227 * 1024 - static_cast<int>(sizeof(WgSmemTail)) - 1024) / stage_bytes;
  if (stages > kWgStages) stages = kWgStages;
  if (stages < 2) return WM_ERR_SHAPE;
  const int smem = stages * stage_bytes + static_cast<int>(sizeof(WgSmemTail)) + 1024;
  if (cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return WM_ERR_CUDA;
  dim3 grid(((Nout + kBM - 1) / kBM + mh - 1) / mh, (Kout + BN - 1) / BN, splits);
  gemm_wgrad_kernel<<<grid, kWgradThreads, smem, stream>>>(tmA, tmB, Mtok, Nout, Kout, BN, tps, workspace, bias_partial, mh, stages);
  WM_COUNT_LAUNCH();
  if (cudaGetLastError() != cudaSuccess) return WM_ERR_CUDA;
  const int64_t n = static_cast<int64_t>(rows_valid) * cols_valid;
  const int threads = 256;
  const int bias_blocks = (dbias && fuse) ? (rows_valid + threads - 1) / threads : 0;  // folded by the same launch
  if (cols_valid == Kout && ld_dw == Kout && (Kout & 3) == 0 && (reinterpret_cast<uintptr_t>(dW) & 15u) == 0 &&
      (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0) {
    // contiguous [rows_valid, Kout] block: same element order and summation order, four columns per thread
    const int64_t n4 = n / 4;
    const int main_blocks = static_cast<int>((n4 + threads - 1) / threads);
    wgrad_reduce4_kernel<<<main_blocks + bias_blocks, threads, 0, stream>>>(
        reinterpret_cast<const float4*>(workspace), reinterpret_cast<float4*>(dW),
        static_cast<int64_t>(Nout) * Kout / 4, n4, splits, accumulate, main_blocks, bias_partial, dbias, Nout, rows_valid);
  } else {
    const int main_blocks = static_cast<int>((n + threads - 1) / threads);
    wgrad_reduce_kernel<<<main_blocks + bias_blocks, threads, 0, stream>>>(workspace, dW, Nout, Kout, rows_valid, cols_valid, ld_dw,
                                                                           splits, accumulate, main_blocks, bias_partial, dbias);
  }
  WM_COUNT_LAUNCH();
  if (cudaGetLastError() != cudaSuccess) return WM_ERR_CUDA;
  if (dbias && !fuse) {  // 256-wide tiles leave no room for the ones chunk: separate column-sum pass (stream-ordered reuse
                         // of the workspace: the partials above have been consumed by wgrad_reduce)
    float* tmp = workspace + colsum_workspace_bytes(Mtok, Nout) / sizeof(float);
    const int rc2 = launch_colsum(reinterpret_cast<const __nv_bfloat16*>(A), lda, Mtok, Nout, tmp, workspace, stream);
    if (rc2) return rc2;
    if (cudaMemcpyAsync(dbias, tmp, static_cast<size_t>(rows_valid) * sizeof(float), cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
      return WM_ERR_CUDA;
    return WM_OK;
  }
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}
int launch_gemm_wgrad(const void* A, int lda, const void* B, int ldb, int Mtok, int Nout, int Kout, float* dW,
                      int accumulate, float* workspace, float* dbias, cudaStream_t stream) {
  return launch_gemm_wgrad_impl(A, lda, B, ldb, Mtok, Nout, Kout, dW, Nout, Kout, Kout, accumulate, workspace, dbias,
                                stream);
}
// only the top-left [rows_valid, cols_valid] block of the product is written, with row pitch ld_dw
int launch_gemm_wgrad_ex(const void* A, int lda, const void* B, int ldb, int Mtok, int Nout, int Kout, float* dW,
                         int rows_valid, int cols_valid, int ld_dw, float* workspace, float* dbias,
                         cudaStream_t stream) {
  return launch_gemm_wgrad_impl(A, lda, B, ldb, Mtok, Nout, Kout, dW, rows_valid, cols_valid, ld_dw, 0, workspace,
                                dbias, stream);
}

// ------------------------------------------------------------------------------------------------
// umma_probe: one CTA, one 128 x N x K product with operands staged by hand into the UNSWIZZLED
// canonical core-matrix layouts the attention kernels use. Validates descriptor semantics
// (K-major / MN-major, LBO / SBO) on hardware independently of TMA.
//   A: [128, K] bf16 row-major in global, B: [N, K] bf16 row-major, D: [128, N] fp32.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                  float* __restrict__ D, int N, int K, int a_mn, int b_mn) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align_up(smem_raw, 1024);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int kg = K / 8;  // core matrices along K
  // K-major : elem(r,k) at (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2,  LBO=128, SBO=kg*128
  // MN-major: elem(r,k) at (r/8)*SBO + (k/8)*LBO + (k%8)*16 + (r%8)*2,  SBO=128, LBO=(rows/8)*128
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;
  const uint32_t a_lbo = a_mn ? (128 / 8) * 128 : 128, a_sbo = a_mn ? 128 : kg * 128;
  const uint32_t b_lbo = b_mn ? (N / 8) * 128 : 128, b_sbo = b_mn ? 128 : kg * 128;
  for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    const uint32_t off = (r / 8) * a_sbo + (k / 8) * a_lbo + (a_mn ? (k % 8) * 16 + (r % 8) * 2 : (r % 8) * 16 + (k % 8) * 2);
    *reinterpret_cast<__nv_bfloat16*>(sA + off) = A[i];
  }
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    const uint32_t off = (r / 8) * b_sbo + (k / 8) * b_lbo + (b_mn ? (k % 8) * 16 + (r % 8) * 2 : (r % 8) * 16 + (k % 8) * 2);
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = B[i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(N), a_mn, b_mn);
    for (int k = 0; k < K / 16; ++k) {
      // one UMMA covers 16 k = two core matrices along K
      const uint64_t da = umma_smem_desc(smem_u32(sA) + k * 2 * a_lbo, a_lbo, a_sbo, UMMA_SWZ_NONE);
      const uint64_t db = umma_smem_desc(smem_u32(sB) + k * 2 * b_lbo, b_lbo, b_sbo, UMMA_SWZ_NONE);
      umma_ss(tmem_base, da, db, idesc, k != 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0, 31);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_base + c0 + (static_cast<uint32_t>(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[static_cast<size_t>(row) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

int launch_umma_probe(const void* A, const void* B, float* D, int N, int K, int a_mn, int b_mn,
                      cudaStream_t stream) {
  if (N % 16 || N > 256 || K % 16 || K > 256) return WM_ERR_SHAPE;
  const int smem = (128 + N) * K * 2 + 1024;
  if (cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return WM_ERR_CUDA;
  umma_probe_kernel<<<1, 128, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(A),
                                              reinterpret_cast<const __nv_bfloat16*>(B), D, N, K, a_mn, b_mn);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

}  // namespace wm
