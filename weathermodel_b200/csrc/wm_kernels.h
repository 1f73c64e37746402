// wm_kernels.h -- internal (C++) launcher declarations shared by the .cu translation units.
// The public C-ABI lives in include/wm_b200.h; wm_api.cu maps one onto the other.
#pragma once

#include "wm_common.cuh"
#include <stddef.h>

#define WM_OK 0
#define WM_ERR_SHAPE 1
#define WM_ERR_ALIGN 2
#define WM_ERR_CUDA 3
#define WM_ERR_DRIVER 4
#define WM_ERR_ARG 5
#define WM_ERR_DEVICE 6

namespace wm {

// Fused GEMM epilogue: v = acc (+bias) -> [relu] -> [dropout] -> [gate: aux>0 ? v*s : 0] -> (+residual)
struct GemmEpilogue {
  const float* bias = nullptr;           // [N] fp32
  int relu = 0;
  uint32_t drop_thresh = 0;              // round(p * 65536); 0 = no dropout (the kernels compare 15 bits: p_eff = round(p * 32768) / 32768)
  float drop_scale = 1.0f;               // drop_keep_scale(drop_thresh)
  uint64_t seed = 0;                     // dropout key
  uint64_t stream = 0;                   // dropout stream: (step, layer, site)
  DropKeys dkeys = {0u, 0u, 0u};            // filled in by the launcher from (seed, stream)
  const __nv_bfloat16* gate = nullptr;   // [M, ld_gate] saved post-activation (dgrad through ReLU+dropout)
  int ld_gate = 0;
  float gate_scale = 1.0f;
  const __nv_bfloat16* residual = nullptr;  // [M, ld_res]
  int ld_res = 0;
  // Sign side channel for the ReLU-gate dgrad: the producing GEMM writes one bit per output element (value > 0
  // after ReLU / dropout), the dgrad GEMM reads the bits instead of the bf16 activation (1/16 of the bytes, one
  // coalesced 64-byte access per warp and 16-column chunk instead of 32 row-strided sectors). Layout: one uint16
  // per (row, 16-column chunk): [row / 32][chunk][row % 32], element j at bit (j & 1) * 8 + (j >> 1);
  // gemm_sign_bits_bytes(M, N) bytes. Only meaningful behind a ReLU (outputs >= 0).
  uint16_t* sign_bits_out = nullptr;
  const uint16_t* gate_bits = nullptr;   // replaces `gate` when given (gate_scale still applies)
  void* out = nullptr;                   // [M, ld_out] bf16 or fp32
  int ld_out = 0;
#ifdef WM_DIAG
  int diag = 0;                          // diagnostic builds only (tools/gemm_diag.py)
#endif
};

int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                   uint32_t box_cols, uint32_t box_rows);

size_t gemm_sign_bits_bytes(int M, int N);
uint32_t gemm_signature(const GemmEpilogue& ep, int out_fp32);
int gemm_set_variant(int M, int N, int K, uint32_t sig, int two_cta, int epi_warps, int staged);
int launch_gemm_tn(const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                   const GemmEpilogue& ep, int out_fp32, int bn_override, cudaStream_t stream);
int launch_gemm_tn_rows(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int b_rows,
                        const GemmEpilogue& ep, int out_fp32, cudaStream_t stream);
int launch_gemm_wgrad_ex(const void* A, int lda, const void* B, int ldb, int Mtok, int Nout, int Kout, float* dW,
                         int rows_valid, int cols_valid, int ld_dw, float* workspace, float* dbias,
                         cudaStream_t stream);
size_t wgrad_workspace_bytes(int Mtok, int Nout, int Kout);
bool wgrad_fuses_bias(int Mtok, int Nout, int Kout);
// dbias (optional): column sums of A over the tokens = bias gradient of the layer whose output gradient A is
int launch_gemm_wgrad(const void* A, int lda, const void* B, int ldb, int Mtok, int Nout, int Kout,
                      float* dW, int accumulate, float* workspace, float* dbias, cudaStream_t stream);
int launch_umma_probe(const void* A, const void* B, float* D, int N, int K, int a_mn, int b_mn,
                      cudaStream_t stream);

// ---- memory-bound kernels (wm_elementwise.cu) ------------------------------------------------
int launch_mask_bert(uint64_t seed, uint64_t philox_offset, int grid_x, float p, int64_t numel,
                     uint8_t* mask, float* rand_out, cudaStream_t stream);
int launch_mask_former(uint64_t seed, uint64_t philox_offset, int grid_x, int n_masked, int64_t n_samples,
                       int n_features, uint8_t* mask, cudaStream_t stream);
int launch_embed_fwd(const float* weather, const uint8_t* mask, int64_t mask_stride_b, int64_t mask_stride_s,
                     const float* year, const float* coords, const float* w_in, const float* b_in,
                     const float* pe, __nv_bfloat16* out, __nv_bfloat16* xin, int B, int S, int F, int D,
                     cudaStream_t stream);
int launch_layernorm_fwd(const __nv_bfloat16* x, const float* gamma, const float* beta, __nv_bfloat16* y,
                         float* mean, float* rstd, int M, int D, float eps, cudaStream_t stream);
size_t layernorm_bwd_workspace_bytes(int M, int D);
int launch_layernorm_bwd(const __nv_bfloat16* dy, const __nv_bfloat16* x, const float* gamma,
                         const float* mean, const float* rstd, __nv_bfloat16* dx, __nv_bfloat16* dx_drop,
                         float* dgamma, float* dbeta, float* dbias, int M, int D, uint32_t drop_thresh,
                         float drop_scale, uint64_t seed, uint64_t stream_id, float* workspace,
                         cudaStream_t stream);
size_t colsum_workspace_bytes(int M, int N);
int launch_colsum(const __nv_bfloat16* x, int ld, int M, int N, float* out, float* workspace,
                  cudaStream_t stream);
int launch_loss_bert(const float* y, int ldy, const float* weather, const uint8_t* mask, int64_t M, int F,
                     float* scratch, float* loss_out, const float* grad_scale, __nv_bfloat16* dy, int lddy,
                     cudaStream_t stream);
int launch_loss_former(const float* y, int ldy, const float* weather, const uint8_t* mask,
                       int64_t mask_stride_b, int64_t mask_stride_s, int B, int S, int F, float beta,
                       float* scratch, float* loss_out, const float* grad_scale, __nv_bfloat16* dy, int lddy,
                       float* mu_out, float* var_out, cudaStream_t stream);
int launch_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, __nv_bfloat16* shadow,
                int64_t n, float lr, double beta1, double beta2, float eps, float weight_decay, int step,
                float grad_scale, cudaStream_t stream);
int launch_adam_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, __nv_bfloat16* shadow, int64_t n,
                    const float* hyper_dev, double beta1, double beta2, float eps, float weight_decay, float grad_scale,
                    cudaStream_t stream);
int launch_step_params_apply(const uint32_t* dev_words, cudaStream_t stream);
constexpr int kMaxTransposeJobs = 72;  // 4 per layer + the head (kernel parameter space: 72 x 32 bytes)
struct TransposeJob {
  const float* w;      // [rows, cols] fp32
  __nv_bfloat16* wt;   // [cols, ld_out] bf16, wt[c, r] = w[r, c]
  int rows, cols, ld_out, tile0;
};
struct TransposeJobs {
  int n;
  TransposeJob job[kMaxTransposeJobs];
};
int launch_cast_transpose_multi(TransposeJobs& jobs, cudaStream_t stream);
int launch_cast_transpose(const float* w, __nv_bfloat16* wt, int rows, int cols, int ld_out, cudaStream_t stream);
int launch_cast_bf16(const float* src, __nv_bfloat16* dst, int64_t n, cudaStream_t stream);

// ---- attention (wm_attn.cu) --------------------------------------------------------------------
// drop_words: optional [B*H][12][384] uint32 keep-bit buffer written by the forward call and read by the backward
// call (required by both when dropout is on); workspace: attn_bwd_workspace_bytes() bytes.
size_t attn_dropout_words_bytes(int B, int S, int H);
size_t attn_bwd_workspace_bytes(int B, int S, int H);
int launch_attn_fwd(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, float* lse, uint32_t* drop_words, int B, int S, int H,
                    int dh, uint32_t drop_thresh, uint64_t seed, uint64_t stream_id, cudaStream_t stream);
int launch_attn_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* ctx, const __nv_bfloat16* dctx, const float* lse,
                    __nv_bfloat16* dqkv, const uint32_t* drop_words, void* workspace, int B, int S, int H, int dh,
                    uint32_t drop_thresh, cudaStream_t stream);


// ---- crop-yield head (wm_yield.cu) -------------------------------------------------------------
struct YieldHeadW;
int yield_head_param_count(int F, int np, int HM);

}  // namespace wm
