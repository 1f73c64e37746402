// wm_api.cu -- extern "C" surface of libwm_b200.so (declared in include/wm_b200.h): argument checks and
// the mapping from plain pointers / sizes onto the internal launchers. No torch types anywhere.
#include "wm_kernels.h"
#include "../../include/wm_b200.h"

using namespace wm;

namespace {
inline cudaStream_t S_(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline const __nv_bfloat16* CB(const void* p) { return reinterpret_cast<const __nv_bfloat16*>(p); }
inline __nv_bfloat16* MB(void* p) { return reinterpret_cast<__nv_bfloat16*>(p); }
inline uint32_t thresh16(float p) { return p > 0.0f ? static_cast<uint32_t>(p * 65536.0f + 0.5f) : 0u; }
inline float keep_scale(uint32_t t) { return wm::drop_keep_scale(t); }  // the kernels compare 15 bits: p_eff = round(32768 p) / 32768
}  // namespace

extern "C" {

int wm_abi_version(void) { return WM_B200_ABI_VERSION; }

const char* wm_strerror(int code) {
  switch (code) {
    case WM_OK: return "ok";
    case WM_ERR_SHAPE: return "unsupported shape";
    case WM_ERR_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case WM_ERR_CUDA: return "CUDA runtime error";
    case WM_ERR_DRIVER: return "CUDA driver entry point (cuTensorMapEncodeTiled) failed";
    case WM_ERR_ARG: return "invalid argument";
    case WM_ERR_DEVICE: return "device is not compute capability 10.x (sm_100a required, no fallback)";
    default: return "unknown error";
  }
}

int wm_device_error(void) {
  unsigned int v = 0, zero = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&v, g_wm_dev_error, sizeof(v)) != cudaSuccess) return -1;
  if (v) cudaMemcpyToSymbol(g_wm_dev_error, &zero, sizeof(zero));
  return static_cast<int>(v);
}

long long wm_launch_count(void) { return g_wm_launches; }

int wm_debug_ticks(long long* out_host, int n) {
  if (!out_host || n <= 0 || n > 64) return WM_ERR_ARG;
  if (cudaDeviceSynchronize() != cudaSuccess) return WM_ERR_CUDA;
  return cudaMemcpyFromSymbol(out_host, g_wm_ticks, sizeof(long long) * n) == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

namespace wm { extern int g_gemm_two_cta, g_gemm_epi_warps, g_gemm_staged, g_ln_bwd_width, g_ln_fwd_width, g_wgrad_bn, g_ln_bwd_rows, g_wgrad_mh; }
int wm_set_option(const char* name, int value) {
  if (!name) return WM_ERR_ARG;
  struct Opt { const char* name; int* slot; };
  const Opt opts[] = {{"gemm_two_cta", &wm::g_gemm_two_cta}, {"gemm_epi_warps", &wm::g_gemm_epi_warps},
                      {"gemm_staged", &wm::g_gemm_staged}, {"ln_bwd_width", &wm::g_ln_bwd_width},
                      {"ln_fwd_width", &wm::g_ln_fwd_width}, {"wgrad_bn", &wm::g_wgrad_bn}, {"ln_bwd_rows", &wm::g_ln_bwd_rows}, {"wgrad_mh", &wm::g_wgrad_mh}};
  for (const Opt& o : opts) {
    const char* a = name;
    const char* b = o.name;
    while (*a && *a == *b) { ++a; ++b; }
    if (!*a && !*b) {
      *o.slot = value;
      return WM_OK;
    }
  }
#ifdef WM_DIAG
  {
    const char* a = name;
    const char* b = "gemm_diag";
    while (*a && *a == *b) { ++a; ++b; }
    if (!*a && !*b) return wm::gemm_set_diag(value);
  }
#endif
  return WM_ERR_ARG;
}

int wm_rand_grid_x(int64_t numel) {
  int dev = 0, sms = 0, tpm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&tpm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
  const int64_t want = (numel + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sms) * (tpm / 256);
  return static_cast<int>(want < cap ? want : cap);
}

int wm_mask_bert(uint64_t seed, uint64_t philox_offset, int grid_x, float masking_prob, int64_t numel,
                 uint8_t* mask_out, float* rand_out, void* stream) {
  if (!mask_out) return WM_ERR_ARG;
  return launch_mask_bert(seed, philox_offset, grid_x, masking_prob, numel, mask_out, rand_out, S_(stream));
}
int wm_mask_former(uint64_t seed, uint64_t philox_offset, int grid_x, int n_masked_features, int64_t n_samples,
                   int n_features, uint8_t* mask_out, void* stream) {
  if (!mask_out) return WM_ERR_ARG;
  return launch_mask_former(seed, philox_offset, grid_x, n_masked_features, n_samples, n_features, mask_out,
                            S_(stream));
}

int wm_embed_fwd(const float* weather, const uint8_t* mask, int64_t mask_stride_b, int64_t mask_stride_s,
                 const float* year, const float* coords, const float* w_in, const float* b_in,
                 const float* pos_encoding, void* out_bf16, void* xin_bf16, int B, int S, int F, int D,
                 void* stream) {
  if (!weather || !mask || !year || !coords || !w_in || !b_in || !pos_encoding || !out_bf16) return WM_ERR_ARG;
  return launch_embed_fwd(weather, mask, mask_stride_b, mask_stride_s, year, coords, w_in, b_in, pos_encoding,
                          MB(out_bf16), MB(xin_bf16), B, S, F, D, S_(stream));
}

size_t wm_gemm_sign_bits_bytes(int M, int N) { return gemm_sign_bits_bytes(M, N); }
static int fill_epilogue(GemmEpilogue& ep, const wm_gemm_epilogue* e) {
  if (!e) return WM_OK;
  ep.bias = e->bias;
  ep.relu = e->relu;
  ep.drop_thresh = thresh16(e->dropout_p);
  ep.drop_scale = keep_scale(ep.drop_thresh);
  ep.seed = e->seed;
  ep.stream = e->stream_id;
  ep.gate = CB(e->gate_bf16);
  ep.ld_gate = e->ld_gate;
  ep.gate_scale = e->gate_scale;
  ep.residual = CB(e->residual_bf16);
  ep.ld_res = e->ld_res;
  ep.sign_bits_out = static_cast<uint16_t*>(e->sign_bits_out);
  ep.gate_bits = static_cast<const uint16_t*>(e->gate_bits);
  if ((ep.gate && (ep.ld_gate & 7)) || (ep.residual && (ep.ld_res & 7))) return WM_ERR_ALIGN;
  return WM_OK;
}
int wm_gemm_tn(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const wm_gemm_epilogue* e,
               void* out, int ld_out, int out_is_fp32, int tile_n, void* stream) {
  if (!A || !B || !out) return WM_ERR_ARG;
  GemmEpilogue ep;
  const int rc = fill_epilogue(ep, e);
  if (rc) return rc;
  ep.out = out;
  ep.ld_out = ld_out;
  return launch_gemm_tn(A, lda, B, ldb, M, N, K, ep, out_is_fp32, tile_n, S_(stream));
}
int wm_gemm_set_variant(int M, int N, int K, const wm_gemm_epilogue* e, int out_is_fp32, int two_cta, int epi_warps,
                        int staged) {
  GemmEpilogue ep;
  const int rc = fill_epilogue(ep, e);
  if (rc) return rc;
  return gemm_set_variant(M, N, K, gemm_signature(ep, out_is_fp32), two_cta, epi_warps, staged);
}
size_t wm_gemm_wgrad_workspace_bytes(int Mtok, int Nout, int Kout) { return wgrad_workspace_bytes(Mtok, Nout, Kout); }
int wm_gemm_wgrad(const void* A, int lda, const void* B, int ldb, int Mtok, int Nout, int Kout, float* dW,
                  int accumulate, float* workspace, float* dbias, void* stream) {
  if (!A || !B || !dW || !workspace) return WM_ERR_ARG;
  return launch_gemm_wgrad(A, lda, B, ldb, Mtok, Nout, Kout, dW, accumulate, workspace, dbias, S_(stream));
}
int wm_umma_probe(const void* A, const void* B, float* D, int N, int K, int a_mn, int b_mn, void* stream) {
  if (!A || !B || !D) return WM_ERR_ARG;
  return launch_umma_probe(A, B, D, N, K, a_mn, b_mn, S_(stream));
}

size_t wm_attn_dropout_words_bytes(int B, int S, int H) { return attn_dropout_words_bytes(B, S, H); }
size_t wm_attn_bwd_workspace_bytes(int B, int S, int H) { return attn_bwd_workspace_bytes(B, S, H); }
int wm_attn_fwd(const void* qkv, void* ctx, float* lse, void* drop_words, int B, int S, int H, int dh, float dropout_p,
                uint64_t seed, uint64_t stream_id, void* stream) {
  if (!qkv || !ctx) return WM_ERR_ARG;
  return launch_attn_fwd(CB(qkv), MB(ctx), lse, static_cast<uint32_t*>(drop_words), B, S, H, dh, thresh16(dropout_p), seed,
                         stream_id, S_(stream));
}
int wm_attn_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv, const void* drop_words,
                void* workspace, int B, int S, int H, int dh, float dropout_p, void* stream) {
  if (!qkv || !ctx || !dctx || !lse || !dqkv) return WM_ERR_ARG;
  return launch_attn_bwd(CB(qkv), CB(ctx), CB(dctx), lse, MB(dqkv), static_cast<const uint32_t*>(drop_words), workspace, B,
                         S, H, dh, thresh16(dropout_p), S_(stream));
}

int wm_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M,
                     int D, float eps, void* stream) {
  if (!x || !gamma || !beta || !y) return WM_ERR_ARG;
  return launch_layernorm_fwd(CB(x), gamma, beta, MB(y), mean, rstd, M, D, eps, S_(stream));
}
size_t wm_layernorm_bwd_workspace_bytes(int M, int D) { return layernorm_bwd_workspace_bytes(M, D); }
int wm_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                     void* dx, void* dx_dropped, float* dgamma, float* dbeta, float* dbias, int M, int D,
                     float dropout_p, uint64_t seed, uint64_t stream_id, float* workspace, void* stream) {
  if (!dy || !x || !gamma || !mean || !rstd || !dx || !workspace) return WM_ERR_ARG;
  const uint32_t t = thresh16(dropout_p);
  return launch_layernorm_bwd(CB(dy), CB(x), gamma, mean, rstd, MB(dx), MB(dx_dropped), dgamma, dbeta, dbias, M, D,
                              t, keep_scale(t), seed, stream_id, workspace, S_(stream));
}
size_t wm_colsum_workspace_bytes(int M, int N) { return colsum_workspace_bytes(M, N); }
int wm_colsum(const void* x, int ld, int M, int N, float* out, float* workspace, void* stream) {
  if (!x || !out || !workspace) return WM_ERR_ARG;
  return launch_colsum(CB(x), ld, M, N, out, workspace, S_(stream));
}

int wm_loss_bert(const float* y, int ldy, const float* weather, const uint8_t* mask, int64_t M, int F,
                 float* scratch, float* loss_out, void* dy, int lddy, void* stream) {
  if (!y || !weather || !mask || !scratch || !loss_out) return WM_ERR_ARG;
  return launch_loss_bert(y, ldy, weather, mask, M, F, scratch, loss_out, nullptr, MB(dy), lddy, S_(stream));
}
int wm_loss_bert_grad(const float* y, int ldy, const float* weather, const uint8_t* mask, int64_t M, int F,
                      const float* scratch, const float* grad_scale, void* dy, int lddy, void* stream) {
  if (!y || !weather || !mask || !scratch || !dy) return WM_ERR_ARG;
  return launch_loss_bert(y, ldy, weather, mask, M, F, const_cast<float*>(scratch), nullptr, grad_scale, MB(dy), lddy,
                          S_(stream));
}
int wm_loss_former(const float* y, int ldy, const float* weather, const uint8_t* mask, int64_t mask_stride_b,
                   int64_t mask_stride_s, int B, int S, int F, float beta, float* scratch, float* loss_out, void* dy,
                   int lddy, float* mu_out, float* var_out, void* stream) {
  if (!y || !weather || !mask || !scratch || !loss_out) return WM_ERR_ARG;
  if ((mu_out == nullptr) != (var_out == nullptr)) return WM_ERR_ARG;
  return launch_loss_former(y, ldy, weather, mask, mask_stride_b, mask_stride_s, B, S, F, beta, scratch, loss_out,
                            nullptr, MB(dy), lddy, mu_out, var_out, S_(stream));
}
int wm_loss_former_grad(const float* y, int ldy, const float* weather, const uint8_t* mask, int64_t mask_stride_b,
                        int64_t mask_stride_s, int B, int S, int F, float beta, const float* scratch,
                        const float* grad_scale, void* dy, int lddy, void* stream) {
  if (!y || !weather || !mask || !scratch || !dy) return WM_ERR_ARG;
  return launch_loss_former(y, ldy, weather, mask, mask_stride_b, mask_stride_s, B, S, F, beta,
                            const_cast<float*>(scratch), nullptr, grad_scale, MB(dy), lddy, nullptr, nullptr, S_(stream));
}

int wm_adam_fused(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow, int64_t n,
                  float lr, double beta1, double beta2, float eps, float weight_decay, int step, float grad_scale,
                  void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq) return WM_ERR_ARG;
  return launch_adam(param, grad, exp_avg, exp_avg_sq, MB(shadow), n, lr, beta1, beta2, eps, weight_decay, step,
                     grad_scale, S_(stream));
}

int wm_adam_fused_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow, int64_t n,
                      const float* hyper_dev, double beta1, double beta2, float eps, float weight_decay, float grad_scale,
                      void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !hyper_dev) return WM_ERR_ARG;
  return launch_adam_dev(param, grad, exp_avg, exp_avg_sq, MB(shadow), n, hyper_dev, beta1, beta2, eps, weight_decay,
                         grad_scale, S_(stream));
}

int wm_step_params_apply(const void* dev_words, void* stream) {
  return launch_step_params_apply(reinterpret_cast<const uint32_t*>(dev_words), S_(stream));
}

int wm_yield_head_param_count(int F, int n_past, int HM) {
  if (F <= 0 || F > 32 || n_past < 0 || HM <= 0 || HM > 128) return -1;
  return yield_head_param_count(F, n_past, HM);
}

int wm_yield_head_fwd(const float* y, int ldy, int is_former, const float* weather, const uint8_t* mask,
                      int64_t mask_stride_b, int64_t mask_stride_s, const float* eps, const float* y_past, int n_past,
                      const float* att_w1, const float* att_b1, const float* att_w2, const float* att_b2,
                      const float* mlp_w1, const float* mlp_b1, const float* mlp_w2, const float* mlp_b2,
                      float* z_out, float* pred, int B, int S, int F, int HM, void* stream) {
  if (!att_w1 || !att_b1 || !att_w2 || !att_b2 || !mlp_w1 || !mlp_b1 || !mlp_w2 || !mlp_b2) return WM_ERR_ARG;
  const YieldHeadW W{att_w1, att_b1, att_w2, att_b2, mlp_w1, mlp_b1, mlp_w2, mlp_b2};
  return launch_yield_head_fwd(y, ldy, is_former, weather, mask, mask_stride_b, mask_stride_s, eps, y_past, n_past, W,
                               z_out, pred, B, S, F, HM, S_(stream));
}

int wm_yield_head_bwd(const float* dpred, const float* y, int ldy, int is_former, const uint8_t* mask,
                      int64_t mask_stride_b, int64_t mask_stride_s, const float* eps, const float* z_saved,
                      const float* y_past, int n_past, const float* att_w1, const float* att_b1, const float* att_w2,
                      const float* att_b2, const float* mlp_w1, const float* mlp_b1, const float* mlp_w2,
                      const float* mlp_b2, float* dy, float* partial, float* grads, int B, int S, int F, int HM,
                      void* stream) {
  if (!att_w1 || !att_b1 || !att_w2 || !att_b2 || !mlp_w1 || !mlp_b1 || !mlp_w2 || !mlp_b2) return WM_ERR_ARG;
  const YieldHeadW W{att_w1, att_b1, att_w2, att_b2, mlp_w1, mlp_b1, mlp_w2, mlp_b2};
  return launch_yield_head_bwd(dpred, y, ldy, is_former, mask, mask_stride_b, mask_stride_s, eps, z_saved, y_past, n_past,
                               W, dy, partial, grads, B, S, F, HM, S_(stream));
}

}  // extern "C"
