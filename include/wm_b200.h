/* wm_b200.h -- C ABI of libwm_b200.so: the B200 (sm_100a) implementation of the WeatherModel
 * training hot path (WeatherBERT / WeatherFormer encoder forward + backward, masks, loss heads, Adam).
 *
 * The reference (Neehan/WeatherModel) has no FFI of its own: its hot path is reached through PyTorch
 * modules. Each entry point below names the reference interface (file:line under /root/reference, or
 * torch:<path> for the torch install the reference calls into) whose arithmetic it replaces. The
 * Python host side (weathermodel_b200/) binds these with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; nothing is retained past the call
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no allocation inside
 *   - return value: 0 = WM_OK, otherwise a WM_ERR_* code (wm_strerror); there is NO CPU fallback and
 *     wm_encoder_create refuses devices that are not compute capability 10.x
 *   - bf16 tensors are passed as void* (raw __nv_bfloat16 storage)
 */
#ifndef WM_B200_H_
#define WM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WM_B200_ABI_VERSION 3

int wm_abi_version(void);
const char* wm_strerror(int code);
/* Reads and clears the device-side error word (non-zero = a bounded mbarrier wait timed out inside a
 * tcgen05 kernel; the code identifies the wait site). Synchronises the device. */
int wm_device_error(void);
/* number of kernels this library has launched so far in this process (host-side counter) */
long long wm_launch_count(void);
/* debug: clock64() timestamps that CTA 0 of the attention kernels records at its phase boundaries (HOST pointer) */
int wm_debug_ticks(long long* out_host, int n);
/* tuning switches (tests / A-B measurements; results are bit-identical either way): "gemm_two_cta" = -1 auto
 * (default) / 0 / 1 selects the cta_group::2 GEMM for M >= 1024, "gemm_epi_warps" = 0 auto (default) / 8 / 16 epilogue
 * warps per GEMM CTA, "gemm_staged" = -1 auto (default) / 0 / 1 writes bf16 outputs through a shared-memory staging
 * tile with coalesced stores instead of thread-per-row stores. A forced value overrides wm_gemm_set_variant and the built-in heuristic. */
int wm_set_option(const char* name, int value);
/* torch.rand grid size for `numel` elements on the current device
 * (torch:include/ATen/native/cuda/DistributionTemplates.h:50-63 calc_execution_policy). */
int wm_rand_grid_x(int64_t numel);

/* ---- masks: StreamingDataset.weatherbert_masking_function / weatherformer_masking_function
 *      (src/pretraining/dataloader/pretraining_dataloader.py:56-66, :68-84). Bit-exact replay of
 *      torch.rand on the CUDA generator given its (seed, philox_offset) and launch grid. ---------- */
int wm_mask_bert(uint64_t seed, uint64_t philox_offset, int grid_x, float masking_prob, int64_t numel,
                 uint8_t* mask_out, float* rand_out /* optional */, void* stream);
int wm_mask_former(uint64_t seed, uint64_t philox_offset, int grid_x, int n_masked_features, int64_t n_samples,
                   int n_features, uint8_t* mask_out /* [n_samples, n_features] */, void* stream);

/* ---- input embedding: WeatherBERT.forward up to positional_encoding
 *      (src/pretraining/models/weatherbert.py:101-115; src/utils/utils.py:63-74;
 *       src/base_models/vanilla_pos_encoding.py:39-58). out: bf16 [B*S, D]; xin: optional bf16 [B*S, 64]. */
int wm_embed_fwd(const float* weather, const uint8_t* mask, int64_t mask_stride_b, int64_t mask_stride_s,
                 const float* year, const float* coords, const float* w_in, const float* b_in,
                 const float* pos_encoding, void* out_bf16, void* xin_bf16, int B, int S, int F, int D,
                 void* stream);

/* ---- dense layers: nn.Linear / MHA in- and out-projection / FFN
 *      (src/pretraining/models/weatherbert.py:34,45-56; torch:nn/modules/transformer.py:944-982).
 *      C[M,N] = epilogue(A[M,K] . B[N,K]^T): +bias, ReLU, dropout, ReLU-gate, +residual (all optional).
 *      sign_bits_out / gate_bits: one bit per output element ("> 0 after ReLU / dropout"), written by the forward
 *      GEMM of a ReLU layer and read by the dgrad GEMM through that ReLU in place of gate_bf16
 *      (wm_gemm_sign_bits_bytes(M, N) bytes; uint16 per (row, 16-column chunk), [row/32][chunk][row%32]). */
typedef struct wm_gemm_epilogue {
  const float* bias;
  int relu;
  float dropout_p;
  uint64_t seed;
  uint64_t stream_id;
  const void* gate_bf16;
  int ld_gate;
  float gate_scale;
  const void* residual_bf16;
  int ld_res;
  void* sign_bits_out;
  const void* gate_bits;
} wm_gemm_epilogue;
size_t wm_gemm_sign_bits_bytes(int M, int N);
int wm_gemm_tn(const void* A_bf16, int lda, const void* B_bf16, int ldb, int M, int N, int K,
               const wm_gemm_epilogue* epilogue /* may be NULL */, void* out, int ld_out, int out_is_fp32,
               int tile_n /* 0 = auto */, void* stream);
/* Record the faster of the (bit-identical) kernel variants for the call site (M, N, K, which epilogue fields are
 * set): two_cta 0 / 1, epi_warps 8 / 16, staged_stores 0 / 1. The host-side tuner (weathermodel_b200/ops.py: tune_gemm_sites) times the
 * variants of an encoder's eight GEMM sites once per shape; untuned sites use a heuristic (K >= 1024: CTA pairs). */
int wm_gemm_set_variant(int M, int N, int K, const wm_gemm_epilogue* epilogue /* may be NULL */, int out_is_fp32,
                        int two_cta, int epi_warps, int staged_stores);
/* dW[Nout,Kout] (+)= A[Mtok,Nout]^T . B[Mtok,Kout]; workspace from wm_gemm_wgrad_workspace_bytes.
 * dbias (optional, [Nout]): column sums of A over the tokens (the bias gradient), fused into the same kernel. */
size_t wm_gemm_wgrad_workspace_bytes(int Mtok, int Nout, int Kout);
int wm_gemm_wgrad(const void* A_bf16, int lda, const void* B_bf16, int ldb, int Mtok, int Nout, int Kout,
                  float* dW, int accumulate, float* workspace, float* dbias /* optional */, void* stream);
/* hardware probe of the unswizzled UMMA operand layouts used by the attention kernels (tests only) */
int wm_umma_probe(const void* A_bf16, const void* B_bf16, float* D, int N, int K, int a_mn_major, int b_mn_major,
                  void* stream);

/* ---- attention: F.scaled_dot_product_attention inside nn.MultiheadAttention
 *      (torch:nn/functional.py:6666-6696). qkv: bf16 [B*S, 3*H*dh] (Q|K|V, head-major columns);
 *      ctx: bf16 [B*S, H*dh]; lse: fp32 [B*H, S]. S <= 384, dh in {12..48} a multiple of 4, H*dh a multiple of 8.
 *      With dropout_p > 0 the forward call also writes its keep decisions (one bit per (query, key), buffer of
 *      wm_attn_dropout_words_bytes) and the backward call reads them back instead of re-deriving the Philox stream;
 *      drop_words may be NULL when dropout_p == 0. workspace: wm_attn_bwd_workspace_bytes bytes (row statistics). */
size_t wm_attn_dropout_words_bytes(int B, int S, int H);
size_t wm_attn_bwd_workspace_bytes(int B, int S, int H);
int wm_attn_fwd(const void* qkv_bf16, void* ctx_bf16, float* lse, void* drop_words, int B, int S, int H, int dh,
                float dropout_p, uint64_t seed, uint64_t stream_id, void* stream);
int wm_attn_bwd(const void* qkv_bf16, const void* ctx_bf16, const void* dctx_bf16, const float* lse,
                void* dqkv_bf16, const void* drop_words, void* workspace, int B, int S, int H, int dh, float dropout_p,
                void* stream);

/* ---- LayerNorm (post-LN norm1 / norm2, torch:nn/modules/transformer.py:951-957; eps 1e-5) ------------- */
int wm_layernorm_fwd(const void* x_bf16, const float* gamma, const float* beta, void* y_bf16, float* mean,
                     float* rstd, int M, int D, float eps, void* stream);
size_t wm_layernorm_bwd_workspace_bytes(int M, int D);
int wm_layernorm_bwd(const void* dy_bf16, const void* x_bf16, const float* gamma, const float* mean,
                     const float* rstd, void* dx_bf16, void* dx_dropped_bf16 /* NULL if p == 0 */, float* dgamma,
                     float* dbeta, float* dbias /* optional: column sums of the (dropped) dx; NULL selects a leaner
                     kernel (15 instead of 8 row warps per SM) -- the encoder takes this gradient from wm_gemm_wgrad */, int M, int D,
                     float dropout_p, uint64_t seed, uint64_t stream_id, float* workspace, void* stream);
size_t wm_colsum_workspace_bytes(int M, int N);
int wm_colsum(const void* x_bf16, int ld, int M, int N, float* out, float* workspace, void* stream);

/* ---- loss heads --------------------------------------------------------------------------------------
 * wm_loss_bert:   WeatherBertTrainer.compute_train_loss (src/pretraining/trainers/weatherbert_trainer.py:46-62):
 *                 loss_out[0] = sum m (x - y)^2 / sum m, loss_out[1] = sum m; dy = 2 m (y - x) / sum m.
 * wm_loss_former: WeatherFormer head + WeatherFormerTrainer.compute_elbo_loss
 *                 (src/pretraining/models/weatherformer.py:87-92; src/pretraining/trainers/
 *                 weatherformer_trainer.py:68-111; src/utils/losses.py:10-47):
 *                 loss_out = {total, reconstruction, kl_term, sum m}; dy = dLoss/d[mu | logvar].
 * y: fp32 [M, ldy]; dy: bf16 [M, lddy] (pad columns zeroed); scratch: >= 4 * 592 floats. */
int wm_loss_bert(const float* y, int ldy, const float* weather, const uint8_t* mask, int64_t M, int F,
                 float* scratch, float* loss_out, void* dy_bf16, int lddy, void* stream);
int wm_loss_former(const float* y, int ldy, const float* weather, const uint8_t* mask, int64_t mask_stride_b,
                   int64_t mask_stride_s, int B, int S, int F, float beta, float* scratch, float* loss_out,
                   void* dy_bf16, int lddy, float* mu_out /* optional [M,F] */, float* var_out /* optional */,
                   void* stream);
/* The autograd backward of the two losses (torch: `loss.backward()` reaching MSELoss / the ELBO sum, same files):
 * dy = grad_scale[0] * dLoss/dY from the partial sums an earlier wm_loss_* call (with dy_bf16 == NULL) left in
 * `scratch`. grad_scale is a DEVICE scalar (the upstream gradient; NULL = 1): no host read-back, no fp32 copy of dy. */
int wm_loss_bert_grad(const float* y, int ldy, const float* weather, const uint8_t* mask, int64_t M, int F,
                      const float* scratch, const float* grad_scale, void* dy_bf16, int lddy, void* stream);
int wm_loss_former_grad(const float* y, int ldy, const float* weather, const uint8_t* mask, int64_t mask_stride_b,
                        int64_t mask_stride_s, int B, int S, int F, float beta, const float* scratch,
                        const float* grad_scale, void* dy_bf16, int lddy, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Crop-yield head (BASELINE configs[5]), one forward and one backward kernel on the raw encoder head output:
 *   WeatherBERTYieldModel._impute_weather + yield_model  src/crop_yield/models/weatherbert_yield_model.py:40-67
 *   WeatherFormerYieldModel.forward (z = mu + sqrt(var) eps)  src/crop_yield/models/weatherformer_yield_model.py:58-60
 * y: fp32 [B*S, ldy] raw head output ([pred | pad] or [mu | logvar | pad]); eps: fp32 [B,S,F] standard-normal noise
 * (WeatherFormer only; drawn by the caller so that the torch generator is consumed exactly as randn_like does);
 * y_past [B, n_past]; the head's parameters as the reference modules hold them: att_w1 [16,F], att_b1 [16],
 * att_w2 [1,16], att_b2 [1], mlp_w1 [HM, F + n_past], mlp_b1 [HM], mlp_w2 [1,HM], mlp_b2 [1].
 * Forward: z_out fp32 [B,S,F] (imputed / sampled weather; also what backward needs), pred fp32 [B].
 * Backward: dy fp32 [B*S, ldy] = dLoss/dy through the head (pad columns zeroed); grads: the parameter gradients
 * flattened in the order above (wm_yield_head_param_count floats); partial: B * that many floats of workspace.
 * Deterministic (fixed-order reductions): yield_main sets torch.use_deterministic_algorithms(True). S <= 384, F <= 32. */
int wm_yield_head_param_count(int F, int n_past, int HM);
int wm_yield_head_fwd(const float* y, int ldy, int is_former, const float* weather, const uint8_t* mask,
                      int64_t mask_stride_b, int64_t mask_stride_s, const float* eps, const float* y_past, int n_past,
                      const float* att_w1, const float* att_b1, const float* att_w2, const float* att_b2,
                      const float* mlp_w1, const float* mlp_b1, const float* mlp_w2, const float* mlp_b2,
                      float* z_out, float* pred, int B, int S, int F, int HM, void* stream);
int wm_yield_head_bwd(const float* dpred, const float* y, int ldy, int is_former, const uint8_t* mask,
                      int64_t mask_stride_b, int64_t mask_stride_s, const float* eps, const float* z_saved,
                      const float* y_past, int n_past, const float* att_w1, const float* att_b1, const float* att_w2,
                      const float* att_b2, const float* mlp_w1, const float* mlp_b1, const float* mlp_w2,
                      const float* mlp_b2, float* dy, float* partial, float* grads, int B, int S, int F, int HM,
                      void* stream);

/* ---- optimiser: torch.optim.Adam as built at src/base_trainer/base_trainer.py:337 ------------------- */
/* beta1 / beta2 are doubles (ABI 3): torch.optim.Adam forms 1 - beta^step from Python floats, i.e. in double
 * precision from the decimal value; bias corrections formed from the float-rounded betas are off by ~1e-5 relative. */
int wm_adam_fused(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                  void* shadow_bf16 /* optional */, int64_t n, float lr, double beta1, double beta2, float eps,
                  float weight_decay, int step, float grad_scale, void* stream);
/* The same update with the step-dependent scalars read from DEVICE memory: hyper_dev = {lr, 1 - beta1^t,
 * sqrt(1 - beta2^t)} (3 floats). For a training step captured in a CUDA graph (the BaseTrainer step body,
 * src/base_trainer/base_trainer.py:239-252, replayed without host work): the launch is replayed unchanged while the
 * host refreshes the three numbers in a pinned buffer that a captured copy brings over. */
int wm_adam_fused_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16 /* optional */,
                      int64_t n, const float* hyper_dev, double beta1, double beta2, float eps, float weight_decay,
                      float grad_scale, void* stream);
/* Installs three per-replay words (device memory; NULL = zeros) that every dropout site of the kernels above folds
 * into its keys: replays of a captured step then draw fresh masks although their launch parameters are frozen.
 * Zero words (the initial state) reproduce the un-captured behaviour exactly. */
int wm_step_params_apply(const void* dev_words /* 3 x uint32 */, void* stream);

/* ---- the encoder as one object: WeatherBERT.forward / WeatherFormer.forward + autograd backward
 *      (src/pretraining/models/weatherbert.py:84-121; src/pretraining/models/weatherformer.py:60-94) ---- */
typedef struct wm_encoder_config {
  int B, S, F;      /* per-GPU batch, sequence length (<= 384), weather features (31) */
  int D, H, L, FF;  /* hidden, heads, layers, feed-forward */
  int out_dim;      /* 31 (WeatherBERT) or 62 (WeatherFormer: mu | logvar) */
  float dropout_p;  /* 0.1 in the reference (nn.TransformerEncoderLayer default) */
  float ln_eps;     /* 1e-5 */
  int eval_only;    /* 1: validation / inference handle (the reference's model.eval() + torch.no_grad(),
                       src/base_trainer/base_trainer.py:262-285): workspace holds ONE layer of activations and no
                       backward temporaries; wm_encoder_forward must be called with save_for_backward = 0 */
} wm_encoder_config;
typedef struct wm_encoder wm_encoder;

/* flat fp32 parameter / gradient buffer: reference named_parameters() order, each tensor start aligned
 * to 64 floats, gaps zero. Returns the number of tensors and fills their start offsets. */
int64_t wm_encoder_param_count(const wm_encoder_config* cfg);
int wm_encoder_param_layout(const wm_encoder_config* cfg, int64_t* offsets, int max_n);
size_t wm_encoder_workspace_bytes(const wm_encoder_config* cfg);
int wm_encoder_create(const wm_encoder_config* cfg, void* workspace, size_t workspace_bytes, wm_encoder** out);
int wm_encoder_destroy(wm_encoder* enc);
int wm_encoder_refresh_weights(wm_encoder* enc, const float* params, void* stream);
/* The bf16 shadow of the flat parameter buffer inside the workspace (same offsets as the fp32 buffer): pass it as
 * shadow_bf16 to wm_adam_fused / wm_adam_fused_dev and the optimiser writes it while it updates the parameters; then only
 * the transposed copies have to be rebuilt before the next forward: */
void* wm_encoder_shadow(wm_encoder* enc);
int wm_encoder_refresh_transposes(wm_encoder* enc, const float* params, void* stream);
/* y_out: fp32 [B*S, 32 or 64] (columns >= out_dim are padding). training: dropout live (nn.Module.train()).
 * save_for_backward = 1 keeps every layer's activations for wm_encoder_backward_*; 0 is the lean schedule of a
 * forward that no backward follows (torch.no_grad()): all layers reuse one set of buffers and nothing that only the
 * backward pass reads is written; the backward entry points then return an error until the next saving forward. */
int wm_encoder_forward(wm_encoder* enc, const float* params, const float* weather, const uint8_t* mask,
                       int64_t mask_stride_b, int64_t mask_stride_s, const float* year, const float* coords,
                       const float* pos_encoding, float* y_out, int training, int save_for_backward, uint64_t seed,
                       uint64_t step, void* stream);
int wm_encoder_backward_head(wm_encoder* enc, const void* dy_bf16, float* grads, void* stream);
int wm_encoder_backward_layers(wm_encoder* enc, const float* params, int layer_hi, int layer_lo, float* grads,
                               void* stream);
int wm_encoder_backward_embed(wm_encoder* enc, float* grads, void* stream);
const void* wm_encoder_activation(wm_encoder* enc, int layer, int which);

#ifdef __cplusplus
}
#endif
#endif /* WM_B200_H_ */
