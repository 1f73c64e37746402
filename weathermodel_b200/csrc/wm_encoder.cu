// wm_encoder.cu -- the whole WeatherBERT / WeatherFormer encoder step as a fixed kernel schedule.
//
// Host-side orchestration only (no kernels here): embedding -> L x post-LN transformer layer -> output
// head, and the matching backward, each as a sequence of launches of the kernels in wm_gemm.cu,
// wm_attn.cu and wm_elementwise.cu on ONE caller-supplied stream. Mirrors
//   WeatherBERT.forward                      src/pretraining/models/weatherbert.py:101-121
//   nn.TransformerEncoderLayer (post-LN)     torch/nn/modules/transformer.py:944-982
//   autograd backward of the same            
// Memory: the caller owns one workspace (wm_encoder_workspace_bytes) that holds the bf16 weight
// shadows (+ transposes for dgrad), the per-layer saved activations and the backward temporaries.
// Parameters and gradients are flat fp32 buffers in reference named_parameters() order, each tensor
// start aligned to 64 floats (wm_encoder_param_layout).
#include "wm_kernels.h"
#include "../../include/wm_b200.h"

#include <new>
#include <vector>

namespace wm {

constexpr int kParamAlign = 64;   // floats
constexpr int kXinCols = 64;      // bf16 padded input row (34 -> 64) for the in_proj wgrad

struct LayerParams {  // offsets (floats) into the flat parameter / gradient buffer
  int64_t w_qkv, b_qkv, w_o, b_o, w1, b1, w2, b2, g1, be1, g2, be2;
};
struct ParamLayout {
  int64_t w_in, b_in, w_out, b_out, total;
  std::vector<LayerParams> layers;
};

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static ParamLayout make_layout(const wm_encoder_config& c) {
  ParamLayout p;
  int64_t off = 0;
  auto take = [&](int64_t n) {
    const int64_t o = off;
    off = align_up(off + n, kParamAlign);
    return o;
  };
  const int64_t D = c.D, FF = c.FF, Fin = c.F + 3;
  p.w_in = take(D * Fin);
  p.b_in = take(D);
  p.layers.resize(c.L);
  for (int l = 0; l < c.L; ++l) {
    LayerParams& q = p.layers[l];
    q.w_qkv = take(3 * D * D);
    q.b_qkv = take(3 * D);
    q.w_o = take(D * D);
    q.b_o = take(D);
    q.w1 = take(FF * D);
    q.b1 = take(FF);
    q.w2 = take(D * FF);
    q.b2 = take(D);
    q.g1 = take(D);
    q.be1 = take(D);
    q.g2 = take(D);
    q.be2 = take(D);
  }
  p.w_out = take(static_cast<int64_t>(c.out_dim) * D);
  p.b_out = take(c.out_dim);
  p.total = off;
  return p;
}

struct LayerAct {
  __nv_bfloat16 *x, *qkv, *ctx, *r1, *u, *h, *r2;
  float *lse, *mean1, *rstd1, *mean2, *rstd2;
  uint32_t* dropw;  // attention dropout keep bits (forward -> backward); nullptr when the model has no dropout
  uint16_t* hbits;  // sign bits of h (post ReLU / dropout): the gate of the linear2 dgrad
};
struct LayerWt {  // transposed bf16 copies for dgrad
  __nv_bfloat16 *wqkv_t, *wo_t, *w1_t, *w2_t;
};

}  // namespace wm

struct wm_encoder {
  wm_encoder_config cfg;
  wm::ParamLayout lay;
  int64_t M;
  int outP;
  uint8_t* ws;
  size_t ws_bytes;
  __nv_bfloat16* shadow;  // bf16 copy of the flat parameter buffer (same offsets)
  std::vector<wm::LayerWt> wt;
  __nv_bfloat16* wout_t;  // [D, outP]
  std::vector<wm::LayerAct> act;
  __nv_bfloat16 *x_final, *xin;
  __nv_bfloat16 *gX, *gR, *gF, *gU, *gC, *gH, *gQKV;
  float* scratch;  // wgrad partials / colsum / LN-bwd partials
  size_t scratch_bytes;
  uint64_t seed, step;
  int training;
  int saved;  // the last forward call kept its activations: backward may run
};

namespace wm {

static size_t scratch_need(const wm_encoder_config& c, int64_t M, int outP) {
  size_t need = 0;
  auto upd = [&](size_t v) { if (v > need) need = v; };
  const int Mi = static_cast<int>(M);
  upd(wgrad_workspace_bytes(Mi, 3 * c.D, c.D));
  upd(wgrad_workspace_bytes(Mi, c.D, c.D));
  upd(wgrad_workspace_bytes(Mi, c.FF, c.D));
  upd(wgrad_workspace_bytes(Mi, c.D, c.FF));
  upd(wgrad_workspace_bytes(Mi, outP, c.D));
  upd(wgrad_workspace_bytes(Mi, c.D, kXinCols));
  upd(colsum_workspace_bytes(Mi, 3 * c.D));
  upd(colsum_workspace_bytes(Mi, c.FF));
  upd(layernorm_bwd_workspace_bytes(Mi, c.D));
  upd(attn_bwd_workspace_bytes(c.B, c.S, c.H));
  return need;
}

// carve (or, with base == nullptr, just size) the workspace
static size_t carve(wm_encoder* e, uint8_t* base) {
  const wm_encoder_config& c = e->cfg;
  const int64_t M = e->M, D = c.D, FF = c.FF;
  size_t off = 0;
  auto take = [&](size_t bytes) -> uint8_t* {
    uint8_t* p = base ? base + off : nullptr;
    off = (off + bytes + 255) & ~static_cast<size_t>(255);
    return p;
  };
  auto bf = [&](int64_t n) { return reinterpret_cast<__nv_bfloat16*>(take(static_cast<size_t>(n) * 2)); };
  auto f32 = [&](int64_t n) { return reinterpret_cast<float*>(take(static_cast<size_t>(n) * 4)); };
  e->shadow = bf(e->lay.total);
  e->wt.resize(c.L);
  for (int l = 0; l < c.L; ++l) {
    e->wt[l].wqkv_t = bf(D * 3 * D);
    e->wt[l].wo_t = bf(D * D);
    e->wt[l].w1_t = bf(D * FF);
    e->wt[l].w2_t = bf(FF * D);
  }
  e->wout_t = bf(D * e->outP);
  // eval_only: one layer's worth of activation buffers (every layer reuses them, x ping-pongs between act[0].x and
  // x_final), no saved statistics / sign bits / keep bits, no backward temporaries: 12 M D bytes instead of
  // (12 L + 14) M D + ... -- 2.6 GB instead of 22 GB at WeatherFormer large, B = 512
  const int n_act = c.eval_only ? 1 : c.L;
  e->act.resize(n_act);
  for (int l = 0; l < n_act; ++l) {
    LayerAct& a = e->act[l];
    a.x = bf(M * D);
    a.qkv = bf(M * 3 * D);
    a.ctx = bf(M * D);
    a.r1 = bf(M * D);
    a.u = bf(M * D);
    a.h = bf(M * FF);
    a.r2 = bf(M * D);
    if (c.eval_only) {
      a.lse = a.mean1 = a.rstd1 = a.mean2 = a.rstd2 = nullptr;
      a.hbits = nullptr;
      a.dropw = nullptr;
      continue;
    }
    a.lse = f32(static_cast<int64_t>(c.B) * c.H * c.S);
    a.hbits = reinterpret_cast<uint16_t*>(take(gemm_sign_bits_bytes(static_cast<int>(M), c.FF)));
    a.dropw = c.dropout_p > 0.0f ? reinterpret_cast<uint32_t*>(take(attn_dropout_words_bytes(c.B, c.S, c.H))) : nullptr;
    a.mean1 = f32(M);
    a.rstd1 = f32(M);
    a.mean2 = f32(M);
    a.rstd2 = f32(M);
  }
  e->x_final = bf(M * D);
  if (c.eval_only) {
    e->xin = e->gX = e->gR = e->gF = e->gU = e->gC = e->gH = e->gQKV = nullptr;
    e->scratch_bytes = 0;
    e->scratch = nullptr;
    return off;
  }
  e->xin = bf(M * kXinCols);
  e->gX = bf(M * D);
  e->gR = bf(M * D);
  e->gF = bf(M * D);
  e->gU = bf(M * D);
  e->gC = bf(M * D);
  e->gH = bf(M * FF);
  e->gQKV = bf(M * 3 * D);
  e->scratch_bytes = scratch_need(c, M, e->outP);
  e->scratch = reinterpret_cast<float*>(take(e->scratch_bytes));
  return off;
}

static int check_cfg(const wm_encoder_config* c) {
  if (!c) return WM_ERR_ARG;
  if (c->B <= 0 || c->S <= 0 || c->S > 384 || c->F <= 0 || c->F > 37 || c->L <= 0 || c->H <= 0) return WM_ERR_SHAPE;
  if (c->D <= 0 || c->D % c->H || (c->D & 7) || c->D > 768) return WM_ERR_SHAPE;
  const int dh = c->D / c->H;
  if ((dh & 3) || dh < 12 || dh > 48) return WM_ERR_SHAPE;
  if (c->FF <= 0 || (c->FF & 7)) return WM_ERR_SHAPE;
  if (c->out_dim <= 0 || c->out_dim > 64) return WM_ERR_SHAPE;
  if (c->dropout_p < 0.0f || c->dropout_p >= 1.0f) return WM_ERR_ARG;
  return WM_OK;
}

static inline uint64_t stream_id(uint64_t step, int layer, int site) {
  return (step << 16) | (static_cast<uint64_t>(layer) << 4) | static_cast<uint64_t>(site);
}

#define WM_TRY(expr)            \
  do {                          \
    const int rc__ = (expr);    \
    if (rc__ != WM_OK) return rc__; \
  } while (0)

}  // namespace wm

using namespace wm;

extern "C" {

int64_t wm_encoder_param_count(const wm_encoder_config* cfg) {
  if (check_cfg(cfg) != WM_OK) return -1;
  return make_layout(*cfg).total;
}

int wm_encoder_param_layout(const wm_encoder_config* cfg, int64_t* offsets, int max_n) {
  if (check_cfg(cfg) != WM_OK) return -1;
  const ParamLayout p = make_layout(*cfg);
  std::vector<int64_t> o;
  o.push_back(p.w_in);
  o.push_back(p.b_in);
  for (const LayerParams& q : p.layers) {
    const int64_t v[12] = {q.w_qkv, q.b_qkv, q.w_o, q.b_o, q.w1, q.b1, q.w2, q.b2, q.g1, q.be1, q.g2, q.be2};
    o.insert(o.end(), v, v + 12);
  }
  o.push_back(p.w_out);
  o.push_back(p.b_out);
  const int n = static_cast<int>(o.size());
  if (offsets) {
    if (max_n < n) return -1;
    for (int i = 0; i < n; ++i) offsets[i] = o[i];
  }
  return n;
}

size_t wm_encoder_workspace_bytes(const wm_encoder_config* cfg) {
  if (check_cfg(cfg) != WM_OK) return 0;
  wm_encoder tmp;
  tmp.cfg = *cfg;
  tmp.lay = make_layout(*cfg);
  tmp.M = static_cast<int64_t>(cfg->B) * cfg->S;
  tmp.outP = cfg->out_dim <= 32 ? 32 : 64;
  return carve(&tmp, nullptr) + 256;
}

int wm_encoder_create(const wm_encoder_config* cfg, void* workspace, size_t workspace_bytes, wm_encoder** out) {
  WM_TRY(check_cfg(cfg));
  if (!workspace || !out) return WM_ERR_ARG;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return WM_ERR_CUDA;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return WM_ERR_DEVICE;  // sm_100a only: no fallback path exists
  wm_encoder* e = new (std::nothrow) wm_encoder();
  if (!e) return WM_ERR_ARG;
  e->cfg = *cfg;
  e->lay = make_layout(*cfg);
  e->M = static_cast<int64_t>(cfg->B) * cfg->S;
  e->outP = cfg->out_dim <= 32 ? 32 : 64;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  const size_t need = carve(e, base);
  if (need + (base - reinterpret_cast<uint8_t*>(workspace)) > workspace_bytes) {
    delete e;
    return WM_ERR_SHAPE;
  }
  e->ws = base;
  e->ws_bytes = need;
  e->seed = 0;
  e->step = 0;
  e->training = 0;
  e->saved = 0;
  *out = e;
  return WM_OK;
}

int wm_encoder_destroy(wm_encoder* e) {
  delete e;
  return WM_OK;
}

// fp32 master -> bf16 shadow (same offsets) + transposed copies used as the B operand of dgrad GEMMs
static int refresh_impl(wm_encoder* e, const float* params, bool cast_shadow, cudaStream_t st) {
  const wm_encoder_config& c = e->cfg;
  if (cast_shadow) WM_TRY(launch_cast_bf16(params, e->shadow, e->lay.total, st));
  if (cudaMemsetAsync(e->wout_t, 0, static_cast<size_t>(c.D) * e->outP * 2, st) != cudaSuccess) return WM_ERR_CUDA;
  TransposeJobs jobs;
  jobs.n = 0;
  auto add = [&](const float* w, __nv_bfloat16* wt, int rows, int cols, int ld_out) {
    jobs.job[jobs.n++] = TransposeJob{w, wt, rows, cols, ld_out, 0};
  };
  for (int l = 0; l < c.L; ++l) {
    const LayerParams& q = e->lay.layers[l];
    if (jobs.n + 4 >= kMaxTransposeJobs) {  // (very deep models: flush in batches)
      WM_TRY(launch_cast_transpose_multi(jobs, st));
      jobs.n = 0;
    }
    add(params + q.w_qkv, e->wt[l].wqkv_t, 3 * c.D, c.D, 3 * c.D);
    add(params + q.w_o, e->wt[l].wo_t, c.D, c.D, c.D);
    add(params + q.w1, e->wt[l].w1_t, c.FF, c.D, c.FF);
    add(params + q.w2, e->wt[l].w2_t, c.D, c.FF, c.D);
  }
  add(params + e->lay.w_out, e->wout_t, c.out_dim, c.D, e->outP);
  return launch_cast_transpose_multi(jobs, st);
}

int wm_encoder_refresh_weights(wm_encoder* e, const float* params, void* stream_) {
  if (!e || !params) return WM_ERR_ARG;
  return refresh_impl(e, params, true, reinterpret_cast<cudaStream_t>(stream_));
}

// The same when the bf16 shadow is already current (wm_adam_fused / _dev wrote it through shadow_bf16 =
// wm_encoder_shadow()): only the transposed copies are rebuilt.
int wm_encoder_refresh_transposes(wm_encoder* e, const float* params, void* stream_) {
  if (!e || !params) return WM_ERR_ARG;
  return refresh_impl(e, params, false, reinterpret_cast<cudaStream_t>(stream_));
}

void* wm_encoder_shadow(wm_encoder* e) { return e ? e->shadow : nullptr; }

int wm_encoder_forward(wm_encoder* e, const float* params, const float* weather, const uint8_t* mask,
                       int64_t mask_stride_b, int64_t mask_stride_s, const float* year, const float* coords,
                       const float* pos_encoding, float* y_out, int training, int save_for_backward, uint64_t seed,
                       uint64_t step, void* stream_) {
  if (!e || !params || !weather || !mask || !year || !coords || !pos_encoding || !y_out) return WM_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const wm_encoder_config& c = e->cfg;
  if (save_for_backward && c.eval_only) return WM_ERR_ARG;  // an eval-only handle has nowhere to save activations
  const bool save = save_for_backward != 0;
  const int M = static_cast<int>(e->M), D = c.D, FF = c.FF, dh = c.D / c.H;
  const float p = training ? c.dropout_p : 0.0f;
  const uint32_t thr = static_cast<uint32_t>(p * 65536.0f + 0.5f);
  const float dscale = drop_keep_scale(thr);  // the kernels compare 15 bits: p_eff = round(32768 p) / 32768
  e->seed = seed;
  e->step = step;
  e->training = training;
  e->saved = save ? 1 : 0;

  // save: layer l keeps its own activations for the backward pass. Otherwise (validation / inference: the reference's
  // model.eval() + torch.no_grad(), src/base_trainer/base_trainer.py:262-285) every layer reuses layer 0's buffers, x
  // ping-pongs between act[0].x and x_final, and nothing that only backward reads is written (row statistics, lse,
  // sign bits, keep bits, the bf16 input copy).
  __nv_bfloat16* x_cur = e->act[0].x;
  __nv_bfloat16* x_alt = e->x_final;
  WM_TRY(launch_embed_fwd(weather, mask, mask_stride_b, mask_stride_s, year, coords, params + e->lay.w_in,
                          params + e->lay.b_in, pos_encoding, x_cur, save ? e->xin : nullptr, c.B, c.S, c.F, D, st));
  for (int l = 0; l < c.L; ++l) {
    const LayerParams& q = e->lay.layers[l];
    LayerAct& a = e->act[save ? l : 0];
    __nv_bfloat16* x_in = save ? a.x : x_cur;
    __nv_bfloat16* x_next = save ? (l + 1 < c.L ? e->act[l + 1].x : e->x_final) : x_alt;
    {  // QKV projection
      GemmEpilogue ep;
      ep.bias = params + q.b_qkv;
      ep.out = a.qkv;
      ep.ld_out = 3 * D;
      WM_TRY(launch_gemm_tn(x_in, D, e->shadow + q.w_qkv, D, M, 3 * D, D, ep, 0, 0, st));
    }
    WM_TRY(launch_attn_fwd(a.qkv, a.ctx, save ? a.lse : nullptr, save ? a.dropw : nullptr, c.B, c.S, c.H, dh, thr, seed,
                           stream_id(step, l, 0), st));
    {  // out-proj + dropout1 + residual
      GemmEpilogue ep;
      ep.bias = params + q.b_o;
      ep.drop_thresh = thr;
      ep.drop_scale = dscale;
      ep.seed = seed;
      ep.stream = stream_id(step, l, 1);
      ep.residual = x_in;
      ep.ld_res = D;
      ep.out = a.r1;
      ep.ld_out = D;
      WM_TRY(launch_gemm_tn(a.ctx, D, e->shadow + q.w_o, D, M, D, D, ep, 0, 0, st));
    }
    WM_TRY(launch_layernorm_fwd(a.r1, params + q.g1, params + q.be1, a.u, save ? a.mean1 : nullptr,
                                save ? a.rstd1 : nullptr, M, D, c.ln_eps, st));
    {  // linear1 + ReLU + dropout
      GemmEpilogue ep;
      ep.bias = params + q.b1;
      ep.relu = 1;
      ep.drop_thresh = thr;
      ep.drop_scale = dscale;
      ep.seed = seed;
      ep.stream = stream_id(step, l, 2);
      ep.sign_bits_out = save ? a.hbits : nullptr;
      ep.out = a.h;
      ep.ld_out = FF;
      WM_TRY(launch_gemm_tn(a.u, D, e->shadow + q.w1, D, M, FF, D, ep, 0, 0, st));
    }
    {  // linear2 + dropout2 + residual
      GemmEpilogue ep;
      ep.bias = params + q.b2;
      ep.drop_thresh = thr;
      ep.drop_scale = dscale;
      ep.seed = seed;
      ep.stream = stream_id(step, l, 3);
      ep.residual = a.u;
      ep.ld_res = D;
      ep.out = a.r2;
      ep.ld_out = D;
      WM_TRY(launch_gemm_tn(a.h, FF, e->shadow + q.w2, FF, M, D, FF, ep, 0, 0, st));
    }
    WM_TRY(launch_layernorm_fwd(a.r2, params + q.g2, params + q.be2, x_next, save ? a.mean2 : nullptr,
                                save ? a.rstd2 : nullptr, M, D, c.ln_eps, st));
    if (!save) {
      x_alt = x_cur;
      x_cur = x_next;
    }
  }
  const __nv_bfloat16* x_last = save ? e->x_final : x_cur;
  {  // output head -> fp32 [M, outP]; weight rows >= out_dim are TMA zero-fill, bias pad is the zero gap
    GemmEpilogue ep;
    ep.bias = params + e->lay.b_out;
    ep.out = y_out;
    ep.ld_out = e->outP;
    WM_TRY(launch_gemm_tn_rows(x_last, D, e->shadow + e->lay.w_out, D, M, e->outP, D, c.out_dim, ep, 1, st));
  }
  return WM_OK;
}

// gradient of the loss w.r.t. the padded head output (bf16 [M, outP]) -> head grads + grad of x_L
int wm_encoder_backward_head(wm_encoder* e, const void* dY_, float* grads, void* stream_) {
  if (!e || !dY_ || !grads) return WM_ERR_ARG;
  if (!e->saved) return WM_ERR_ARG;  // the last forward ran the lean (no-save) schedule
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const wm_encoder_config& c = e->cfg;
  const __nv_bfloat16* dY = reinterpret_cast<const __nv_bfloat16*>(dY_);
  const int M = static_cast<int>(e->M), D = c.D;
  WM_TRY(launch_gemm_wgrad_ex(dY, e->outP, e->x_final, D, M, e->outP, D, grads + e->lay.w_out, c.out_dim, D, D,
                              e->scratch, grads + e->lay.b_out, st));
  GemmEpilogue ep;
  ep.out = e->gX;
  ep.ld_out = D;
  WM_TRY(launch_gemm_tn(dY, e->outP, e->wout_t, e->outP, M, D, e->outP, ep, 0, 0, st));
  return WM_OK;
}

// backward through layers layer_hi-1 ... layer_lo (gX holds dLoss/dx_{layer_hi} on entry, dLoss/dx_{layer_lo} on exit)
int wm_encoder_backward_layers(wm_encoder* e, const float* params, int layer_hi, int layer_lo, float* grads,
                               void* stream_) {
  if (!e || !params || !grads || !e->saved) return WM_ERR_ARG;
  const wm_encoder_config& c = e->cfg;
  if (layer_lo < 0 || layer_hi > c.L || layer_lo > layer_hi) return WM_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int M = static_cast<int>(e->M), D = c.D, FF = c.FF, dh = c.D / c.H;
  const float p = e->training ? c.dropout_p : 0.0f;
  const uint32_t thr = static_cast<uint32_t>(p * 65536.0f + 0.5f);
  const float dscale = drop_keep_scale(thr);  // the kernels compare 15 bits: p_eff = round(32768 p) / 32768
  // d b2 / d b_o = column sums of the (dropout-masked) LayerNorm input gradients: taken from the wgrad GEMM of the
  // same layer when its tile leaves room for the fused all-ones chunk (then the LayerNorm backward kernel runs its
  // leaner 15-warp form), else accumulated by the LayerNorm backward kernel itself
  const bool b2_from_wgrad = wgrad_fuses_bias(M, D, FF), bo_from_wgrad = wgrad_fuses_bias(M, D, D);
  for (int l = layer_hi - 1; l >= layer_lo; --l) {
    const LayerParams& q = e->lay.layers[l];
    LayerAct& a = e->act[l];
    __nv_bfloat16* gFd = thr ? e->gF : e->gR;  // gradient seen by the linear output (after dropout mask)
    // LN2 backward: gX -> gR (d r2), gF (dropout2-masked), d gamma2 / d beta2
    WM_TRY(launch_layernorm_bwd(e->gX, a.r2, params + q.g2, a.mean2, a.rstd2, e->gR, thr ? e->gF : nullptr,
                                grads + q.g2, grads + q.be2, b2_from_wgrad ? nullptr : grads + q.b2, M, D, thr, dscale, e->seed,
                                stream_id(e->step, l, 3), e->scratch, st));
    // dW2 [D, FF] = gF^T h, d b2 = column sums of gF (fused all-ones chunk of the same GEMM)
    WM_TRY(launch_gemm_wgrad(gFd, D, a.h, FF, M, D, FF, grads + q.w2, 0, e->scratch, b2_from_wgrad ? grads + q.b2 : nullptr, st));
    {  // d h_pre = (gF W2) * [h > 0] / (1 - p)
      GemmEpilogue ep;
      ep.gate_bits = a.hbits;  // one bit per element instead of re-reading the 2-byte activation
      ep.gate_scale = dscale;
      ep.out = e->gH;
      ep.ld_out = FF;
      WM_TRY(launch_gemm_tn(gFd, D, e->wt[l].w2_t, D, M, FF, D, ep, 0, 0, st));
    }
    WM_TRY(launch_gemm_wgrad(e->gH, FF, a.u, D, M, FF, D, grads + q.w1, 0, e->scratch, grads + q.b1, st));
    {  // d u = gH W1 + d r2
      GemmEpilogue ep;
      ep.residual = e->gR;
      ep.ld_res = D;
      ep.out = e->gU;
      ep.ld_out = D;
      WM_TRY(launch_gemm_tn(e->gH, FF, e->wt[l].w1_t, FF, M, D, FF, ep, 0, 0, st));
    }
    // LN1 backward: gU -> gR (d r1), gF (dropout1-masked), d gamma1 / d beta1; d b_o from the wgrad below
    WM_TRY(launch_layernorm_bwd(e->gU, a.r1, params + q.g1, a.mean1, a.rstd1, e->gR, thr ? e->gF : nullptr,
                                grads + q.g1, grads + q.be1, bo_from_wgrad ? nullptr : grads + q.b_o, M, D, thr, dscale, e->seed,
                                stream_id(e->step, l, 1), e->scratch, st));
    WM_TRY(launch_gemm_wgrad(gFd, D, a.ctx, D, M, D, D, grads + q.w_o, 0, e->scratch, bo_from_wgrad ? grads + q.b_o : nullptr, st));
    {  // d ctx = gF W_o
      GemmEpilogue ep;
      ep.out = e->gC;
      ep.ld_out = D;
      WM_TRY(launch_gemm_tn(gFd, D, e->wt[l].wo_t, D, M, D, D, ep, 0, 0, st));
    }
    WM_TRY(launch_attn_bwd(a.qkv, a.ctx, e->gC, a.lse, e->gQKV, a.dropw, e->scratch, c.B, c.S, c.H, dh, thr, st));
    WM_TRY(launch_gemm_wgrad(e->gQKV, 3 * D, a.x, D, M, 3 * D, D, grads + q.w_qkv, 0, e->scratch, grads + q.b_qkv, st));
    {  // d x_l = gQKV W_qkv + d r1
      GemmEpilogue ep;
      ep.residual = e->gR;
      ep.ld_res = D;
      ep.out = e->gX;
      ep.ld_out = D;
      WM_TRY(launch_gemm_tn(e->gQKV, 3 * D, e->wt[l].wqkv_t, 3 * D, M, D, 3 * D, ep, 0, 0, st));
    }
  }
  return WM_OK;
}

// in_proj gradients from dLoss/dx_0 (left in gX by backward_layers(…, 0))
int wm_encoder_backward_embed(wm_encoder* e, float* grads, void* stream_) {
  if (!e || !grads || !e->saved) return WM_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const wm_encoder_config& c = e->cfg;
  const int M = static_cast<int>(e->M), D = c.D, Fin = c.F + 3;
  WM_TRY(launch_gemm_wgrad_ex(e->gX, D, e->xin, kXinCols, M, D, kXinCols, grads + e->lay.w_in, D, Fin, Fin,
                              e->scratch, grads + e->lay.b_in, st));
  return WM_OK;
}

// debugging / parity access to the saved activations: which = 0 x_l, 1 qkv, 2 ctx, 3 r1, 4 u, 5 h, 6 r2,
// 7 x_final, 8 gX
const void* wm_encoder_activation(wm_encoder* e, int layer, int which) {
  if (!e) return nullptr;
  if (which == 7) return e->x_final;
  if (which == 8) return e->gX;
  if (layer < 0 || layer >= static_cast<int>(e->act.size())) return nullptr;
  const LayerAct& a = e->act[layer];
  switch (which) {
    case 0: return a.x;
    case 1: return a.qkv;
    case 2: return a.ctx;
    case 3: return a.r1;
    case 4: return a.u;
    case 5: return a.h;
    case 6: return a.r2;
    default: return nullptr;
  }
}

}  // extern "C"
