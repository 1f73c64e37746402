// wm_yield.cu -- the crop-yield head as ONE forward and ONE backward kernel (BASELINE
// configs[5]). Replaces the ~40 eager launches per step of
//   WeatherBERTYieldModel._impute_weather / yield_model   src/crop_yield/models/weatherbert_yield_model.py:40-67
//   WeatherFormerYieldModel.forward (reparameterisation)   src/crop_yield/models/weatherformer_yield_model.py:58-60
// on the raw encoder head output:
//   z[s, f]   = mask ? (BERT: y[s, f] | Former: mu + sqrt(var) * eps,  var = clamp(exp(logvar), 1e-6, 1)) : weather[s, f]
//   score[s]  = W2 . gelu(W1 z[s] + b1) + b2             (31 -> 16 -> 1, exact erf GELU like nn.GELU())
//   a         = softmax over the sequence;  pooled = sum_s a[s] z[s, :]
//   pred      = W4 . gelu(W3 [pooled, y_past] + b3) + b4  (31 + n_past + 1 -> 120 -> 1)
// One CTA per sequence (B <= 64 in the reference's runs: a latency-bound launch either way), one thread per token,
// every reduction a fixed-order warp-shuffle tree + cross-warp fold: deterministic, as yield_main requires
// (torch.use_deterministic_algorithms(True), src/crop_yield/yield_main.py:124). fp32 throughout.
// The backward kernel recomputes the forward quantities from the saved z, writes dLoss/d(raw head output) in fp32 and
// per-sequence partial parameter gradients that a second tiny kernel folds over the batch in index order.
#include "wm_kernels.h"

namespace wm {

constexpr int kYhThreads = 384;  // >= S (S <= 384)
constexpr int kYhZLd = 33;       // smem row pitch of z (odd: conflict-free for thread = row)
constexpr int kYhHA = 16;        // attention hidden width (nn.Linear(weather_dim, 16))
constexpr int kYhHMax = 128;     // >= MLP hidden width (120)
constexpr int kYhUMax = 64;      // >= F + n_past + 1

WM_DEVICE float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
WM_DEVICE float gelu_erf_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// fixed-order block reductions (kYhThreads / 32 warps); `red` is a [12] float scratch in shared memory
WM_DEVICE float yh_block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
#pragma unroll
  for (int w = 0; w < kYhThreads / 32; ++w) t += red[w];
  return t;
}
WM_DEVICE float yh_block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
#pragma unroll
  for (int w = 0; w < kYhThreads / 32; ++w) t = fmaxf(t, red[w]);
  return t;
}

struct YieldHeadW {
  const float *w1, *b1, *w2, *b2, *w3, *b3, *w4, *b4;
};

// shared forward part: z row of this thread (registers + smem), attention scores, softmax, pooled, MLP hidden.
// Returns a[s]; leaves pooled / y_past in sU[0 .. F + np), h3 (pre-activation) in sH3[0 .. HM).
template <bool kFromSavedZ>
WM_DEVICE float yield_head_forward_part(const float* __restrict__ y, int ldy, int is_former,
                                        const float* __restrict__ weather, const uint8_t* __restrict__ mask, int64_t msb,
                                        int64_t mss, const float* __restrict__ eps, const float* __restrict__ z_saved,
                                        const float* __restrict__ y_past, int np, const YieldHeadW& W, int S, int F, int HM,
                                        float (&zr)[32], float (&h1)[kYhHA], float* sW1, float* sZ, float* sU, float* sH3,
                                        float* sPool, float* red, float* z_out) {
  const int b = blockIdx.x, s = threadIdx.x;
  const bool valid = s < S;
  for (int i = threadIdx.x; i < kYhHA * F; i += kYhThreads) sW1[i] = W.w1[i];
#pragma unroll
  for (int f = 0; f < 32; ++f) zr[f] = 0.0f;
  if (valid) {
    const size_t row = static_cast<size_t>(b) * S + s;
    if (kFromSavedZ) {
      for (int f = 0; f < F; ++f) zr[f] = z_saved[row * F + f];
    } else {
      const float* yr = y + row * ldy;
      const float* wr = weather + row * F;
      const uint8_t* mr = mask + b * msb + s * mss;
#pragma unroll
      for (int f = 0; f < 32; ++f) {
        if (f < F) {
          float v = wr[f];
          if (mr[f]) {
            v = yr[f];
            if (is_former) {
              const float var = fminf(fmaxf(expf(yr[F + f]), 1e-6f), 1.0f);
              v = v + sqrtf(var) * eps[row * F + f];
            }
          }
          zr[f] = v;
        }
      }
      if (z_out)
        for (int f = 0; f < F; ++f) z_out[row * F + f] = zr[f];
    }
  }
#pragma unroll
  for (int f = 0; f < 32; ++f) sZ[s * kYhZLd + f] = zr[f];
  __syncthreads();
  float sc = W.b2[0];
#pragma unroll
  for (int j = 0; j < kYhHA; ++j) {
    float acc = W.b1[j];
#pragma unroll
    for (int f = 0; f < 32; ++f)
      if (f < F) acc = fmaf(sW1[j * F + f], zr[f], acc);
    h1[j] = acc;
    sc = fmaf(W.w2[j], gelu_erf(acc), sc);
  }
  const float mx = yh_block_max(valid ? sc : -INFINITY, red);
  const float e = valid ? expf(sc - mx) : 0.0f;
  const float tot = yh_block_sum(e, red);
  const float a = e / tot;
  // pooled[f] = sum_s a[s] z[s, f]: warp tree per feature, then a fold over the 12 warps in index order
  __syncthreads();
#pragma unroll
  for (int f = 0; f < 32; ++f) {
    const float v = warp_sum(a * zr[f]);
    if ((threadIdx.x & 31) == 0) sPool[(threadIdx.x >> 5) * 32 + f] = v;
  }
  __syncthreads();
  if (threadIdx.x < F) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < kYhThreads / 32; ++w) t += sPool[w * 32 + threadIdx.x];
    sU[threadIdx.x] = t;
  } else if (threadIdx.x < F + np) {
    sU[threadIdx.x] = y_past[static_cast<size_t>(b) * np + (threadIdx.x - F)];
  }
  __syncthreads();
  if (threadIdx.x < HM) {
    const int nu = F + np;
    float acc = W.b3[threadIdx.x];
    for (int i = 0; i < nu; ++i) acc = fmaf(W.w3[threadIdx.x * nu + i], sU[i], acc);
    sH3[threadIdx.x] = acc;
  }
  __syncthreads();
  return a;
}

struct YieldSmem {
  float w1[kYhHA * 32];
  float z[kYhThreads * kYhZLd];
  float u[kYhUMax];
  float h3[kYhHMax];
  float pool[(kYhThreads / 32) * 32];
  float red[16];
};

__global__ void __launch_bounds__(kYhThreads)
yield_head_fwd_kernel(const float* __restrict__ y, int ldy, int is_former, const float* __restrict__ weather,
                      const uint8_t* __restrict__ mask, int64_t msb, int64_t mss, const float* __restrict__ eps,
                      const float* __restrict__ y_past, int np, YieldHeadW W, float* __restrict__ z_out,
                      float* __restrict__ pred, int S, int F, int HM) {
  extern __shared__ __align__(16) uint8_t yh_smem_raw[];
  YieldSmem& sm = *reinterpret_cast<YieldSmem*>(yh_smem_raw);
  float zr[32], h1[kYhHA];
  yield_head_forward_part<false>(y, ldy, is_former, weather, mask, msb, mss, eps, nullptr, y_past, np, W, S, F, HM, zr, h1,
                                 sm.w1, sm.z, sm.u, sm.h3, sm.pool, sm.red, z_out);
  const float part = threadIdx.x < HM ? W.w4[threadIdx.x] * gelu_erf(sm.h3[threadIdx.x]) : 0.0f;
  const float p = yh_block_sum(part, sm.red);
  if (threadIdx.x == 0) pred[blockIdx.x] = p + W.b4[0];
}

// flat layout of the head's parameter gradients (and of each sequence's partial): see wm_yield_head_param_count
struct YieldGradLayout {
  int w1, b1, w2, b2, w3, b3, w4, b4, total;
};
__host__ __device__ inline YieldGradLayout yield_grad_layout(int F, int np, int HM) {
  YieldGradLayout L;
  int o = 0;
  L.w1 = o; o += kYhHA * F;
  L.b1 = o; o += kYhHA;
  L.w2 = o; o += kYhHA;
  L.b2 = o; o += 1;
  L.w3 = o; o += HM * (F + np);
  L.b3 = o; o += HM;
  L.w4 = o; o += HM;
  L.b4 = o; o += 1;
  L.total = o;
  return L;
}

struct YieldBwdSmem {
  YieldSmem f;
  float dh1[kYhThreads * (kYhHA + 1)];  // d h1[s, j]
  float t2[kYhThreads * (kYhHA + 1)];   // dsc[s] * gelu(h1[s, j])
  float dh3[kYhHMax];
  float dpool[32];
};

__global__ void __launch_bounds__(kYhThreads)
yield_head_bwd_kernel(const float* __restrict__ dpred, const float* __restrict__ y, int ldy, int is_former,
                      const uint8_t* __restrict__ mask, int64_t msb, int64_t mss, const float* __restrict__ eps,
                      const float* __restrict__ z_saved, const float* __restrict__ y_past, int np, YieldHeadW W,
                      float* __restrict__ dy, float* __restrict__ partial, int S, int F, int HM) {
  extern __shared__ __align__(16) uint8_t yh_smem_raw[];
  YieldBwdSmem& sm = *reinterpret_cast<YieldBwdSmem*>(yh_smem_raw);
  const int b = blockIdx.x, s = threadIdx.x, tid = threadIdx.x;
  const bool valid = s < S;
  const YieldGradLayout L = yield_grad_layout(F, np, HM);
  float* gp = partial + static_cast<size_t>(b) * L.total;
  float zr[32], h1[kYhHA];
  const float a = yield_head_forward_part<true>(y, ldy, is_former, nullptr, mask, msb, mss, eps, z_saved, y_past, np, W, S, F,
                                                HM, zr, h1, sm.f.w1, sm.f.z, sm.f.u, sm.f.h3, sm.f.pool, sm.f.red, nullptr);
  const float dp = dpred[b];
  const int nu = F + np;
  // ---- MLP: thread j owns hidden unit j
  if (tid < HM) {
    const float h3 = sm.f.h3[tid];
    const float dh3 = dp * W.w4[tid] * gelu_erf_grad(h3);
    sm.dh3[tid] = dh3;
    gp[L.w4 + tid] = dp * gelu_erf(h3);
    gp[L.b3 + tid] = dh3;
    for (int i = 0; i < nu; ++i) gp[L.w3 + tid * nu + i] = dh3 * sm.f.u[i];
  }
  if (tid == 0) gp[L.b4] = dp;
  __syncthreads();
  if (tid < F) {  // d pooled[f] = sum_j W3[j, f] dh3[j]
    float t = 0.0f;
    for (int j = 0; j < HM; ++j) t = fmaf(W.w3[j * nu + tid], sm.dh3[j], t);
    sm.dpool[tid] = t;
  }
  __syncthreads();
  // ---- softmax pooling backward, thread = token
  float da = 0.0f;
#pragma unroll
  for (int f = 0; f < 32; ++f)
    if (f < F) da = fmaf(sm.dpool[f], zr[f], da);
  const float dot = yh_block_sum(a * da, sm.f.red);
  const float dsc = valid ? a * (da - dot) : 0.0f;
  const float db2 = yh_block_sum(dsc, sm.f.red);
  if (tid == 0) gp[L.b2] = db2;
  float dz[32];
#pragma unroll
  for (int f = 0; f < 32; ++f) dz[f] = f < F ? a * sm.dpool[f] : 0.0f;
#pragma unroll
  for (int j = 0; j < kYhHA; ++j) {
    const float d1 = dsc * W.w2[j] * gelu_erf_grad(h1[j]);
    sm.dh1[s * (kYhHA + 1) + j] = d1;
    sm.t2[s * (kYhHA + 1) + j] = dsc * gelu_erf(h1[j]);
#pragma unroll
    for (int f = 0; f < 32; ++f)
      if (f < F) dz[f] = fmaf(d1, sm.f.w1[j * F + f], dz[f]);
  }
  __syncthreads();
  // ---- reductions over the sequence (index order): dW1[j, f], db1[j], dW2[j]
  for (int o = tid; o < kYhHA * F; o += kYhThreads) {
    const int j = o / F, f = o - j * F;
    float t = 0.0f;
    for (int q = 0; q < S; ++q) t = fmaf(sm.dh1[q * (kYhHA + 1) + j], sm.f.z[q * kYhZLd + f], t);
    gp[L.w1 + o] = t;
  }
  if (tid < 2 * kYhHA) {
    const int j = tid & (kYhHA - 1);
    const float* src = tid < kYhHA ? sm.dh1 : sm.t2;
    float t = 0.0f;
    for (int q = 0; q < S; ++q) t += src[q * (kYhHA + 1) + j];
    gp[(tid < kYhHA ? L.b1 : L.w2) + j] = t;
  }
  // ---- d(raw head output): through the imputation (and the reparameterisation for WeatherFormer)
  if (valid) {
    const size_t row = static_cast<size_t>(b) * S + s;
    float* dyr = dy + row * ldy;
    const uint8_t* mr = mask + b * msb + s * mss;
    const float* yr = y + row * ldy;
#pragma unroll
    for (int f = 0; f < 32; ++f) {
      if (f < F) {
        const float g = mr[f] ? dz[f] : 0.0f;
        dyr[f] = g;
        if (is_former) {
          const float e = expf(yr[F + f]);
          const float var = fminf(fmaxf(e, 1e-6f), 1.0f);
          // z = mu + sqrt(var) eps: d/dvar = eps / (2 sqrt(var)); clamp passes the gradient on [1e-6, 1]; dvar/dlogvar = e
          dyr[F + f] = (e >= 1e-6f && e <= 1.0f) ? g * eps[row * F + f] * 0.5f * rsqrtf(var) * e : 0.0f;
        }
      }
    }
    for (int c = (is_former ? 2 * F : F); c < ldy; ++c) dyr[c] = 0.0f;
  }
}

// grads[i] = sum_b partial[b][i] in index order (deterministic)
__global__ void yield_head_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grads, int total, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float t = 0.0f;
  for (int b = 0; b < B; ++b) t += partial[static_cast<size_t>(b) * total + i];
  grads[i] = t;
}

static bool yield_shape_ok(int B, int S, int F, int np, int HM, int ldy, int is_former) {
  return B > 0 && S > 0 && S <= kYhThreads && F > 0 && F <= 32 && np >= 0 && F + np <= kYhUMax && HM > 0 && HM <= kYhHMax &&
         ldy >= (is_former ? 2 * F : F);
}

int yield_head_param_count(int F, int np, int HM) { return yield_grad_layout(F, np, HM).total; }

int launch_yield_head_fwd(const float* y, int ldy, int is_former, const float* weather, const uint8_t* mask, int64_t msb,
                          int64_t mss, const float* eps, const float* y_past, int np, const YieldHeadW& W, float* z_out,
                          float* pred, int B, int S, int F, int HM, cudaStream_t stream) {
  if (!yield_shape_ok(B, S, F, np, HM, ldy, is_former)) return WM_ERR_SHAPE;
  if (!y || !weather || !mask || !pred || (is_former && !eps) || (np > 0 && !y_past)) return WM_ERR_ARG;
  const int smem = static_cast<int>(sizeof(YieldSmem));
  if (cudaFuncSetAttribute(yield_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return WM_ERR_CUDA;
  yield_head_fwd_kernel<<<B, kYhThreads, smem, stream>>>(y, ldy, is_former, weather, mask, msb, mss, eps, y_past, np, W, z_out,
                                                         pred, S, F, HM);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

int launch_yield_head_bwd(const float* dpred, const float* y, int ldy, int is_former, const uint8_t* mask, int64_t msb,
                          int64_t mss, const float* eps, const float* z_saved, const float* y_past, int np,
                          const YieldHeadW& W, float* dy, float* partial, float* grads, int B, int S, int F, int HM,
                          cudaStream_t stream) {
  if (!yield_shape_ok(B, S, F, np, HM, ldy, is_former)) return WM_ERR_SHAPE;
  if (!dpred || !y || !mask || !z_saved || !dy || !partial || !grads || (is_former && !eps) || (np > 0 && !y_past)) return WM_ERR_ARG;
  const int smem = static_cast<int>(sizeof(YieldBwdSmem));
  if (cudaFuncSetAttribute(yield_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return WM_ERR_CUDA;
  yield_head_bwd_kernel<<<B, kYhThreads, smem, stream>>>(dpred, y, ldy, is_former, mask, msb, mss, eps, z_saved, y_past, np, W,
                                                         dy, partial, S, F, HM);
  WM_COUNT_LAUNCH();
  const int total = yield_grad_layout(F, np, HM).total;
  yield_head_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(partial, grads, total, B);
  WM_COUNT_LAUNCH();
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

}  // namespace wm
