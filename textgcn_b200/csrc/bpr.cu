// Fused BPR step: user/pos/neg gathers, dot-product scores, SELU loss (the reference uses SELU, not
// softplus: base_model.py:194), L2 regulariser on the layer-0 rows (:200-210), and the gradients
// scatter-added into dense (N, d) buffers with 128-bit vector atomics (red.global.add.v4.f32).
//
// Roofline: HBM/L2 gather bound, batch·(2 + n_neg)·2·4d bytes read, the same again in atomics.
// One warp per batch row.  Loss terms are written per warp and summed in a fixed order by a second
// one-block kernel, so the reported loss is deterministic; the gradient scatter uses float atomics
// (order-dependent in the last bit when a row repeats in the batch; inside the 1e-5 parity budget).
#include "common.cuh"

namespace tgcn {

constexpr float kSeluScale = 1.0507009873554804934193349852946f;
constexpr float kSeluNegCoef = (float)(1.6732632423543772848170429916717 * 1.0507009873554804934193349852946);

struct BprArgs {
  int n_users, n_items, d, batch, n_neg;
  const int* users;
  const int* pos;
  const int* negs;  // (n_neg, batch)
  const float* emb;  // (N, d)
  const float* user_w;
  const float* item_w;
  float reg_coef;   // reg_lambda / batch  (gradient of reg_lambda/(2·batch)·‖row‖² is reg_coef·row)
  float grad_coef;  // 1 / (batch · n_neg)
  float* grad_emb;
  float* grad_w0;
  float2* partials;  // per warp: (Σ selu, Σ ‖row‖²)
};

__device__ __forceinline__ void atomic_add4(float* p, const float4& v) {
  atomicAdd(reinterpret_cast<float4*>(p), v);
}

constexpr int kMaxChunks = 4;  // d <= 512

__global__ void __launch_bounds__(256) bpr_kernel(const BprArgs a) {
  const int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= a.batch) return;
  const int d4 = a.d >> 2;
  const int u = __ldg(a.users + warp);
  const int p = __ldg(a.pos + warp);
  // Rows the device sampler could not complete carry -1 (tgcn_sample_bpr_batch: a user with no train item, or no
  // non-positive item left): such a row — or any id outside the tables — contributes nothing and touches no memory.
  if ((unsigned)u >= (unsigned)a.n_users || (unsigned)p >= (unsigned)a.n_items) {
    if (lane == 0) a.partials[warp] = make_float2(0.f, 0.f);
    return;
  }
  const size_t d = a.d;
  const float* eu_p = a.emb + (size_t)u * d;
  const float* ep_p = a.emb + (size_t)(a.n_users + p) * d;
  float4 eu[kMaxChunks], ep[kMaxChunks], gu[kMaxChunks], gp[kMaxChunks];
  float pos_part = 0.f, reg_part = 0.f;
#pragma unroll
  for (int w = 0; w < kMaxChunks; ++w) {
    const int chunk = lane + 32 * w;
    gu[w] = gp[w] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (chunk < d4) {
      eu[w] = ldg4(eu_p + chunk * 4);
      ep[w] = ldg4(ep_p + chunk * 4);
      pos_part += dot4(eu[w], ep[w]);
      const float4 wu = ldg4(a.user_w + (size_t)u * d + chunk * 4);
      const float4 wp = ldg4(a.item_w + (size_t)p * d + chunk * 4);
      reg_part += dot4(wu, wu) + dot4(wp, wp);
      if (a.grad_w0) {
        atomic_add4(a.grad_w0 + (size_t)u * d + chunk * 4, make_float4(a.reg_coef * wu.x, a.reg_coef * wu.y, a.reg_coef * wu.z, a.reg_coef * wu.w));
        atomic_add4(a.grad_w0 + (size_t)(a.n_users + p) * d + chunk * 4, make_float4(a.reg_coef * wp.x, a.reg_coef * wp.y, a.reg_coef * wp.z, a.reg_coef * wp.w));
      }
    } else {
      eu[w] = ep[w] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float pos_score = warp_sum(pos_part);
  float loss_sum = 0.f;
  for (int j = 0; j < a.n_neg; ++j) {
    const int n = __ldg(a.negs + (size_t)j * a.batch + warp);
    if ((unsigned)n >= (unsigned)a.n_items) continue;  // sentinel negative: skipped (warp-uniform)
    const float* en_p = a.emb + (size_t)(a.n_users + n) * d;
    float4 en[kMaxChunks];
    float neg_part = 0.f;
#pragma unroll
    for (int w = 0; w < kMaxChunks; ++w) {
      const int chunk = lane + 32 * w;
      if (chunk < d4) {
        en[w] = ldg4(en_p + chunk * 4);
        neg_part += dot4(eu[w], en[w]);
        const float4 wn = ldg4(a.item_w + (size_t)n * d + chunk * 4);
        reg_part += dot4(wn, wn);
        if (a.grad_w0)
          atomic_add4(a.grad_w0 + (size_t)(a.n_users + n) * d + chunk * 4, make_float4(a.reg_coef * wn.x, a.reg_coef * wn.y, a.reg_coef * wn.z, a.reg_coef * wn.w));
      } else {
        en[w] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float x = warp_sum(neg_part) - pos_score;
    // SELU as torch evaluates it: x > 0 ? scale·x : (exp(x) - 1)·(alpha·scale); derivative uses the x <= 0 branch at 0
    const float ex = expf(x);
    loss_sum += x > 0.f ? kSeluScale * x : (ex - 1.f) * kSeluNegCoef;
    const float g = (x > 0.f ? kSeluScale : kSeluNegCoef * ex) * a.grad_coef;
    if (a.grad_emb) {
#pragma unroll
      for (int w = 0; w < kMaxChunks; ++w) {
        const int chunk = lane + 32 * w;
        if (chunk < d4) {
          gu[w].x += g * (en[w].x - ep[w].x);
          gu[w].y += g * (en[w].y - ep[w].y);
          gu[w].z += g * (en[w].z - ep[w].z);
          gu[w].w += g * (en[w].w - ep[w].w);
          fma4(gp[w], -g, eu[w]);
          atomic_add4(a.grad_emb + (size_t)(a.n_users + n) * d + chunk * 4, make_float4(g * eu[w].x, g * eu[w].y, g * eu[w].z, g * eu[w].w));
        }
      }
    }
  }
  if (a.grad_emb) {
#pragma unroll
    for (int w = 0; w < kMaxChunks; ++w) {
      const int chunk = lane + 32 * w;
      if (chunk < d4) {
        atomic_add4(a.grad_emb + (size_t)u * d + chunk * 4, gu[w]);
        atomic_add4(a.grad_emb + (size_t)(a.n_users + p) * d + chunk * 4, gp[w]);
      }
    }
  }
  const float reg_total = warp_sum(reg_part);
  if (lane == 0) a.partials[warp] = make_float2(loss_sum, reg_total);
}

// Fixed-order reduction of the per-row partials (double accumulation), one block.
__global__ void __launch_bounds__(1024) bpr_finalize_kernel(const float2* __restrict__ partials, int batch, int n_neg,
                                                          float reg_lambda, float* __restrict__ losses) {
  __shared__ double s_l[1024], s_r[1024];
  double l = 0.0, r = 0.0;
  for (int i = threadIdx.x; i < batch; i += 1024) {
    l += (double)partials[i].x;
    r += (double)partials[i].y;
  }
  s_l[threadIdx.x] = l;
  s_r[threadIdx.x] = r;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_l[threadIdx.x] += s_l[threadIdx.x + o];
      s_r[threadIdx.x] += s_r[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    losses[0] = (float)(s_l[0] / ((double)batch * (double)n_neg));
    losses[1] = (float)((double)reg_lambda * s_r[0] / (double)batch / 2.0);
  }
}

// Dense Adam step, torch.optim.Adam defaults (no amsgrad, no weight decay), float4-vectorised.
// bc != NULL: the bias corrections come from device memory (written by adam_prepare_kernel), so the launch carries no
// per-step host scalar and the whole training step can be replayed from a CUDA graph.
__global__ void __launch_bounds__(256) adam_kernel(int64_t n4, float4* __restrict__ p, const float4* __restrict__ g,
                                                   float4* __restrict__ m, float4* __restrict__ v, float lr, float b1, float b2,
                                                   float eps, float bc1, float bc2_sqrt, const float* __restrict__ bc) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  if (bc != nullptr) {
    bc1 = __ldg(bc);
    bc2_sqrt = __ldg(bc + 1);
  }
  float4 pi = p[i], mi = m[i], vi = v[i];
  const float4 gi = g[i];
  const float step = lr / bc1;
#define TGCN_ADAM(c)                                   \
  mi.c = mi.c + (gi.c - mi.c) * (1.f - b1);            \
  vi.c = vi.c * b2 + gi.c * gi.c * (1.f - b2);         \
  pi.c = pi.c - step * (mi.c / (sqrtf(vi.c) / bc2_sqrt + eps));
  TGCN_ADAM(x) TGCN_ADAM(y) TGCN_ADAM(z) TGCN_ADAM(w)
#undef TGCN_ADAM
  p[i] = pi;
  m[i] = mi;
  v[i] = vi;
}

// ++*step, then the bias corrections of that step: bc[0] = 1 - b1^step, bc[1] = sqrt(1 - b2^step) (one thread, fp64)
__global__ void adam_prepare_kernel(long long* __restrict__ step, float* __restrict__ bc, double b1, double b2) {
  const long long t = *step + 1;
  *step = t;
  bc[0] = (float)(1.0 - pow(b1, (double)t));
  bc[1] = (float)sqrt(1.0 - pow(b2, (double)t));
}

}  // namespace tgcn

using namespace tgcn;

extern "C" {

int64_t tgcn_bpr_workspace_bytes(int64_t batch) { return batch > 0 ? batch * (int64_t)sizeof(float2) + 256 : 256; }

int tgcn_bpr_fwd_bwd(int64_t n_users, int64_t n_items, int64_t d, int64_t batch, int32_t n_neg, const int32_t* d_users,
                     const int32_t* d_pos, const int32_t* d_negs, const float* d_emb, const float* d_user_w,
                     const float* d_item_w, float reg_lambda, float* d_losses, float* d_grad_emb, float* d_grad_w0,
                     void* d_workspace, int64_t workspace_bytes, tgcn_stream_t stream) {
  TGCN_REQUIRE(n_users > 0 && n_items > 0 && batch > 0 && n_neg > 0, "bad sizes: batch=%lld n_neg=%d", (long long)batch, n_neg);
  TGCN_REQUIRE(d > 0 && d % 4 == 0 && d <= 128 * kMaxChunks, "embedding width d=%lld must be a multiple of 4 and <= %d", (long long)d, 128 * kMaxChunks);
  TGCN_REQUIRE(d_users && d_pos && d_negs && d_emb && d_user_w && d_item_w && d_losses, "NULL argument");
  TGCN_REQUIRE(d_workspace && workspace_bytes >= tgcn_bpr_workspace_bytes(batch), "workspace too small");
  TGCN_REQUIRE(batch < (1ll << 26), "batch too large");
  BprArgs a;
  a.n_users = (int)n_users;
  a.n_items = (int)n_items;
  a.d = (int)d;
  a.batch = (int)batch;
  a.n_neg = n_neg;
  a.users = d_users;
  a.pos = d_pos;
  a.negs = d_negs;
  a.emb = d_emb;
  a.user_w = d_user_w;
  a.item_w = d_item_w;
  a.reg_coef = (float)((double)reg_lambda / (double)batch);
  a.grad_coef = (float)(1.0 / ((double)batch * (double)n_neg));
  a.grad_emb = d_grad_emb;
  a.grad_w0 = d_grad_w0;
  a.partials = (float2*)d_workspace;
  cudaStream_t s = (cudaStream_t)stream;
  const int threads = 256;
  const int64_t blocks = (batch * 32 + threads - 1) / threads;
  bpr_kernel<<<(unsigned)blocks, threads, 0, s>>>(a);
  TGCN_CHECK_LAUNCH();
  bpr_finalize_kernel<<<1, 1024, 0, s>>>(a.partials, a.batch, n_neg, reg_lambda, d_losses);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int tgcn_adam_step(int64_t n, float* d_p, const float* d_g, float* d_m, float* d_v, float lr, float beta1, float beta2,
                   float eps, int64_t step, tgcn_stream_t stream) {
  TGCN_REQUIRE(n > 0 && n % 4 == 0, "n=%lld must be a positive multiple of 4", (long long)n);
  TGCN_REQUIRE(d_p && d_g && d_m && d_v && step >= 1, "bad argument");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const int64_t n4 = n / 4;
  const int threads = 256;
  const int64_t blocks = (n4 + threads - 1) / threads;
  adam_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(n4, (float4*)d_p, (const float4*)d_g, (float4*)d_m, (float4*)d_v,
                                                                     lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), nullptr);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int tgcn_adam_prepare(int64_t* d_step, float* d_bc, float beta1, float beta2, tgcn_stream_t stream) {
  TGCN_REQUIRE(d_step && d_bc, "NULL argument");
  adam_prepare_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long*)d_step, d_bc, (double)beta1, (double)beta2);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int tgcn_adam_step_dev(int64_t n, float* d_p, const float* d_g, float* d_m, float* d_v, float lr, float beta1, float beta2,
                       float eps, const float* d_bc, tgcn_stream_t stream) {
  TGCN_REQUIRE(n > 0 && n % 4 == 0, "n=%lld must be a positive multiple of 4", (long long)n);
  TGCN_REQUIRE(d_p && d_g && d_m && d_v && d_bc, "bad argument");
  const int64_t n4 = n / 4;
  const int threads = 256;
  const int64_t blocks = (n4 + threads - 1) / threads;
  adam_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(n4, (float4*)d_p, (const float4*)d_g, (float4*)d_m, (float4*)d_v,
                                                                     lr, beta1, beta2, eps, 1.f, 1.f, d_bc);
  TGCN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
