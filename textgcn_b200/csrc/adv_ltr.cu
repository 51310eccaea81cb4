// (d) AdvSamplModel hardest-negative selection and (e) LTR feature assembly.
//
// adv_select_kernel: one CTA per batch row.  Scores the row's candidates against the user (gather of
// n_cand·4d bytes: HBM/L2 bound), bitonic-sorts (score desc, candidate position asc) in shared memory,
// drops the user's train items and writes the first kmax survivors (advanced_sampling.py:61-65).
// ltr_*: gathers + row-wise dot products of the text tables (ltr_models.py:148-166) and the packing that
// turns score_batchwise_ltr into a single contraction for the fused eval kernel.
#include <limits.h>
#include <math.h>

#include "common.cuh"

namespace tgcn {

constexpr int kAdvThreads = 256;
constexpr int kMaxChunks = 4;  // d <= 512 with 32 lanes

struct AdvArgs {
  int n_users, n_items, d, batch, n_cand, p2, kmax;
  const int* users;
  const int* cands;  // (batch, n_cand) item ids
  const float* emb;  // (N, d)
  const int* mrowptr;
  const int* mcol;
  int* out_negs;
  int* out_counts;
  float* out_scores;  // optional (batch, n_cand)
};

__global__ void __launch_bounds__(kAdvThreads) adv_select_kernel(const AdvArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_score = reinterpret_cast<float*>(smem_raw);  // [p2]
  int* s_pos = reinterpret_cast<int*>(s_score + a.p2);   // [p2]
  int* s_scan = s_pos + a.p2;                            // [kAdvThreads]
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u = __ldg(a.users + b);
  TGCN_DASSERT(u >= 0 && u < a.n_users);
  const int d4 = a.d >> 2;
  // lanes per candidate: smallest power of two >= d/4, capped at 32
  int lpn = 1;
  while (lpn < d4 && lpn < 32) lpn <<= 1;
  const int grp = lane / lpn, sub = lane % lpn, groups = 32 / lpn;
  float4 ue[kMaxChunks];
#pragma unroll
  for (int w = 0; w < kMaxChunks; ++w) {
    const int chunk = sub + w * lpn;
    ue[w] = chunk < d4 ? ldg4(a.emb + (size_t)u * a.d + chunk * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int* cand = a.cands + (size_t)b * a.n_cand;
  const int per_warp_step = groups;
  for (int c0 = warp * per_warp_step; c0 < a.p2; c0 += (kAdvThreads / 32) * per_warp_step) {
    const int c = c0 + grp;
    float part = 0.f;
    if (c < a.n_cand) {
      const int item = __ldg(cand + c);
      TGCN_DASSERT(item >= 0 && item < a.n_items);
      const float* ip = a.emb + (size_t)(a.n_users + item) * a.d;
#pragma unroll
      for (int w = 0; w < kMaxChunks; ++w) {
        const int chunk = sub + w * lpn;
        if (chunk < d4) part += dot4(ue[w], ldg4(ip + chunk * 4));
      }
    }
    for (int o = lpn >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (sub == 0 && c < a.p2) {
      const bool real = c < a.n_cand;
      s_score[c] = real ? part : -INFINITY;
      s_pos[c] = real ? c : INT_MAX;
      if (real && a.out_scores) a.out_scores[(size_t)b * a.n_cand + c] = part;
    }
  }
  __syncthreads();
  // bitonic sort, "ranks_before" order ascending in index
  for (int size = 2; size <= a.p2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (a.p2 >> 1); t += kAdvThreads) {
        const int i = 2 * t - (t & (stride - 1));
        const int j = i + stride;
        const bool up = (i & size) == 0;  // this block sorts best-first
        const float si = s_score[i], sj = s_score[j];
        const int pi = s_pos[i], pj = s_pos[j];
        const bool j_first = ranks_before(sj, pj, si, pi);
        if (j_first == up) {
          s_score[i] = sj;
          s_score[j] = si;
          s_pos[i] = pj;
          s_pos[j] = pi;
        }
      }
      __syncthreads();
    }
  }
  // order-preserving removal of the user's train items; keep the first kmax
  const int per = (a.p2 + kAdvThreads - 1) / kAdvThreads;
  const int lo = tid * per, hi = min(a.p2, lo + per);
  const int mlo = a.mrowptr ? __ldg(a.mrowptr + u) : 0, mhi = a.mrowptr ? __ldg(a.mrowptr + u + 1) : 0;
  unsigned keep_bits = 0;
  int cnt = 0;
  for (int t = lo; t < hi; ++t) {
    const int p = s_pos[t];
    bool keep = p != INT_MAX;
    if (keep && a.mrowptr) keep = !sorted_contains(a.mcol, mlo, mhi, __ldg(cand + p) + a.n_users);
    if (keep) {
      keep_bits |= 1u << (t - lo);
      ++cnt;
    }
  }
  s_scan[tid] = cnt;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int i = 0; i < kAdvThreads; ++i) {
      const int c = s_scan[i];
      s_scan[i] = run;
      run += c;
    }
    a.out_counts[b] = min(run, a.kmax);
  }
  __syncthreads();
  int o = s_scan[tid];
  for (int t = lo; t < hi && o < a.kmax; ++t) {
    if (keep_bits & (1u << (t - lo))) {
      a.out_negs[(size_t)b * a.kmax + o] = __ldg(cand + s_pos[t]);
      ++o;
    }
  }
  __syncthreads();
  const int total = a.out_counts[b];
  for (int r = total + tid; r < a.kmax; r += kAdvThreads) a.out_negs[(size_t)b * a.kmax + r] = -1;
}

struct LtrPairArgs {
  int n_users, d, D, batch, n_feat;
  const int* users;
  const int* items;
  const float* emb;
  const float *users_rev, *users_desc, *items_rev, *items_desc, *pop_users, *pop_items;
  float* out;
};

__global__ void __launch_bounds__(256) ltr_pairwise_kernel(const LtrPairArgs a) {
  const int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= a.batch) return;
  const int u = __ldg(a.users + warp), it = __ldg(a.items + warp);
  TGCN_DASSERT(u >= 0 && u < a.n_users && it >= 0);
  float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f, f4 = 0.f;
  const float* ue = a.emb + (size_t)u * a.d;
  const float* ie = a.emb + (size_t)(a.n_users + it) * a.d;
  for (int c = lane; c < (a.d >> 2); c += 32) f0 += dot4(ldg4(ue + c * 4), ldg4(ie + c * 4));
  const float* ur = a.users_rev + (size_t)u * a.D;
  const float* ud = a.users_desc + (size_t)u * a.D;
  const float* ir = a.items_rev + (size_t)it * a.D;
  const float* id = a.items_desc + (size_t)it * a.D;
  for (int c = lane; c < (a.D >> 2); c += 32) {
    const float4 vur = ldg4(ur + c * 4), vud = ldg4(ud + c * 4), vir = ldg4(ir + c * 4), vid = ldg4(id + c * 4);
    f1 += dot4(vur, vir);
    f2 += dot4(vud, vid);
    f3 += dot4(vur, vid);
    f4 += dot4(vud, vir);
  }
  f0 = warp_sum(f0);
  f1 = warp_sum(f1);
  f2 = warp_sum(f2);
  f3 = warp_sum(f3);
  f4 = warp_sum(f4);
  if (lane == 0) {
    float* o = a.out + (size_t)warp * a.n_feat;
    o[0] = f0;
    o[1] = f1;
    o[2] = f2;
    o[3] = f3;
    o[4] = f4;
    if (a.n_feat == 7) {
      o[5] = __ldg(a.pop_users + u);
      o[6] = __ldg(a.pop_items + it);
    }
  }
}

__global__ void __launch_bounds__(256) ltr_pair_emb_bwd_kernel(int n_users, int d, int batch, const int* __restrict__ users,
                                                               const int* __restrict__ items, const float* __restrict__ emb,
                                                               const float* __restrict__ gf0, float* __restrict__ grad_emb) {
  const int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= batch) return;
  const float g = __ldg(gf0 + warp);
  if (g == 0.f) return;
  const size_t ru = (size_t)__ldg(users + warp) * d, ri = (size_t)(n_users + __ldg(items + warp)) * d;
  for (int c = lane; c < (d >> 2); c += 32) {
    const float4 eu = ldg4(emb + ru + c * 4), ei = ldg4(emb + ri + c * 4);
    atomicAdd(reinterpret_cast<float4*>(grad_emb + ru + c * 4), make_float4(g * ei.x, g * ei.y, g * ei.z, g * ei.w));
    atomicAdd(reinterpret_cast<float4*>(grad_emb + ri + c * 4), make_float4(g * eu.x, g * eu.y, g * eu.z, g * eu.w));
  }
}

struct W5 {
  float w[5];
};

// out row i = [w0·Ie | w1·Ir + w3·Id | w2·Id + w4·Ir], width d + 2D
__global__ void __launch_bounds__(256) ltr_pack_items_kernel(int64_t n_items, int d, int D, const float* __restrict__ ie,
                                                             const float* __restrict__ ir, const float* __restrict__ id, W5 w,
                                                             float* __restrict__ out) {
  const int row4 = (d + 2 * D) >> 2;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_items * row4) return;
  const int64_t i = t / row4;
  const int c = (int)(t % row4) * 4;
  float4 r;
  if (c < d) {
    const float4 e = ldg4(ie + i * d + c);
    r = make_float4(w.w[0] * e.x, w.w[0] * e.y, w.w[0] * e.z, w.w[0] * e.w);
  } else if (c < d + D) {
    const float4 x = ldg4(ir + i * D + (c - d)), y = ldg4(id + i * D + (c - d));
    r = make_float4(w.w[1] * x.x + w.w[3] * y.x, w.w[1] * x.y + w.w[3] * y.y, w.w[1] * x.z + w.w[3] * y.z, w.w[1] * x.w + w.w[3] * y.w);
  } else {
    const float4 x = ldg4(ir + i * D + (c - d - D)), y = ldg4(id + i * D + (c - d - D));
    r = make_float4(w.w[2] * y.x + w.w[4] * x.x, w.w[2] * y.y + w.w[4] * x.y, w.w[2] * y.z + w.w[4] * x.z, w.w[2] * y.w + w.w[4] * x.w);
  }
  *reinterpret_cast<float4*>(out + i * (d + 2 * D) + c) = r;
}

// out row r = [Ue[u] | Ur[u] | Ud[u]] for u = users[r]
__global__ void __launch_bounds__(256) ltr_pack_users_kernel(int64_t n_rank, const int* __restrict__ users, int d, int D,
                                                             const float* __restrict__ ue, const float* __restrict__ ur,
                                                             const float* __restrict__ ud, float* __restrict__ out) {
  const int row4 = (d + 2 * D) >> 2;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_rank * row4) return;
  const int64_t r = t / row4;
  const int c = (int)(t % row4) * 4;
  const int64_t u = users ? __ldg(users + r) : r;
  float4 v;
  if (c < d) v = ldg4(ue + u * d + c);
  else if (c < d + D) v = ldg4(ur + u * D + (c - d));
  else v = ldg4(ud + u * D + (c - d - D));
  *reinterpret_cast<float4*>(out + r * (d + 2 * D) + c) = v;
}

}  // namespace tgcn

using namespace tgcn;

extern "C" {

int tgcn_adv_select(const tgcn_graph_t* mask_graph, int64_t d, int64_t batch, int32_t n_cand, const int32_t* d_users,
                    const int32_t* d_cands, const float* d_emb, int32_t kmax, int32_t* d_out_negs, int32_t* d_out_counts,
                    float* d_out_scores, tgcn_stream_t stream) {
  TGCN_REQUIRE(mask_graph != nullptr, "graph is NULL");
  TGCN_REQUIRE(mask_graph->row_begin == 0 && mask_graph->n_rows >= mask_graph->n_users && !mask_graph->is_block, "adv_select needs a whole-graph handle");
  TGCN_REQUIRE(d > 0 && d % 4 == 0 && d <= 128 * kMaxChunks, "embedding width d=%lld must be a multiple of 4 and <= %d", (long long)d, 128 * kMaxChunks);
  TGCN_REQUIRE(batch > 0 && n_cand > 0 && n_cand <= TGCN_ADV_MAX_CANDIDATES, "bad sizes: batch=%lld n_cand=%d (max %d)", (long long)batch, n_cand, TGCN_ADV_MAX_CANDIDATES);
  TGCN_REQUIRE(kmax > 0, "kmax must be positive");
  TGCN_REQUIRE(d_users && d_cands && d_emb && d_out_negs && d_out_counts, "NULL argument");
  AdvArgs a;
  a.n_users = (int)mask_graph->n_users;
  a.n_items = (int)mask_graph->n_items;
  a.d = (int)d;
  a.batch = (int)batch;
  a.n_cand = n_cand;
  a.p2 = 2;
  while (a.p2 < n_cand) a.p2 <<= 1;
  a.kmax = kmax;
  a.users = d_users;
  a.cands = d_cands;
  a.emb = d_emb;
  a.mrowptr = mask_graph->rowptr;
  a.mcol = mask_graph->col;
  a.out_negs = d_out_negs;
  a.out_counts = d_out_counts;
  a.out_scores = d_out_scores;
  const size_t smem = (size_t)a.p2 * 8 + kAdvThreads * 4;
  adv_select_kernel<<<(unsigned)batch, kAdvThreads, smem, (cudaStream_t)stream>>>(a);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int tgcn_ltr_pairwise_features(int64_t n_users, int64_t d, int64_t D, int64_t batch, int32_t n_feat, const int32_t* d_users,
                               const int32_t* d_items, const float* d_emb, const float* d_users_rev,
                               const float* d_users_desc, const float* d_items_rev, const float* d_items_desc,
                               const float* d_pop_users, const float* d_pop_items, float* d_out, tgcn_stream_t stream) {
  TGCN_REQUIRE(batch > 0 && d > 0 && d % 4 == 0 && D > 0 && D % 4 == 0, "bad sizes: batch=%lld d=%lld D=%lld (multiples of 4)", (long long)batch, (long long)d, (long long)D);
  TGCN_REQUIRE(n_feat == 5 || n_feat == 7, "n_feat must be 5 or 7");
  TGCN_REQUIRE(d_users && d_items && d_emb && d_users_rev && d_users_desc && d_items_rev && d_items_desc && d_out, "NULL argument");
  TGCN_REQUIRE(n_feat == 5 || (d_pop_users && d_pop_items), "popularity tables required for n_feat == 7");
  LtrPairArgs a{(int)n_users, (int)d, (int)D, (int)batch, n_feat, d_users, d_items, d_emb, d_users_rev, d_users_desc,
                d_items_rev, d_items_desc, d_pop_users, d_pop_items, d_out};
  const int64_t blocks = (batch * 32 + 255) / 256;
  ltr_pairwise_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int tgcn_ltr_pairwise_emb_bwd(int64_t n_users, int64_t d, int64_t batch, const int32_t* d_users, const int32_t* d_items,
                              const float* d_emb, const float* d_gf0, float* d_grad_emb, tgcn_stream_t stream) {
  TGCN_REQUIRE(batch > 0 && d > 0 && d % 4 == 0, "bad sizes");
  TGCN_REQUIRE(d_users && d_items && d_emb && d_gf0 && d_grad_emb, "NULL argument");
  const int64_t blocks = (batch * 32 + 255) / 256;
  ltr_pair_emb_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((int)n_users, (int)d, (int)batch, d_users, d_items, d_emb, d_gf0, d_grad_emb);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int tgcn_ltr_pack_items(int64_t n_items, int64_t d, int64_t D, const float* d_items_emb, const float* d_items_rev,
                        const float* d_items_desc, const float* h_w5, float* d_out, tgcn_stream_t stream) {
  TGCN_REQUIRE(n_items > 0 && d > 0 && d % 4 == 0 && D > 0 && D % 4 == 0, "bad sizes");
  TGCN_REQUIRE(d_items_emb && d_items_rev && d_items_desc && h_w5 && d_out, "NULL argument");
  W5 w;
  for (int i = 0; i < 5; ++i) w.w[i] = h_w5[i];
  const int64_t total = n_items * ((d + 2 * D) / 4);
  ltr_pack_items_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_items, (int)d, (int)D, d_items_emb, d_items_rev, d_items_desc, w, d_out);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int tgcn_ltr_pack_users(int64_t n_rank, const int32_t* d_users, int64_t d, int64_t D, const float* d_users_emb,
                        const float* d_users_rev, const float* d_users_desc, float* d_out, tgcn_stream_t stream) {
  TGCN_REQUIRE(n_rank > 0 && d > 0 && d % 4 == 0 && D > 0 && D % 4 == 0, "bad sizes");
  TGCN_REQUIRE(d_users_emb && d_users_rev && d_users_desc && d_out, "NULL argument");
  const int64_t total = n_rank * ((d + 2 * D) / 4);
  ltr_pack_users_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_rank, d_users, (int)d, (int)D, d_users_emb, d_users_rev, d_users_desc, d_out);
  TGCN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
