// Full-ranking evaluation: score GEMM (exact fp32 FMA) fused with train-item masking and a streaming
// per-user top-k, so the (users × items) score matrix never reaches HBM (base_model.py:255-261).
//
// This file holds the exact-fp32 SIMT kernel: a 128×128×16 register-tiled SGEMM (8×8 outputs per
// thread) whose epilogue keeps, per user row, a sorted k-entry list in shared memory.  The thread
// layout gives every warp exclusive ownership of 16 user rows, so list updates need only warp-level
// synchronisation.  The fast path is one compare per score against the row's current k-th value;
// candidates that pass are checked against the user's train items (binary search in the user row of
// Â) and inserted cooperatively by the warp.  Ordering is the strict total order (score desc, item id
// asc), so the result does not depend on tile or split order.  Item ranges can be split across CTAs
// (small user counts) or GPUs (item sharding); partial lists are merged by topk_merge_kernel, which also
// applies the reference's "-inf tail" behaviour.
//
// Roofline: compute bound on the fp32 FMA pipe at 2·K flops per score; HBM traffic is the two vector
// tables (L2-resident at Electronics scale) plus (n_rank, k) outputs.
#include <limits.h>
#include <math.h>

#include "common.cuh"

namespace tgcn {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;
constexpr int kEvalThreads = 256;

struct EvalArgs {
  const int* users;
  int n_rank;
  int by_pos;  // user vectors / bias are indexed by list position instead of user id
  const float* uvec;
  int64_t ldu;
  const float* ivec;
  int64_t ldi;
  int K;
  int item_begin, item_end;
  const float* ubias;
  const float* ibias;
  const int* mrowptr;  // mask CSR rows are user ids - mrow_begin; entries are n_users + item id
  const int* mcol;
  int mrow_begin;
  int mcol_off;
  int k;
  int tiles_per_split;
  int* part_ids;  // (n_splits, n_rank, k)
  float* part_scores;
};

// Insert (s, id) into the descending list (ls, li) of length k held in shared memory; warp-cooperative.
__device__ __noinline__ void warp_list_insert(float* ls, int* li, int k, float s, int id, int lane) {
  int before = 0;
  for (int base = 0; base < k; base += 32) {
    const int j = base + lane;
    const bool b = j < k && ranks_before(ls[j], li[j], s, id);
    before += __popc(__ballot_sync(0xffffffffu, b));
  }
  const int pos = before;
  if (pos >= k) return;
  for (int base = ((k - 1) >> 5) << 5; base >= 0; base -= 32) {
    const int j = base + lane;
    float ps = 0.f;
    int pi = 0;
    const bool mv = j < k && j > pos;
    if (mv) {
      ps = ls[j - 1];
      pi = li[j - 1];
    }
    __syncwarp();
    if (mv) {
      ls[j] = ps;
      li[j] = pi;
    }
    __syncwarp();
  }
  if (lane == 0) {
    ls[pos] = s;
    li[pos] = id;
  }
  __syncwarp();
}

__device__ __forceinline__ bool is_masked(const EvalArgs& a, int user, int item) {
  if (a.mrowptr == nullptr) return false;
  const int r = user - a.mrow_begin;
  TGCN_DASSERT(r >= 0 && item >= a.item_begin && item < a.item_end);
  return sorted_contains(a.mcol, __ldg(a.mrowptr + r), __ldg(a.mrowptr + r + 1), item + a.mcol_off);
}

__global__ void __launch_bounds__(kEvalThreads, 1) eval_topk_simt_kernel(const EvalArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);             // [BK][BM+PAD]
  float* Bs = As + BK * (BM + PAD);                            // [BK][BN+PAD]
  float* list_s = Bs + BK * (BN + PAD);                        // [BM][k]
  int* list_i = reinterpret_cast<int*>(list_s + BM * a.k);     // [BM][k]
  int* urow = list_i + BM * a.k;                               // [BM]

  const int tid = threadIdx.x, lane = tid & 31;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM;
  const int k = a.k;

  for (int i = tid; i < BM * k; i += kEvalThreads) {
    list_s[i] = -INFINITY;
    list_i[i] = INT_MAX;
  }
  if (tid < BM) {
    const int m = m0 + tid;
    urow[tid] = m < a.n_rank ? (a.users ? __ldg(a.users + m) : m) : -1;
  }
  __syncthreads();

  // global->smem staging assignments: each thread moves 2 float4 of A and 2 of B per K-chunk
  const int ld_row = tid >> 2;        // 0..63 (+64 on the second pass)
  const int ld_k = (tid & 3) * 4;     // 0,4,8,12
  const float* a_ptr[2];
  bool a_ok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int u = urow[ld_row + 64 * h];
    a_ok[h] = u >= 0;
    const int vrow = a.by_pos ? m0 + ld_row + 64 * h : u;
    a_ptr[h] = a.uvec + (size_t)(u >= 0 ? vrow : 0) * a.ldu;
  }
  float ub[8];
  int um[8];  // local user rows of this thread: ty*4+i and 64+ty*4+i
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    um[i] = (i < 4 ? 0 : 64) + ty * 4 + (i & 3);
    const int u = urow[um[i]];
    ub[i] = (a.ubias && u >= 0) ? __ldg(a.ubias + (a.by_pos ? m0 + um[i] : u)) : 0.f;
  }

  const int n_items_range = a.item_end - a.item_begin;
  const int n_tiles = (n_items_range + BN - 1) / BN;
  const int tile_begin = blockIdx.y * a.tiles_per_split;
  const int tile_end = min(n_tiles, tile_begin + a.tiles_per_split);
  const int n_kchunks = (a.K + BK - 1) / BK;

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int n0 = a.item_begin + tile * BN;
    const float* b_ptr[2];
    bool b_ok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + ld_row + 64 * h;
      b_ok[h] = n < a.item_end;
      b_ptr[h] = a.ivec + (size_t)(b_ok[h] ? n : a.item_begin) * a.ldi;
    }
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    auto fetch = [&](int kc) {
      const int kk = kc * BK + ld_k;
      const bool kin = kk < a.K;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        ra[h] = (a_ok[h] && kin) ? ldg4(a_ptr[h] + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
        rb[h] = (b_ok[h] && kin) ? ldg4(b_ptr[h] + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    fetch(0);
    for (int kc = 0; kc < n_kchunks; ++kc) {
      __syncthreads();  // previous chunk fully consumed
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = ld_row + 64 * h;
        As[(ld_k + 0) * (BM + PAD) + r] = ra[h].x;
        As[(ld_k + 1) * (BM + PAD) + r] = ra[h].y;
        As[(ld_k + 2) * (BM + PAD) + r] = ra[h].z;
        As[(ld_k + 3) * (BM + PAD) + r] = ra[h].w;
        Bs[(ld_k + 0) * (BN + PAD) + r] = rb[h].x;
        Bs[(ld_k + 1) * (BN + PAD) + r] = rb[h].y;
        Bs[(ld_k + 2) * (BN + PAD) + r] = rb[h].z;
        Bs[(ld_k + 3) * (BN + PAD) + r] = rb[h].w;
      }
      __syncthreads();
      if (kc + 1 < n_kchunks) fetch(kc + 1);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(As + kk * (BM + PAD) + ty * 4);
        const float4 a1 = *reinterpret_cast<const float4*>(As + kk * (BM + PAD) + 64 + ty * 4);
        const float4 b0 = *reinterpret_cast<const float4*>(Bs + kk * (BN + PAD) + tx * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(Bs + kk * (BN + PAD) + 64 + tx * 4);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }

    // ---- fused epilogue: bias, threshold filter, masked insert -------------------------------------
    int item_id[8];
    float ib[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      item_id[j] = n0 + (j < 4 ? 0 : 64) + tx * 4 + (j & 3);
      ib[j] = (a.ibias && item_id[j] < a.item_end) ? __ldg(a.ibias + item_id[j]) : 0.f;
    }
    unsigned long long pass = 0ull;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool uok = urow[um[i]] >= 0;
      const float thr = list_s[um[i] * k + k - 1];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[i][j] += ub[i] + ib[j];
        const bool p = uok && item_id[j] < a.item_end && acc[i][j] >= thr;
        pass |= (unsigned long long)p << (i * 8 + j);
      }
    }
    if (__any_sync(0xffffffffu, pass != 0ull)) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          unsigned m = __ballot_sync(0xffffffffu, (pass >> (i * 8 + j)) & 1ull);
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float s = __shfl_sync(0xffffffffu, acc[i][j], src);
            const int ml = __shfl_sync(0xffffffffu, um[i], src);
            const int it = __shfl_sync(0xffffffffu, item_id[j], src);
            float* ls = list_s + ml * k;
            int* li = list_i + ml * k;
            if (!ranks_before(s, it, ls[k - 1], li[k - 1])) continue;
            if (is_masked(a, urow[ml], it)) continue;
            warp_list_insert(ls, li, k, s, it, lane);
          }
        }
      }
    }
  }

  __syncthreads();
  for (int i = tid; i < BM * k; i += kEvalThreads) {
    const int m = m0 + i / k;
    if (m < a.n_rank) {
      const size_t o = ((size_t)blockIdx.y * a.n_rank + m) * k + (i % k);
      a.part_ids[o] = list_i[i];
      a.part_scores[o] = list_s[i];
    }
  }
}

struct MergeArgs {
  const int* users;
  int n_rows, n_parts, k;
  const int* part_ids;
  const float* part_scores;
  const int* mrowptr;
  const int* mcol;
  int mrow_begin, mcol_off;
  int finalize;
  int* out_ids;
  float* out_scores;
  const int* n_rows_dev;  // device-gated second pass of the screened path: merge only the first *n_rows_dev rows ...
  const int* out_rows;    // ... and write row r to out_rows[r]
};

// One warp per row: insert the n_parts·k candidates into a shared-memory list, then complete short
// lists with masked (train) items in ascending id order at score -inf (G9).
__global__ void __launch_bounds__(128) topk_merge_kernel(const MergeArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + wib;
  const int k = a.k;
  float* ls = reinterpret_cast<float*>(smem_raw) + wib * k;
  int* li = reinterpret_cast<int*>(reinterpret_cast<float*>(smem_raw) + 4 * k) + wib * k;
  if (row >= a.n_rows || (a.n_rows_dev && row >= __ldg(a.n_rows_dev))) return;
  const size_t orow = a.out_rows ? (size_t)__ldg(a.out_rows + row) : (size_t)row;
  for (int j = lane; j < k; j += 32) {
    ls[j] = -INFINITY;
    li[j] = INT_MAX;
  }
  __syncwarp();
  for (int p = 0; p < a.n_parts; ++p) {
    const size_t base = ((size_t)p * a.n_rows + row) * k;
    for (int j = 0; j < k; ++j) {
      const int id = __ldg(a.part_ids + base + j);
      if (id == INT_MAX) break;  // partial lists are sorted: sentinels are last
      const float s = __ldg(a.part_scores + base + j);
      if (!ranks_before(s, id, ls[k - 1], li[k - 1])) break;  // the rest of this part ranks even lower
      warp_list_insert(ls, li, k, s, id, lane);
    }
  }
  int real = 0;
  for (int base = 0; base < k; base += 32) {
    const int j = base + lane;
    real += __popc(__ballot_sync(0xffffffffu, j < k && li[j] != INT_MAX));
  }
  int lo = 0, deg = 0;
  if (a.finalize && a.mrowptr) {
    const int u = (a.users ? __ldg(a.users + row) : row) - a.mrow_begin;
    lo = __ldg(a.mrowptr + u);
    deg = __ldg(a.mrowptr + u + 1) - lo;
  }
  for (int j = lane; j < k; j += 32) {
    int id = li[j];
    float s = ls[j];
    if (a.finalize && j >= real) {
      const int t = j - real;
      id = t < deg ? __ldg(a.mcol + lo + t) - a.mcol_off : -1;
      s = -INFINITY;
    }
    a.out_ids[orow * k + j] = id;
    a.out_scores[orow * k + j] = s;
  }
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
    else return 148;
  }
  return n;
}

// (Round 2 experiment, reverted: for the STREAMED tensor-core variant — LTR ranking, K = 1600, whose 822 MB item operand every
// user tile re-streams; ncu: 61.9 GB of DRAM reads per call — splitting the item range until one split's share of the operand
// is L2-sized (17 splits instead of 2) made the call 4 % SLOWER, 18.9 -> 19.6 ms: the per-CTA pipeline fill, list restarts and
// the wider merge cost more than the DRAM traffic they remove.  profiles/r02/README.md.)
void eval_split_plan(int64_t n_rank, int64_t n_items_range, int bn, int* n_splits, int* tiles_per_split) {
  const int64_t m_tiles = (n_rank + BM - 1) / BM;
  const int64_t n_tiles = (n_items_range + bn - 1) / bn;
  int64_t want = (2ll * sm_count() + m_tiles - 1) / m_tiles;  // aim for >= 2 CTAs per SM
  if (want < 1) want = 1;
  if (want > n_tiles) want = n_tiles;
  if (want > 64) want = 64;
  const int64_t tps = (n_tiles + want - 1) / want;
  *tiles_per_split = (int)(tps > 0 ? tps : 1);
  *n_splits = (int)((n_tiles + *tiles_per_split - 1) / *tiles_per_split);
  if (*n_splits < 1) *n_splits = 1;
}

// tensor-core path (eval_tc.cu)
bool eval_tc_eligible(int64_t K, int32_t k, bool has_bias);
int eval_tc_tile_n(int64_t K, bool has_bias);
int64_t eval_tc_workspace_bytes(int64_t n_rank, int64_t n_range, int64_t K, int32_t k, bool has_bias, int n_splits);
int eval_topk_tc(const int* mrowptr, const int* mcol, int mrow_begin, int mcol_off, int64_t n_rank, const int32_t* d_users,
                 int by_pos, const float* d_user_vecs, int64_t ldu, const float* d_item_vecs, int64_t ldi, int64_t K, int64_t item_begin,
                 int64_t item_end, const float* d_user_bias, const float* d_item_bias, int32_t k, int finalize, int* d_out_ids,
                 float* d_out_scores, void* d_workspace, int64_t workspace_bytes, int* n_splits_out, int** part_ids_out,
                 float** part_scores_out, cudaStream_t s, const TcGate* gate);
bool eval_tc_screen_eligible(int64_t K, int32_t k, bool has_bias);
bool eval_tc_screen_auto(int64_t n_range, int64_t K, int32_t k, bool has_bias);
int64_t eval_tc_screen_queue_offset(int64_t n_rank, int64_t n_range, int64_t K, int32_t k, bool has_bias);
int eval_tc_screen_rescore(const TcGate* gate, int64_t n_rank, const int32_t* d_users, int by_pos, const float* d_user_vecs, int64_t ldu,
                           const float* d_item_vecs, int64_t ldi, int64_t K, int32_t k, const float* d_user_bias, const float* d_item_bias,
                           int* d_out_ids, float* d_out_scores, cudaStream_t s);
int eval_topk_screen(const int* mrowptr, const int* mcol, int mrow_begin, int mcol_off, int64_t n_rank, const int32_t* d_users, int by_pos,
                     const float* d_user_vecs, int64_t ldu, const float* d_item_vecs, int64_t ldi, int64_t K, int64_t item_begin,
                     int64_t item_end, const float* d_user_bias, const float* d_item_bias, int32_t k, int finalize, int* d_out_ids,
                     float* d_out_scores, void* d_workspace, int64_t workspace_bytes, int* n_splits_out, int** part_ids_out,
                     float** part_scores_out, TcGate* gate_out, cudaStream_t s);

static int mask_fields(const tgcn_graph* g, const int** rowptr, const int** col, int* row_begin, int* col_off) {
  if (g) {
    // rows of the handle are users row_begin .. row_begin + n_rows - 1 (the caller only ranks users it covers)
    *rowptr = g->rowptr;
    *col = g->col;
    *row_begin = (int)g->row_begin;
    *col_off = g->mask_col_off;
  } else {
    *rowptr = nullptr;
    *col = nullptr;
    *row_begin = 0;
    *col_off = 0;
  }
  return 0;
}

}  // namespace tgcn

using namespace tgcn;

extern "C" {

int64_t tgcn_eval_workspace_bytes(int64_t n_rank, int64_t n_items_range, int64_t K, int32_t k) {
  if (n_rank <= 0 || n_items_range <= 0 || k <= 0 || K <= 0) return -1;
  int ns, tps;
  eval_split_plan(n_rank, n_items_range, BN, &ns, &tps);
  int64_t need = (int64_t)ns * n_rank * k * 8 + 512;
  if (eval_tc_eligible(K, k, true)) {  // sized for the widest variant (with the bias chunk)
    eval_split_plan(n_rank, n_items_range, eval_tc_tile_n(K, true), &ns, &tps);
    int64_t tc = eval_tc_workspace_bytes(n_rank, n_items_range, K, k, true, ns);
    eval_split_plan(n_rank, n_items_range, eval_tc_tile_n(K, false), &ns, &tps);
    const int64_t tc2 = eval_tc_workspace_bytes(n_rank, n_items_range, K, k, true, ns);
    if (tc2 > tc) tc = tc2;
    if (tc > need) need = tc;
  }
  return need;
}

static int launch_merge(const tgcn_graph_t* mask_graph, int64_t n_rows, const int32_t* d_users, int32_t n_parts, int32_t k,
                        const int32_t* d_part_ids, const float* d_part_scores, int32_t finalize, int32_t* d_out_ids, float* d_out_scores,
                        const int* n_rows_dev, const int* out_rows, tgcn_stream_t stream) {
  TGCN_REQUIRE(n_rows > 0 && n_parts > 0 && k > 0 && k <= TGCN_MAX_TOPK, "bad sizes: n_rows=%lld n_parts=%d k=%d", (long long)n_rows, n_parts, k);
  TGCN_REQUIRE(d_part_ids && d_part_scores && d_out_ids && d_out_scores, "NULL argument");
  MergeArgs m;
  m.n_rows_dev = n_rows_dev;
  m.out_rows = out_rows;
  m.users = d_users;
  m.n_rows = (int)n_rows;
  m.n_parts = n_parts;
  m.k = k;
  m.part_ids = d_part_ids;
  m.part_scores = d_part_scores;
  if (int rc = mask_fields(mask_graph, &m.mrowptr, &m.mcol, &m.mrow_begin, &m.mcol_off)) return rc;
  m.finalize = finalize;
  m.out_ids = d_out_ids;
  m.out_scores = d_out_scores;
  const int64_t blocks = (n_rows + 3) / 4;
  topk_merge_kernel<<<(unsigned)blocks, 128, 4 * k * 8, (cudaStream_t)stream>>>(m);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int64_t tgcn_eval_screen_queue_offset(int64_t n_rank, int64_t n_items_range, int64_t K, int32_t k, int32_t has_bias) {
  if (n_rank <= 0 || n_items_range <= 0 || k <= 0 || K <= 0 || !eval_tc_eligible(K, k, has_bias != 0)) return -1;
  return eval_tc_screen_queue_offset(n_rank, n_items_range, K, k, has_bias != 0);
}

int32_t tgcn_eval_resolve_precision(int64_t n_items_range, int64_t K, int32_t k, int32_t has_bias, int32_t precision) {
  if (precision != 0) return precision;
  if (!eval_tc_eligible(K, k, has_bias != 0)) return 1;
  return eval_tc_screen_auto(n_items_range, K, k, has_bias != 0) ? 3 : 2;
}

int tgcn_topk_merge(const tgcn_graph_t* mask_graph, int64_t n_rows, const int32_t* d_users, int32_t n_parts, int32_t k,
                    const int32_t* d_part_ids, const float* d_part_scores, int32_t finalize, int32_t* d_out_ids,
                    float* d_out_scores, tgcn_stream_t stream) {
  return launch_merge(mask_graph, n_rows, d_users, n_parts, k, d_part_ids, d_part_scores, finalize, d_out_ids, d_out_scores, nullptr, nullptr,
                      stream);
}

int tgcn_eval_topk(const tgcn_graph_t* mask_graph, int64_t n_rank, const int32_t* d_users, const float* d_user_vecs,
                   int64_t ldu, const float* d_item_vecs, int64_t ldi, int64_t K, int64_t item_begin, int64_t item_end,
                   const float* d_user_bias, const float* d_item_bias, int32_t vecs_by_position, int32_t precision, int32_t k,
                   int32_t finalize, int32_t* d_out_ids, float* d_out_scores, void* d_workspace, int64_t workspace_bytes,
                   tgcn_stream_t stream) {
  TGCN_REQUIRE(n_rank > 0 && K > 0 && K % 4 == 0 && ldu % 4 == 0 && ldi % 4 == 0 && ldu >= K && ldi >= K,
               "bad shapes: n_rank=%lld K=%lld ldu=%lld ldi=%lld (K, ld multiples of 4)", (long long)n_rank, (long long)K, (long long)ldu, (long long)ldi);
  TGCN_REQUIRE(item_begin >= 0 && item_end > item_begin && item_end < (1ll << 31), "bad item range [%lld, %lld)", (long long)item_begin, (long long)item_end);
  TGCN_REQUIRE(k > 0 && k <= TGCN_MAX_TOPK, "k=%d out of range (1..%d)", k, TGCN_MAX_TOPK);
  TGCN_REQUIRE(d_user_vecs && d_item_vecs && d_out_ids && d_out_scores, "NULL argument");
  TGCN_REQUIRE(((uintptr_t)d_user_vecs % 16 == 0) && ((uintptr_t)d_item_vecs % 16 == 0), "vector tables must be 16-byte aligned");
  TGCN_REQUIRE(precision >= 0 && precision <= 3, "precision must be 0 (auto), 1 (fp32), 2 (3xTF32) or 3 (TF32 screen + exact re-scoring)");
  const int64_t need = tgcn_eval_workspace_bytes(n_rank, item_end - item_begin, K, k);
  TGCN_REQUIRE(d_workspace && workspace_bytes >= need, "workspace too small: need %lld bytes", (long long)need);
  const bool tc_ok = eval_tc_eligible(K, k, d_user_bias || d_item_bias) && ldu % 4 == 0 && ldi % 4 == 0;
  TGCN_REQUIRE((precision != 2 && precision != 3) || tc_ok, "the tensor-core paths need k <= 64 (and K <= 8192)");
  if (tc_ok && precision != 1) {
    const int* mrowptr;
    const int* mcol;
    int mrow_begin, mcol_off, n_splits;
    int* part_ids;
    float* part_scores;
    if (int rc = mask_fields(mask_graph, &mrowptr, &mcol, &mrow_begin, &mcol_off)) return rc;
    const bool screen_ok = eval_tc_screen_eligible(K, k, d_user_bias || d_item_bias);
    TGCN_REQUIRE(precision != 3 || screen_ok, "the screened path needs k <= 24");
    if (precision == 3 || (precision == 0 && eval_tc_screen_auto(item_end - item_begin, K, k, d_user_bias || d_item_bias))) {
      // pass 1: one TF32 product per score, exact re-scoring of the candidates, certificate; pass 2 (gated on the device by the
      // length of the queue pass 1 leaves): the 3xTF32 variant on the rows that could not be certified
      TcGate gate;
      if (int rc = eval_topk_screen(mrowptr, mcol, mrow_begin, mcol_off, n_rank, d_users, vecs_by_position, d_user_vecs, ldu, d_item_vecs, ldi,
                                    K, item_begin, item_end, d_user_bias, d_item_bias, k, finalize, d_out_ids, d_out_scores, d_workspace,
                                    workspace_bytes, &n_splits, &part_ids, &part_scores, &gate, (cudaStream_t)stream))
        return rc;
      if (n_splits > 1)
        if (int rc = tgcn_topk_merge(mask_graph, n_rank, d_users, n_splits, k, part_ids, part_scores, finalize, d_out_ids, d_out_scores, stream))
          return rc;
      if (int rc = eval_topk_tc(mrowptr, mcol, mrow_begin, mcol_off, n_rank, d_users, vecs_by_position, d_user_vecs, ldu, d_item_vecs, ldi, K,
                                item_begin, item_end, d_user_bias, d_item_bias, k, finalize, d_out_ids, d_out_scores, d_workspace,
                                workspace_bytes, &n_splits, &part_ids, &part_scores, (cudaStream_t)stream, &gate))
        return rc;
      if (n_splits > 1)
        if (int rc = launch_merge(mask_graph, n_rank, gate.users, n_splits, k, part_ids, part_scores, finalize, d_out_ids, d_out_scores,
                                  gate.count, gate.rows, stream))
          return rc;
      // the rows of the second pass get the same exact fp32 scores (and order) as the rest
      return eval_tc_screen_rescore(&gate, n_rank, d_users, vecs_by_position, d_user_vecs, ldu, d_item_vecs, ldi, K, k, d_user_bias, d_item_bias,
                                    d_out_ids, d_out_scores, (cudaStream_t)stream);
    }
    if (int rc = eval_topk_tc(mrowptr, mcol, mrow_begin, mcol_off, n_rank, d_users, vecs_by_position, d_user_vecs, ldu, d_item_vecs,
                              ldi, K, item_begin, item_end, d_user_bias, d_item_bias, k, finalize, d_out_ids, d_out_scores, d_workspace,
                              workspace_bytes, &n_splits, &part_ids, &part_scores, (cudaStream_t)stream, nullptr))
      return rc;
    if (n_splits == 1) return 0;  // the kernel wrote (and completed) the final table itself
    return tgcn_topk_merge(mask_graph, n_rank, d_users, n_splits, k, part_ids, part_scores, finalize, d_out_ids, d_out_scores, stream);
  }
  EvalArgs a;
  a.users = d_users;
  a.n_rank = (int)n_rank;
  a.by_pos = vecs_by_position;
  a.uvec = d_user_vecs;
  a.ldu = ldu;
  a.ivec = d_item_vecs;
  a.ldi = ldi;
  a.K = (int)K;
  a.item_begin = (int)item_begin;
  a.item_end = (int)item_end;
  a.ubias = d_user_bias;
  a.ibias = d_item_bias;
  if (int rc = mask_fields(mask_graph, &a.mrowptr, &a.mcol, &a.mrow_begin, &a.mcol_off)) return rc;
  a.k = k;
  int n_splits;
  eval_split_plan(n_rank, item_end - item_begin, BN, &n_splits, &a.tiles_per_split);
  a.part_ids = (int*)d_workspace;
  a.part_scores = (float*)((char*)d_workspace + (size_t)n_splits * n_rank * k * 4);
  const size_t smem = (size_t)(BK * (BM + PAD) + BK * (BN + PAD)) * 4 + (size_t)BM * k * 8 + BM * 4;
  cudaStream_t s = (cudaStream_t)stream;
  TGCN_CHECK_CUDA(cudaFuncSetAttribute(eval_topk_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((n_rank + BM - 1) / BM), (unsigned)n_splits);
  eval_topk_simt_kernel<<<grid, kEvalThreads, smem, s>>>(a);
  TGCN_CHECK_LAUNCH();
  return tgcn_topk_merge(mask_graph, n_rank, d_users, n_splits, k, a.part_ids, a.part_scores, finalize, d_out_ids, d_out_scores, stream);
}

}  // extern "C"
