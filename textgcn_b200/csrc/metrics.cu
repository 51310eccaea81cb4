// n4: recall / precision / hit / ndcg / f1 @k straight from the (n, kmax) id table the fused eval
// kernel leaves on the device and a ragged (CSR) list of true test items — the reference's utils.calculate_metrics
// (utils.py:11-63) is a pandas row-apply that takes 65 % of `evaluate` and needs a Python list per user, which is not
// feasible at 10M users.  Semantics kept exactly: the intersection counts DISTINCT predicted ids found in y_true
// (np.intersect1d, utils.py:46); recall divides by len(y_true) with duplicates (:15-16, :39); ndcg's relevance marks every
// occurrence of a hit id (:31), its discounts are 1/log2(arange(2, k+2)) and its ideal is min(|y_true|, k) ones (:23-33);
// f1 is 0 where precision + recall is 0 (:55-62); every metric is the mean over rows, accumulated in float64.
//
// One warp per row (grid-stride), per-warp float64 sums in shared memory, per-block partials reduced in a fixed order by
// a second one-block kernel: the result is deterministic.  Bound: HBM read of n·kmax·4 + Σ|y_true|·4 bytes.
#include "common.cuh"

namespace tgcn {

constexpr int kMetricThreads = 256;
constexpr int kMetricWarps = kMetricThreads / 32;
constexpr int kMetricBlocks = 148 * 8;
constexpr int kMaxKs = 8;
constexpr int kNumMetrics = 5;  // recall, precision, hit, ndcg, f1

struct MetricsArgs {
  int64_t n_rows;
  int kmax, n_ks;
  int ks[kMaxKs];
  const int* pred;
  const long long* tptr;
  const int* tids;
  double* partials;  // (gridDim.x, n_ks, 5)
};

__global__ void __launch_bounds__(kMetricThreads) topk_metrics_kernel(const MetricsArgs a) {
  __shared__ double disc[TGCN_MAX_TOPK];   // 1 / log2(j + 2)
  __shared__ double cum[TGCN_MAX_TOPK];    // cum[j] = disc[0] + ... + disc[j]
  __shared__ int spred[kMetricWarps][TGCN_MAX_TOPK];
  __shared__ double sums[kMetricWarps][kMaxKs][kNumMetrics];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < a.kmax) disc[threadIdx.x] = 1.0 / log2((double)(threadIdx.x + 2));
  for (int i = threadIdx.x; i < kMetricWarps * kMaxKs * kNumMetrics; i += kMetricThreads) (&sums[0][0][0])[i] = 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = 0.0;
    for (int j = 0; j < a.kmax; ++j) {
      run += disc[j];
      cum[j] = run;
    }
  }
  __syncthreads();
  const int64_t n_warps = (int64_t)gridDim.x * kMetricWarps;
  const int n_chunks = (a.kmax + 31) >> 5;
  for (int64_t row = (int64_t)blockIdx.x * kMetricWarps + wib; row < a.n_rows; row += n_warps) {
    const int* p = a.pred + row * a.kmax;
    for (int j = lane; j < a.kmax; j += 32) spred[wib][j] = __ldg(p + j);
    __syncwarp();
    const long long t0 = __ldg(a.tptr + row), t1 = __ldg(a.tptr + row + 1);
    const double tlen = (double)(t1 - t0);
    unsigned hit_mask[TGCN_MAX_TOPK / 32], uniq_mask[TGCN_MAX_TOPK / 32];
#pragma unroll
    for (int c = 0; c < TGCN_MAX_TOPK / 32; ++c) {
      hit_mask[c] = uniq_mask[c] = 0u;
      if (c >= n_chunks) continue;
      const int j = c * 32 + lane;
      bool hit = false, first = true;
      if (j < a.kmax) {
        const int id = spred[wib][j];
        if (id >= 0) {
          for (long long t = t0; t < t1; ++t) hit |= __ldg(a.tids + t) == id;
          if (hit)
            for (int q = 0; q < j; ++q) first &= spred[wib][q] != id;  // a repeated id counts once in the intersection
        }
      }
      hit_mask[c] = __ballot_sync(0xffffffffu, hit);
      uniq_mask[c] = __ballot_sync(0xffffffffu, hit && first);
    }
    for (int ki = 0; ki < a.n_ks; ++ki) {
      const int k = a.ks[ki];
      int inter = 0;
      double dcg = 0.0;
#pragma unroll
      for (int c = 0; c < TGCN_MAX_TOPK / 32; ++c) {
        const int rem = k - c * 32;
        if (rem <= 0) continue;
        const unsigned low = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
        inter += __popc(uniq_mask[c] & low);
        if ((hit_mask[c] & low) >> lane & 1u) dcg += disc[c * 32 + lane];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dcg += __shfl_xor_sync(0xffffffffu, dcg, o);
      if (lane == 0) {
        const double rec = (double)inter / tlen, prec = (double)inter / (double)k;
        const int ideal = (int)(tlen < (double)k ? tlen : (double)k);
        const double ndcg = ideal > 0 ? dcg / cum[ideal - 1] : 0.0;
        const double den = rec + prec;
        sums[wib][ki][0] += rec;
        sums[wib][ki][1] += prec;
        sums[wib][ki][2] += inter > 0 ? 1.0 : 0.0;
        sums[wib][ki][3] += ndcg;
        sums[wib][ki][4] += den != 0.0 ? rec * prec * 2.0 / den : 0.0;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < a.n_ks * kNumMetrics) {
    const int ki = threadIdx.x / kNumMetrics, m = threadIdx.x % kNumMetrics;
    double s = 0.0;
    for (int w = 0; w < kMetricWarps; ++w) s += sums[w][ki][m];
    a.partials[(size_t)blockIdx.x * a.n_ks * kNumMetrics + threadIdx.x] = s;
  }
}

// One warp per value: lane l adds the partials of blocks l, l + 32, ... in order, then a fixed shuffle tree — the same
// summation order every run.
__global__ void topk_metrics_finalize_kernel(const double* __restrict__ partials, int n_blocks, int n_vals, double n_rows,
                                             double* __restrict__ out) {
  const int t = blockIdx.x, lane = threadIdx.x;  // one 32-thread block per value
  if (t >= n_vals) return;
  double s = 0.0;
  for (int b = lane; b < n_blocks; b += 32) s += partials[(size_t)b * n_vals + t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[t] = s / n_rows;
}

}  // namespace tgcn

using namespace tgcn;

extern "C" {

int64_t tgcn_topk_metrics_workspace_bytes(void) { return (int64_t)kMetricBlocks * kMaxKs * kNumMetrics * (int64_t)sizeof(double); }

int tgcn_topk_metrics(int64_t n_rows, int32_t kmax, const int32_t* d_pred_ids, const int64_t* d_true_ptr, const int32_t* d_true_ids,
                      int32_t n_ks, const int32_t* h_ks, double* d_out, void* d_workspace, int64_t workspace_bytes,
                      tgcn_stream_t stream) {
  TGCN_REQUIRE(n_rows > 0 && kmax > 0 && kmax <= TGCN_MAX_TOPK, "bad sizes: n_rows=%lld kmax=%d (max %d)", (long long)n_rows, kmax, TGCN_MAX_TOPK);
  TGCN_REQUIRE(n_ks > 0 && n_ks <= kMaxKs && h_ks, "n_ks=%d out of range [1, %d]", n_ks, kMaxKs);
  TGCN_REQUIRE(d_pred_ids && d_true_ptr && d_true_ids && d_out, "NULL argument");
  TGCN_REQUIRE(d_workspace && workspace_bytes >= tgcn_topk_metrics_workspace_bytes(), "workspace too small");
  MetricsArgs a;
  a.n_rows = n_rows;
  a.kmax = kmax;
  a.n_ks = n_ks;
  for (int i = 0; i < n_ks; ++i) {
    TGCN_REQUIRE(h_ks[i] > 0 && h_ks[i] <= kmax, "k=%d out of range [1, kmax=%d]", h_ks[i], kmax);
    a.ks[i] = h_ks[i];
  }
  a.pred = d_pred_ids;
  a.tptr = (const long long*)d_true_ptr;
  a.tids = d_true_ids;
  a.partials = (double*)d_workspace;
  int64_t blocks = (n_rows + kMetricWarps - 1) / kMetricWarps;
  if (blocks > kMetricBlocks) blocks = kMetricBlocks;
  cudaStream_t s = (cudaStream_t)stream;
  topk_metrics_kernel<<<(unsigned)blocks, kMetricThreads, 0, s>>>(a);
  TGCN_CHECK_LAUNCH();
  topk_metrics_finalize_kernel<<<n_ks * kNumMetrics, 32, 0, s>>>(a.partials, (int)blocks, n_ks * kNumMetrics, (double)n_rows, d_out);
  TGCN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
