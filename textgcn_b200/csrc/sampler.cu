// n1: the BPR negative sampler on the GPU.  Replaces BaseDataset._cache_samples / __getitem__
// (dataset.py:167-193), which is 43 % of a CPU training step in the reference and cannot terminate when a user has
// fewer than bucket_len·neg_samples non-positive items (G20).  One thread per batch row: a uniform positive from
// the user's train row of Â and n_neg uniform negatives by rejection (binary search in the same row), distinct
// within the row.  Counter-based RNG (murmur3 finaliser over seed ⊕ row ⊕ draw), so a (seed, batch index) pair
// reproduces its samples.  Only statistical parity with the reference's Python `random` is possible.
#include "common.cuh"

namespace tgcn {

__device__ __forceinline__ uint32_t mix32(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (uint32_t)x;
}

struct SampleArgs {
  const int* rowptr;
  const int* col;
  int n_users, n_items, batch, n_neg, max_tries;
  unsigned long long seed;
  const int* users;
  long long* out;  // (batch, 2 + n_neg)
  int* fail_count;
};

__global__ void __launch_bounds__(256) sample_bpr_kernel(const SampleArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.batch) return;
  const int u = __ldg(a.users + b);
  const int lo = __ldg(a.rowptr + u), hi = __ldg(a.rowptr + u + 1);
  const int deg = hi - lo;
  const unsigned long long key = a.seed ^ ((unsigned long long)(b + 1) * 0x9E3779B97F4A7C15ULL);
  long long* row = a.out + (size_t)b * (2 + a.n_neg);
  row[0] = u;
  row[1] = deg > 0 ? __ldg(a.col + lo + (int)(((unsigned long long)mix32(key) * (unsigned)deg) >> 32)) - a.n_users : -1;
  for (int j = 0; j < a.n_neg; ++j) {
    int cand = -1;
    bool ok = false;
    for (int t = 0; t < a.max_tries && !ok; ++t) {
      cand = (int)(((unsigned long long)mix32(key + 1 + (unsigned long long)j * a.max_tries + t) * (unsigned)a.n_items) >> 32);
      ok = !sorted_contains(a.col, lo, hi, cand + a.n_users);
      for (int q = 0; q < j && ok; ++q) ok = row[2 + q] != cand;
    }
    if (!ok) {  // dense user: walk cyclically from the last candidate (terminates; the reference would spin forever)
      for (int s = 1; s <= a.n_items && !ok; ++s) {
        const int c = (cand + s) % a.n_items;
        ok = !sorted_contains(a.col, lo, hi, c + a.n_users);
        for (int q = 0; q < j && ok; ++q) ok = row[2 + q] != c;
        if (ok) cand = c;
      }
    }
    if (!ok) {
      cand = -1;
      atomicAdd(a.fail_count, 1);
    }
    row[2 + j] = cand;
  }
}

}  // namespace tgcn

using namespace tgcn;

extern "C" int tgcn_sample_bpr_batch(const tgcn_graph_t* g, int64_t batch, int32_t n_neg, const int32_t* d_users, uint64_t seed,
                                     int32_t max_tries, int64_t* d_out, int32_t* d_fail_count, tgcn_stream_t stream) {
  TGCN_REQUIRE(g != nullptr && !g->is_block, "sampler needs a whole-graph handle");
  TGCN_REQUIRE(batch > 0 && n_neg > 0 && n_neg <= 64 && max_tries > 0, "bad sizes: batch=%lld n_neg=%d", (long long)batch, n_neg);
  TGCN_REQUIRE(d_users && d_out && d_fail_count, "NULL argument");
  SampleArgs a{g->rowptr, g->col, (int)g->n_users, (int)g->n_items, (int)batch, n_neg, max_tries, (unsigned long long)seed,
               d_users, (long long*)d_out, d_fail_count};
  sample_bpr_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  TGCN_CHECK_LAUNCH();
  return 0;
}
