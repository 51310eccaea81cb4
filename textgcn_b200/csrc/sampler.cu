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
  TGCN_DASSERT(u >= 0 && u < a.n_users);
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

// n2  Bernoulli(1 - p) edge keep-mask drawn on the device (replaces torch.rand(nnz) on the host + H2D, base_model.py:82):
// one byte per nnz, 16 entries per thread from four 64-bit hashes, written as one 128-bit store.
// draws != NULL: the per-draw part of the seed comes from a device counter (seed += *draws · 0xD6E8FEB86659FD93), so the
// launch carries no per-step host scalar (CUDA-graph replay of the training step).
__global__ void __launch_bounds__(256) dropout_mask_kernel(long long n16, long long nnz, unsigned long long seed, uint32_t thresh,
                                                           const unsigned long long* __restrict__ draws, uint8_t* __restrict__ keep) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n16) return;
  if (draws != nullptr) seed += __ldg(draws) * 0xD6E8FEB86659FD93ULL;
  uint32_t w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t packed = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const unsigned long long idx = (unsigned long long)t * 16 + q * 4 + b;
      const uint32_t r = mix32(seed + idx * 0x9E3779B97F4A7C15ULL);
      packed |= (r < thresh ? 1u : 0u) << (8 * b);
    }
    w[q] = packed;
  }
  if ((t + 1) * 16 <= nnz) {
    *reinterpret_cast<uint4*>(keep + t * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    for (long long i = t * 16; i < nnz; ++i) keep[i] = (uint8_t)((w[(i - t * 16) >> 2] >> (8 * ((i - t * 16) & 3))) & 0xff);
  }
}

// a14  AdvSamplDataset.__getitem__ (advanced_sampling.py:21-22): n_cand DISTINCT uniform items per batch row.
// Distinctness without a set: position q of row b maps through a keyed 4-round Feistel permutation of [0, 2^bits)
// with cycle walking back into [0, n_items), i.e. the first n_cand entries of a random permutation of the items.
__device__ __forceinline__ uint32_t feistel_perm(uint32_t x, int half_bits, unsigned long long key, uint32_t n) {
  const uint32_t mask = (1u << half_bits) - 1u;
  do {
    uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
      const uint32_t f = mix32(key + ((unsigned long long)round << 32) + r) & mask;
      const uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    x = (l << half_bits) | r;
  } while (x >= n);
  return x;
}

__global__ void __launch_bounds__(256) sample_candidates_kernel(int n_items, int half_bits, int batch, int n_cand,
                                                                unsigned long long seed, const int* __restrict__ users,
                                                                long long* __restrict__ out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)batch * n_cand) return;
  const int b = (int)(t / n_cand), q = (int)(t % n_cand);
  const unsigned long long key = seed ^ ((unsigned long long)(b + 1) * 0x9E3779B97F4A7C15ULL);
  long long* row = out + (size_t)b * (1 + n_cand);
  if (q == 0) row[0] = __ldg(users + b);
  row[1 + q] = feistel_perm((uint32_t)q, half_bits, key, (uint32_t)n_items);
}

// advanced_sampling.py:63-64: min(n_pos, deg) distinct random positives per row (head of a keyed permutation of the
// user's train list), -1 padded.
__global__ void __launch_bounds__(256) sample_positives_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int n_users,
                                                               int batch, int n_pos, unsigned long long seed,
                                                               const int* __restrict__ users, long long* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * n_pos) return;
  const int b = t / n_pos, q = t % n_pos;
  const int u = __ldg(users + b);
  const int lo = __ldg(rowptr + u), deg = __ldg(rowptr + u + 1) - lo;
  long long v = -1;
  if (q < deg) {
    int bits = 2;
    while ((1 << bits) < deg) ++bits;
    if (bits & 1) ++bits;
    const unsigned long long key = seed ^ ((unsigned long long)(b + 1) * 0xD6E8FEB86659FD93ULL);
    v = __ldg(col + lo + (int)feistel_perm((uint32_t)q, bits / 2, key, (uint32_t)deg)) - n_users;
  }
  out[(size_t)b * n_pos + q] = v;
}

}  // namespace tgcn

using namespace tgcn;

extern "C" int tgcn_dropout_mask(int64_t nnz, float dropout, uint64_t seed, uint8_t* d_keep, tgcn_stream_t stream) {
  TGCN_REQUIRE(nnz > 0 && d_keep != nullptr, "bad arguments");
  TGCN_REQUIRE(dropout >= 0.f && dropout < 1.f, "dropout=%f out of [0,1)", dropout);
  TGCN_REQUIRE(((uintptr_t)d_keep & 15) == 0, "keep mask must be 16-byte aligned");
  const double keep_p = 1.0 - (double)dropout;
  const uint32_t thresh = keep_p >= 1.0 ? 0xffffffffu : (uint32_t)(keep_p * 4294967296.0);
  const long long n16 = (nnz + 15) / 16;
  dropout_mask_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n16, nnz, (unsigned long long)seed, thresh, nullptr, d_keep);
  TGCN_CHECK_LAUNCH();
  return 0;
}

namespace tgcn {
__global__ void counter_inc_kernel(unsigned long long* c) { *c += 1ULL; }
}  // namespace tgcn

extern "C" int tgcn_counter_inc(uint64_t* d_counter, tgcn_stream_t stream) {
  TGCN_REQUIRE(d_counter != nullptr, "NULL counter");
  tgcn::counter_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)d_counter);
  TGCN_CHECK_LAUNCH();
  return 0;
}

extern "C" int tgcn_dropout_mask_dev(int64_t nnz, float dropout, uint64_t base_seed, const uint64_t* d_draws, uint8_t* d_keep,
                                     tgcn_stream_t stream) {
  TGCN_REQUIRE(nnz > 0 && d_keep != nullptr && d_draws != nullptr, "bad arguments");
  TGCN_REQUIRE(dropout >= 0.f && dropout < 1.f, "dropout=%f out of [0,1)", dropout);
  TGCN_REQUIRE(((uintptr_t)d_keep & 15) == 0, "keep mask must be 16-byte aligned");
  const double keep_p = 1.0 - (double)dropout;
  const uint32_t thresh = keep_p >= 1.0 ? 0xffffffffu : (uint32_t)(keep_p * 4294967296.0);
  const long long n16 = (nnz + 15) / 16;
  dropout_mask_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n16, nnz, (unsigned long long)base_seed, thresh,
                                                                                      (const unsigned long long*)d_draws, d_keep);
  TGCN_CHECK_LAUNCH();
  return 0;
}

extern "C" int tgcn_sample_positives(const tgcn_graph_t* g, int64_t batch, int32_t n_pos, const int32_t* d_users, uint64_t seed,
                                     int64_t* d_out, tgcn_stream_t stream) {
  TGCN_REQUIRE(g != nullptr && !g->is_block, "sampler needs a whole-graph handle");
  TGCN_REQUIRE(batch > 0 && n_pos > 0 && batch * (int64_t)n_pos < (1ll << 31), "bad sizes");
  TGCN_REQUIRE(d_users && d_out, "NULL argument");
  const int total = (int)(batch * n_pos);
  sample_positives_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g->rowptr, g->col, (int)g->n_users, (int)batch, n_pos,
                                                                                (unsigned long long)seed, d_users, (long long*)d_out);
  TGCN_CHECK_LAUNCH();
  return 0;
}

extern "C" int tgcn_sample_candidates(int64_t n_items, int64_t batch, int32_t n_cand, const int32_t* d_users, uint64_t seed,
                                      int64_t* d_out, tgcn_stream_t stream) {
  TGCN_REQUIRE(n_items > 0 && n_items < (1ll << 31) && batch > 0 && n_cand > 0 && n_cand <= n_items, "bad sizes: n_items=%lld n_cand=%d",
               (long long)n_items, n_cand);
  TGCN_REQUIRE(d_users && d_out, "NULL argument");
  int bits = 2;
  while ((1ll << bits) < n_items) ++bits;
  if (bits & 1) ++bits;  // balanced Feistel halves
  const long long total = batch * (long long)n_cand;
  sample_candidates_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((int)n_items, bits / 2, (int)batch, n_cand,
                                                                                           (unsigned long long)seed, d_users, (long long*)d_out);
  TGCN_CHECK_LAUNCH();
  return 0;
}

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
