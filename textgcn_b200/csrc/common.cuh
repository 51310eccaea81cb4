// Shared helpers for libtgcn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tgcn_b200.h"

namespace tgcn {

void set_error(const char* fmt, ...);

#define TGCN_CHECK_CUDA(expr)                                                                  \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::tgcn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

#define TGCN_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::tgcn::set_error(__VA_ARGS__);    \
      return 2;                          \
    }                                    \
  } while (0)

#define TGCN_CHECK_LAUNCH() TGCN_CHECK_CUDA(cudaGetLastError())

// Debug-assert build (textgcn_b200/build.py --debug -> libtgcn_b200_dbg.so, selected with TGCN_B200_LIB): every index a
// kernel dereferences is range-checked with a device assert.  compute-sanitizer is closed on this GPU pool, so the GPU
// parity suite is run once per round against this build instead (profiles/r02/README.md).  Compiled out otherwise.
#ifdef TGCN_DEBUG_BOUNDS
#include <assert.h>
#define TGCN_DASSERT(cond) assert(cond)
#else
#define TGCN_DASSERT(cond) ((void)0)
#endif

constexpr int kSplitThreshold = 128;  // rows longer than this are cut into segments
constexpr int kSegmentLen = 64;       // nnz per segment of a long row

struct Segment {  // one warp's (or lane group's) share of a long row
  int row;        // local row index
  int begin;      // nnz range
  int end;
  int slot;       // index into the partial-sum scratch
  int split;      // index of the row in the split-row table
  int qbegin;     // the same range in the packed layout (quads of four non-zeros, see tgcn_graph::qptr)
  int qend;
  int pad;
};
struct SplitRow {
  int row;
  int first_slot;
  int n_parts;
  int pad;
};

}  // namespace tgcn

struct tgcn_graph {
  int64_t n_users, n_items;
  int64_t row_begin;  // first global row covered (0 unless a row block)
  int64_t n_rows;     // rows covered by this handle
  int64_t nnz;
  int is_block;
  const int* rowptr;
  const int* col;
  const float* val;
  int* tperm;  // owned, built on demand
  int* qptr;      // owned: packed copy of the CSR for the SpMM kernels — row r owns quads [qptr[r], qptr[r+1]) of four
  int4* qcol;     //        (col, val) pairs, rows padded to whole quads with col = the row's last column / val = 0 (8 B/nnz + padding; NULL when
  float4* qval;   //        switched off with TGCN_SPMM_PACKED=0 at creation or when the allocation failed)
  tgcn::Segment* segments;
  int n_segments;
  tgcn::SplitRow* split_rows;
  int n_split_rows;
  int* split_counters;  // arrivals per split row (self-resetting: ONE launch in flight per handle, enforced by launch_spmm)
  cudaEvent_t busy_event;    // recorded after every launch that uses the counters / partial-sum scratch ...
  cudaStream_t busy_stream;  // ... on this stream: a launch on ANOTHER stream while it is pending is refused
  int busy_valid;
  int n_user_segments;       // segments [0, n_user_segments) belong to user rows (whole-graph handles; all of them for a block)
  cudaStream_t host_in, host_out;  // tgcn_propagate_host: copy streams + events, created on first use
  cudaEvent_t host_ev[13];
  int max_degree;
  int mask_col_off;  // eval masks: entry value = mask_col_off + item id (n_users unless overridden)
  int* order;      // owned: rows with <= kSplitThreshold non-zeros, user rows then item rows, each by decreasing
  int n_ordered;   //        number of kUnroll-wide steps (stable), so the lane groups sharing a warp finish together
  int bipartite;  // verified at creation: user rows reference only item columns and vice versa
};

namespace tgcn {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void fma4(float4& a, float s, const float4& x) {
  a.x = fmaf(s, x.x, a.x);
  a.y = fmaf(s, x.y, a.y);
  a.z = fmaf(s, x.z, a.z);
  a.w = fmaf(s, x.w, a.w);
}
__device__ __forceinline__ void add4(float4& a, const float4& x) {
  a.x += x.x;
  a.y += x.y;
  a.z += x.z;
  a.w += x.w;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// strict total order used for every ranking: score descending, then id ascending
__device__ __forceinline__ bool ranks_before(float sa, int ia, float sb, int ib) {
  return sa > sb || (sa == sb && ia < ib);
}
// true if `key` occurs in the sorted range col[lo, hi)
__device__ __forceinline__ bool sorted_contains(const int* __restrict__ col, int lo, int hi, int key) {
  const int end = hi;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    int c = __ldg(col + mid);
    if (c < key) lo = mid + 1;
    else hi = mid;
  }
  return lo < end && __ldg(col + lo) == key;
}

// What the device-gated second pass of the screened eval path ranks (eval_tc.cu): the first *count entries of the queue `rows`
// (rank positions whose certificate failed).
struct TcGate {
  const int* count;
  const int* rows;
  const int* users;  // user id of queue entry j (mask rows)
  const int* src;    // source row of queue entry j in the user table
};

}  // namespace tgcn
