// K-layer LightGCN propagation: CSR SpMM with 128-bit gathers, edge dropout applied as a keep-mask over
// the static CSR, and the layer mean fused into the last pass.
//
// Roofline: HBM-bound.  Algorithmic bytes per layer = nnz·(4d + 8) + N·4d + (N+1)·4.
//
// Work decomposition (spmm_group_kernel, every width the models use): one GROUP of LPN lanes per row, each lane
// owning VPL float4 of the d-wide output row (d = 4·LPN·VPL) and walking the row's non-zeros itself — no shuffles, no
// cross-lane reduction; several rows share a warp when LPN < 32.  Rows longer than kSplitThreshold are cut into
// kSegmentLen-nnz segments handled by separate groups that write partial sums; the last segment of a row to finish
// adds the partials in slot order, so results are deterministic run to run.  spmm_rows_kernel (warp per row, lanes
// split over non-zeros) serves the odd widths.
#include <stdlib.h>

#include "common.cuh"

namespace tgcn {

constexpr int kMaxAddends = TGCN_MAX_LAYERS + 1;
constexpr int kBatch = 8;  // non-zeros in flight per warp per step

struct Epilogue {
  int n_add;                          // y = (add[0] + add[1] + ... + acc) / divisor  (+ y_old if accumulate)
  const float* add_user[kMaxAddends];  // address of row R: R < add_split ? add_user + R·d : add_item + (R - add_split)·d
  const float* add_item[kMaxAddends];
  int add_split;
  float divisor;
  int accumulate;
  // Feature-sliced multi-GPU mode (tgcn_propagate_sliced): this launch computes columns [col_off, col_off + d) of
  // d_full-wide rows and stores them straight into the row-sharded, full-width result tables of the peer GPUs
  // (peer memory over NVLink): user row r goes to its owner r / users_per_rank, item rows go to every peer.
  int n_peers;  // 0 = plain store into y
  int row_base;
  int user_row0;  // user u is stored at row (u - user_row0) % users_per_rank of peer (u - user_row0) / users_per_rank
  int users_per_rank;
  int n_users;
  int d_full;
  int col_off;
  float* peer_user[TGCN_MAX_PEERS];
  float* peer_item[TGCN_MAX_PEERS];
};

struct SpmmArgs {
  const int* rowptr;
  const int* col;
  const float* val;
  const int* qptr;      // packed layout (handle-owned, NULL = not built): row r owns quads [qptr[r], qptr[r+1]) ...
  const int4* qcol;     // ... of four column ids (padding past the row end repeats the row's last column) ...
  const float4* qval;   // ... and four values (0 on padding): two 128-bit loads per four non-zeros instead of eight 32-bit ones
  const uint8_t* keep;  // per-nnz keep mask or NULL
  const int* tperm;     // when set, entry p uses keep[tperm[p]] (transposed dropout matrix)
  float keep_div;       // survivors are divided by (1 - p) (generic kernel) ...
  float keep_scale;     // ... or multiplied by 1/(1 - p) (group kernel)
  const float* x_user;  // gather source: c < x_split ? x_user + c·d : x_item + (c - x_split)·d
  const float* x_item;
  int x_split;
  int n_rows;        // rows walked by this launch (all rows of the handle, or a row range: see RowRange)
  int row_lo;        // first row of the range (0 for a whole launch); only used when `order` is NULL
  int n_units;       // row units of the group / slice kernels: n_rows, or the number of short rows when `order` is set
  const int* order;  // short rows (<= kSplitThreshold non-zeros) by decreasing step count, or NULL
  int d;
  unsigned zero;  // always 0: an operand the compiler cannot fold (see TGCN_GATHER_FMA)
  long long dbg_x_rows, dbg_nnz, dbg_rows;  // bounds for the debug-assert build: rows of the gather source, nnz and rows of the handle
  const Segment* segments;
  int n_segments;
  const SplitRow* split_rows;
  int* split_counters;
  float* partial;  // (n_segments, d)
  float* y;        // (n_rows, d)
  Epilogue ep;
};

__device__ __forceinline__ void epilogue_store(const SpmmArgs& a, int row, int chunk, float4 acc) {
  const Epilogue& ep = a.ep;
  const size_t off = (size_t)chunk * 4;
  float4 s = acc;
  if (ep.n_add > 0) {
    const bool lo = row < ep.add_split;
    const size_t r = lo ? (size_t)row : (size_t)(row - ep.add_split);
    s = ldg4((lo ? ep.add_user[0] : ep.add_item[0]) + r * a.d + off);
    for (int t = 1; t < ep.n_add; ++t) add4(s, ldg4((lo ? ep.add_user[t] : ep.add_item[t]) + r * a.d + off));
    add4(s, acc);
  }
  if (ep.divisor != 1.0f) {
    s.x /= ep.divisor;
    s.y /= ep.divisor;
    s.z /= ep.divisor;
    s.w /= ep.divisor;
  }
  if (ep.n_peers > 0) {
    const size_t coff = (size_t)ep.col_off + off;
    const int grow = row + ep.row_base;  // global row id (row blocks start at row_base)
    if (grow < ep.n_users) {
      const int rel = grow - ep.user_row0;
      const int q = rel / ep.users_per_rank;
      TGCN_DASSERT(rel >= 0 && q < ep.n_peers);
      *reinterpret_cast<float4*>(ep.peer_user[q] + (size_t)(rel - q * ep.users_per_rank) * ep.d_full + coff) = s;
    } else {
      const size_t o = (size_t)(grow - ep.n_users) * ep.d_full + coff;
      for (int q = 0; q < ep.n_peers; ++q) *reinterpret_cast<float4*>(ep.peer_item[q] + o) = s;
    }
    return;
  }
  float4* out = reinterpret_cast<float4*>(a.y + (size_t)row * a.d + off);
  if (ep.accumulate) add4(s, *out);
  *out = s;
}

// Long rows: every segment writes its partial sum, then signals; the LAST segment of a row to arrive adds the partials
// in slot order (deterministic) and applies the epilogue, so no second kernel is needed.  `leader` is one lane of the
// group, `mask` the group's lanes.  Pattern: store, __threadfence, atomic; last arriver fences and reads through L2.
__device__ __forceinline__ bool segment_arrive_is_last(const SpmmArgs& a, int split, bool leader, unsigned mask, int leader_lane) {
  __threadfence();
  int old = 0;
  const int n_parts = a.split_rows[split].n_parts;
  if (leader) old = atomicAdd(a.split_counters + split, 1);
  old = __shfl_sync(mask, old, leader_lane);
  if (old != n_parts - 1) return false;
  __threadfence();
  if (leader) a.split_counters[split] = 0;  // re-arm for the next launch on this handle
  return true;
}
__device__ __forceinline__ float4 sum_partials(const SpmmArgs& a, int split, int chunk) {
  const SplitRow sr = a.split_rows[split];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int part = 0; part < sr.n_parts; ++part)
    add4(acc, __ldcg(reinterpret_cast<const float4*>(a.partial + (size_t)(sr.first_slot + part) * a.d + chunk * 4)));
  return acc;
}

template <int LPN, int VPL>
__global__ void __launch_bounds__(256) spmm_rows_kernel(const SpmmArgs a) {
  constexpr int G = 32 / LPN;
  constexpr int U = (kBatch / G) > 0 ? (kBatch / G) : 1;
  const int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  int row, begin, end, slot, split = -1;
  if (warp < a.n_segments) {
    const Segment s = a.segments[warp];
    row = s.row;
    begin = s.begin;
    end = s.end;
    slot = s.slot;
    split = s.split;
  } else {
    row = warp - a.n_segments;
    if (row >= a.n_rows) return;
    row += a.row_lo;
    begin = __ldg(a.rowptr + row);
    end = __ldg(a.rowptr + row + 1);
    slot = -1;
    if (end - begin > kSplitThreshold) return;  // handled as segments
  }
  const int grp = lane / LPN, sub = lane % LPN;
  const int d4 = a.d >> 2;
  float4 acc[VPL];
#pragma unroll
  for (int w = 0; w < VPL; ++w) acc[w] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int base = begin; base < end; base += 32) {
    const int p = base + lane;
    int c = 0;
    float v = 0.f;
    bool live = p < end;
    if (live) {
      c = __ldg(a.col + p);
      v = __ldg(a.val + p);
    }
    int cnt = min(32, end - base);
    if (a.keep != nullptr) {
      if (live) {
        const int q = a.tperm ? __ldg(a.tperm + p) : p;
        live = __ldg(a.keep + q) != 0;
        v = v / a.keep_div;
      }
      // compact the surviving entries to the low lanes so no gather slot is wasted
      const unsigned m = __ballot_sync(0xffffffffu, live);
      cnt = __popc(m);
      const unsigned src = __fns(m, 0, lane + 1);
      c = __shfl_sync(0xffffffffu, c, src & 31);
      v = __shfl_sync(0xffffffffu, v, src & 31);
    }
    for (int j = 0; j < cnt; j += G * U) {
      float4 xs[U][VPL];
      float vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = j + u * G + grp;
        const int cc = __shfl_sync(0xffffffffu, c, idx & 31);
        vv[u] = __shfl_sync(0xffffffffu, v, idx & 31);
        const bool ok = idx < cnt;
        if (!ok) vv[u] = 0.f;
        TGCN_DASSERT(!ok || (cc >= 0 && cc < a.dbg_x_rows));
        const float* xr = (cc < a.x_split ? a.x_user + (size_t)cc * a.d : a.x_item + (size_t)(cc - a.x_split) * a.d);
#pragma unroll
        for (int w = 0; w < VPL; ++w) {
          const int chunk = sub + w * LPN;
          xs[u][w] = (ok && chunk < d4) ? ldg4(xr + chunk * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int w = 0; w < VPL; ++w) fma4(acc[w], vv[u], xs[u][w]);
    }
  }
  // combine the G groups
#pragma unroll
  for (int o = LPN; o < 32; o <<= 1) {
#pragma unroll
    for (int w = 0; w < VPL; ++w) {
      acc[w].x += __shfl_xor_sync(0xffffffffu, acc[w].x, o);
      acc[w].y += __shfl_xor_sync(0xffffffffu, acc[w].y, o);
      acc[w].z += __shfl_xor_sync(0xffffffffu, acc[w].z, o);
      acc[w].w += __shfl_xor_sync(0xffffffffu, acc[w].w, o);
    }
  }
  if (grp != 0) return;
#pragma unroll
  for (int w = 0; w < VPL; ++w) {
    const int chunk = sub + w * LPN;
    if (chunk >= d4) continue;
    if (slot >= 0) *reinterpret_cast<float4*>(a.partial + (size_t)slot * a.d + chunk * 4) = acc[w];
    else epilogue_store(a, row, chunk, acc[w]);
  }
  if (slot >= 0) {
    const unsigned mask = LPN == 32 ? 0xffffffffu : ((1u << LPN) - 1u);
    if (segment_arrive_is_last(a, split, sub == 0, mask, 0)) {
#pragma unroll
      for (int w = 0; w < VPL; ++w) {
        const int chunk = sub + w * LPN;
        if (chunk < d4) epilogue_store(a, row, chunk, sum_partials(a, split, chunk));
      }
    }
  }
}

// ---- main kernel (exact widths d = 4·LPN·VPL) -----------------------------------------------------------
// One GROUP of LPN lanes per row (two rows per warp at d = 64, four at d = 32, one at d >= 128): each lane owns
// VPL float4 columns of the output row and walks the row's non-zeros itself, kUnroll at a time, so there are
// no shuffles, no cross-lane reduction and the address arithmetic is a compile-time-constant multiply.
// col/val reads are group-uniform (one broadcast transaction, L1-resident after the first touch of a line).
// Ncu on the first (warp-per-row + shuffle) version showed 30 warp instructions per non-zero, 75 % of them
// integer/address work, and 39 % achieved occupancy from 256-thread blocks waiting on their longest row; this
// layout issues ~5 per non-zero and uses 128-thread blocks.
constexpr int kUnroll = 4;
#ifndef TGCN_DEFAULT_LANES_D64
#define TGCN_DEFAULT_LANES_D64 16
#endif
#ifndef TGCN_DEFAULT_QUADS_D64
#define TGCN_DEFAULT_QUADS_D64 1
#endif
constexpr int kGroupThreads = 128;

template <int LPN, int VPL, int QU, bool PACKED, bool MASKED>
__global__ void __launch_bounds__(kGroupThreads) spmm_group_kernel(const SpmmArgs a) {
  constexpr int D = 4 * LPN * VPL;
  const int tid = blockIdx.x * kGroupThreads + threadIdx.x;
  const int unit = tid / LPN, sub = tid % LPN;
  constexpr bool packed = PACKED;
  int row, begin, end, slot, split = -1;
  int qb = 0, qe = 0;  // packed: the quads of this unit
  if (unit < a.n_segments) {
    const Segment s = a.segments[unit];
    row = s.row;
    begin = s.begin;
    end = s.end;
    slot = s.slot;
    split = s.split;
    qb = s.qbegin;
    qe = s.qend;
  } else {
    row = unit - a.n_segments;
    if (row >= (LPN < 32 ? a.n_units : a.n_rows)) return;
    // several rows share a warp when LPN < 32: `order` lists the short rows by decreasing step count so that
    // the lane groups of a warp finish together (ncu: 16 of 32 lanes active on the 16-wide slice without it)
    if (LPN < 32 && a.order) row = __ldg(a.order + row);
    else row += a.row_lo;
    slot = -1;
    if (packed) {
      qb = __ldg(a.qptr + row);
      qe = __ldg(a.qptr + row + 1);
      if (qe - qb > kSplitThreshold / 4) return;  // handled as segments (deg > 128 <=> more than 32 quads)
      begin = end = 0;
      if (MASKED) begin = __ldg(a.rowptr + row);  // mask bytes are indexed by the ORIGINAL nnz position
    } else {
      begin = __ldg(a.rowptr + row);
      end = __ldg(a.rowptr + row + 1);
      if (end - begin > kSplitThreshold) return;  // handled as segments
    }
  }
  TGCN_DASSERT(row >= 0 && row < a.dbg_rows && qb <= qe && begin <= a.dbg_nnz);
  // Â is bipartite: user rows only reference item columns and vice versa, so the table select happens once
  // per row instead of once per gather (launch_spmm rejects a non-bipartite graph with split tables).
  const float* xb = ((a.x_split != 0x7fffffff && row < a.x_split) ? a.x_item - (size_t)a.x_split * D : a.x_user) + sub * 4;
  float4 acc[VPL];
#pragma unroll
  for (int w = 0; w < VPL; ++w) acc[w] = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr bool masked = MASKED;
  constexpr int NZ = 4 * QU;  // non-zeros in flight per lane and step
  // dropped edges (c < 0: no gather at all) and the 1/(1-p) rescale; p = original position of c[0]
  auto apply_mask = [&](int p, int (&c)[NZ], float (&v)[NZ]) {
#pragma unroll
    for (int i = 0; i < NZ; ++i) {
      if (c[i] >= 0) {
        TGCN_DASSERT(p + i >= 0 && p + i < a.dbg_nnz);
        const int q = a.tperm ? __ldg(a.tperm + p + i) : p + i;
        TGCN_DASSERT(q >= 0 && q < a.dbg_nnz);
        if (__ldg(a.keep + q) == 0) c[i] = -1;
        v[i] *= a.keep_scale;
      }
    }
  };
// every gather of a step is issued before the first FMA of the step (NZ independent 128-bit loads in flight per lane); the
// FMAs then run strictly in nnz order, so the per-row summation order never depends on the layout.  Written out in the loop
// bodies (not as a lambda): ptxas otherwise interleaves load / FMA pairs through one register quad and serialises the gathers.
  // packed + unmasked: the padding of a row repeats its last column with value 0, so every gather is unconditional
  constexpr bool kSkip = !PACKED || MASKED;
#define TGCN_GATHER_FMA(c, v)                                                                                          \
  do {                                                                                                                 \
    float4 x[NZ][VPL];                                                                                                 \
    _Pragma("unroll") for (int i = 0; i < NZ; ++i) TGCN_DASSERT(c[i] >= (kSkip ? -1 : 0) && c[i] < a.dbg_x_rows);      \
    _Pragma("unroll") for (int i = 0; i < NZ; ++i) _Pragma("unroll") for (int w = 0; w < VPL; ++w)                     \
        x[i][w] = (kSkip && c[i] < 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : ldg4(xb + (size_t)c[i] * D + w * (LPN * 4)); \
    /* ptxas otherwise threads load / FMA pairs through one register quad and keeps 1-2 gathers in flight: make the   \
       first weight depend on EVERY gathered row (OR with `all & a.zero`, a.zero == 0 at run time), so all NZ·VPL loads \
       are issued before the first FMA.  Three logic instructions per step; values are unchanged. */                   \
    unsigned all_bits = 0xffffffffu;                                                                                   \
    _Pragma("unroll") for (int i = 0; i < NZ; ++i) _Pragma("unroll") for (int w = 0; w < VPL; ++w)                     \
        all_bits &= __float_as_uint(x[i][w].x);                                                                        \
    float v0 = __uint_as_float(__float_as_uint(v[0]) | (all_bits & a.zero));                                           \
    _Pragma("unroll") for (int i = 0; i < NZ; ++i) _Pragma("unroll") for (int w = 0; w < VPL; ++w)                     \
        fma4(acc[w], i == 0 ? v0 : v[i], x[i][w]);                                                                     \
  } while (0)
  if (packed) {
    // ncu on the 32-bit col/val loads at d = 64: L1/TEX was the busiest unit (62 %) and half of its wavefronts were the
    // group-uniform col/val words; one int4 + one float4 per four non-zeros cuts those eightfold and drops the per-element
    // bounds predicates (padding carries c = -1).
#pragma unroll 1
    for (int q = qb; q < qe; q += QU) {
      int c[NZ];
      float v[NZ];
#pragma unroll
      for (int u = 0; u < QU; ++u) {
        const int qq = (QU == 1 || q + u < qe) ? q + u : q;  // odd tail of a two-quad step: re-read quad q with zero weights
        const int4 cc = __ldg(a.qcol + qq);
        float4 vv = __ldg(a.qval + qq);
        if (QU > 1 && qq != q + u) vv = make_float4(0.f, 0.f, 0.f, 0.f);
        c[4 * u + 0] = cc.x, c[4 * u + 1] = cc.y, c[4 * u + 2] = cc.z, c[4 * u + 3] = cc.w;
        v[4 * u + 0] = vv.x, v[4 * u + 1] = vv.y, v[4 * u + 2] = vv.z, v[4 * u + 3] = vv.w;
      }
      if (masked) {
#pragma unroll
        for (int i = 0; i < NZ; ++i)
          if (v[i] == 0.f) c[i] = -1;  // padding (real entries of Â are positive): no mask byte belongs to it
        apply_mask(begin + 4 * (q - qb), c, v);
      }
      TGCN_GATHER_FMA(c, v);
    }
  } else {
    // (Prefetching step p + 1's col/val ahead of step p's gathers was measured slower everywhere: 48 registers instead
    // of 40 cost more occupancy than the shorter dependent chain returned — c5 120.2 -> 124.5 ms, c2 0.371 -> 0.392 ms.)
    for (int p = begin; p < end; p += NZ) {
      int c[NZ];
      float v[NZ];
#pragma unroll
      for (int i = 0; i < NZ; ++i) {
        const bool ok = p + i < end;
        c[i] = ok ? __ldg(a.col + p + i) : -1;
        v[i] = ok ? __ldg(a.val + p + i) : 0.f;
      }
      if (masked) apply_mask(p, c, v);
      TGCN_GATHER_FMA(c, v);
    }
  }
#undef TGCN_GATHER_FMA
#pragma unroll
  for (int w = 0; w < VPL; ++w) {
    const int chunk = sub + w * LPN;
    if (slot >= 0) *reinterpret_cast<float4*>(a.partial + (size_t)slot * D + chunk * 4) = acc[w];
    else epilogue_store(a, row, chunk, acc[w]);
  }
  if (slot >= 0) {
    const int lane = threadIdx.x & 31;
    const int first = lane - sub;  // first lane of this group inside the warp
    const unsigned mask = LPN == 32 ? 0xffffffffu : (((1u << LPN) - 1u) << first);
    if (segment_arrive_is_last(a, split, sub == 0, mask, first)) {
#pragma unroll
      for (int w = 0; w < VPL; ++w) {
        const int chunk = sub + w * LPN;
        epilogue_store(a, row, chunk, sum_partials(a, split, chunk));
      }
    }
  }
}

// keep_t[p] = keep[tperm[p]]: the dropout mask in the order of Â_dropᵀ's entries, gathered ONCE per backward so that the L
// transposed passes read their mask bytes sequentially instead of through a dependent 4-byte + random 1-byte load per non-zero
// (ncu, c2: 0.20 ms per transposed pass against 0.14 ms for the forward masked pass).
__global__ void __launch_bounds__(256) permute_mask_kernel(const uint8_t* __restrict__ keep, const int* __restrict__ tperm, int64_t nnz,
                                                           uint8_t* __restrict__ keep_t) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p < nnz) keep_t[p] = __ldg(keep + __ldg(tperm + p));
}

struct MeanArgs {
  int n_add;
  const float* add[kMaxAddends];
  float divisor;
};

// out = (add[0] + add[1] + ...) / divisor over n4 float4 — the layer mean for tables that are not produced by a
// local SpMM (the all-reduced item table of the bipartite multi-GPU scheme).  Same summation order as the epilogue.
__global__ void __launch_bounds__(256) layer_mean_kernel(const MeanArgs m, int64_t n4, float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 s = ldg4(m.add[0] + i * 4);
  for (int t = 1; t < m.n_add; ++t) add4(s, ldg4(m.add[t] + i * 4));
  if (m.divisor != 1.0f) {
    s.x /= m.divisor;
    s.y /= m.divisor;
    s.z /= m.divisor;
    s.w /= m.divisor;
  }
  *reinterpret_cast<float4*>(out + i * 4) = s;
}

struct MeanScatterArgs {
  MeanArgs m;
  int ds4;       // float4 per source row
  int d_full;    // destination row stride (floats)
  int col_off;   // destination column (floats)
  int64_t row0;  // destination row of source row 0
  int n_dst;
  float* dst[TGCN_MAX_PEERS];
};

// Same mean over (rows, ds) tables, stored at column col_off of the d_full-wide rows [row0, row0 + rows) of n_dst
// destination tables (the peers' replicated items_emb): the all-gather of the feature-sliced scheme as plain stores.
__global__ void __launch_bounds__(256) layer_mean_scatter_kernel(const MeanScatterArgs a, int64_t n4) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 s = ldg4(a.m.add[0] + i * 4);
  for (int t = 1; t < a.m.n_add; ++t) add4(s, ldg4(a.m.add[t] + i * 4));
  if (a.m.divisor != 1.0f) {
    s.x /= a.m.divisor;
    s.y /= a.m.divisor;
    s.z /= a.m.divisor;
    s.w /= a.m.divisor;
  }
  const int64_t r = i / a.ds4;
  const int c4 = (int)(i - r * a.ds4);
  const size_t o = (size_t)(a.row0 + r) * a.d_full + a.col_off + c4 * 4;
  for (int q = 0; q < a.n_dst; ++q) *reinterpret_cast<float4*>(a.dst[q] + o) = s;
}

template <int LPN, int VPL>
static void launch_group(const SpmmArgs& a, cudaStream_t s, int quads) {
  const int64_t units = (int64_t)a.n_segments + (LPN < 32 ? a.n_units : a.n_rows);
  const int64_t blocks = (units * LPN + kGroupThreads - 1) / kGroupThreads;
  const bool packed = a.qptr != nullptr, masked = a.keep != nullptr;
  constexpr int Q2 = VPL == 1 ? 2 : 1;  // two quads per step only for the one-float4-per-lane layouts (register budget)
  const unsigned nb = (unsigned)blocks;
#define TGCN_GROUP_LAUNCH(QU_, P_, M_) spmm_group_kernel<LPN, VPL, QU_, P_, M_><<<nb, kGroupThreads, 0, s>>>(a)
  if (quads == 2 && Q2 == 2) {
    if (packed && masked) TGCN_GROUP_LAUNCH(Q2, true, true);
    else if (packed) TGCN_GROUP_LAUNCH(Q2, true, false);
    else if (masked) TGCN_GROUP_LAUNCH(Q2, false, true);
    else TGCN_GROUP_LAUNCH(Q2, false, false);
  } else {
    if (packed && masked) TGCN_GROUP_LAUNCH(1, true, true);
    else if (packed) TGCN_GROUP_LAUNCH(1, true, false);
    else if (masked) TGCN_GROUP_LAUNCH(1, false, true);
    else TGCN_GROUP_LAUNCH(1, false, false);
  }
#undef TGCN_GROUP_LAUNCH
}

// Lane layout per width, read ONCE per process (A/B switches for profiling; the defaults are the measured best):
// TGCN_SPMM_LANES_D64 / _D128 = lanes per row (16 | 8 at d = 64, 32 | 16 at d = 128); fewer lanes = more rows per warp, two
// float4 per lane, half the (group-redundant) col / val / address instructions per non-zero.
struct SpmmTuning {
  int lanes_d64, lanes_d128;
  int quads_d64, quads_d128;  // quads (4 non-zeros) in flight per lane and step: TGCN_SPMM_QUADS_D64 / _D128 = 1 | 2
};
static const SpmmTuning& spmm_tuning() {
  static const SpmmTuning t = [] {
    SpmmTuning v{TGCN_DEFAULT_LANES_D64, 32, TGCN_DEFAULT_QUADS_D64, 1};
    if (const char* e = getenv("TGCN_SPMM_LANES_D64")) v.lanes_d64 = atoi(e) == 8 ? 8 : 16;
    if (const char* e = getenv("TGCN_SPMM_LANES_D128")) v.lanes_d128 = atoi(e) == 16 ? 16 : 32;
    if (const char* e = getenv("TGCN_SPMM_QUADS_D64")) v.quads_d64 = atoi(e) == 2 ? 2 : 1;
    if (const char* e = getenv("TGCN_SPMM_QUADS_D128")) v.quads_d128 = atoi(e) == 2 ? 2 : 1;
    return v;
  }();
  return t;
}

// The long-row scratch (partial sums in the caller's workspace, arrival counters in the handle) is per HANDLE, so two
// launches on one handle may only overlap if they are ordered on one stream.  Same-stream launches are; a launch on a
// different stream while the previous one has not finished is refused.  (Skipped during stream capture: a captured
// graph replays with the dependencies it was captured with.)
static int handle_guard_enter(const tgcn_graph* g, cudaStream_t s, bool* record) {
  *record = false;
  if (g->n_segments == 0) return 0;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    (void)cudaGetLastError();
    return 0;
  }
  if (g->busy_valid && g->busy_stream != s)
    TGCN_REQUIRE(cudaEventQuery(g->busy_event) != cudaErrorNotReady,
                 "graph handle is busy on another stream: launches sharing a handle must be ordered on one stream");
  *record = true;
  return 0;
}
static void handle_guard_exit(const tgcn_graph* cg, cudaStream_t s) {
  tgcn_graph* g = const_cast<tgcn_graph*>(cg);
  if (!g->busy_event && cudaEventCreateWithFlags(&g->busy_event, cudaEventDisableTiming) != cudaSuccess) return;
  if (cudaEventRecord(g->busy_event, s) == cudaSuccess) {
    g->busy_stream = s;
    g->busy_valid = 1;
  }
}

static int launch_spmm(const tgcn_graph* g, SpmmArgs& a, cudaStream_t s) {
  const int64_t warps = (int64_t)a.n_segments + a.n_rows;
  const int threads = 256;
  const int64_t blocks = (warps * 32 + threads - 1) / threads;
  TGCN_REQUIRE(blocks < (1ll << 31), "grid too large");
  bool record = false;
  if (int rc = handle_guard_enter(g, s, &record)) return rc;
  const int d = a.d;
  const bool contiguous = a.x_split == 0x7fffffff || a.x_item == a.x_user + (size_t)a.x_split * d;
  const bool group_ok = g->bipartite || contiguous;  // per-row table select needs a bipartite Â or one table
  const SpmmTuning& tune = spmm_tuning();
  if (group_ok && d == 16) launch_group<4, 1>(a, s, 1);
  else if (group_ok && d == 32) launch_group<8, 1>(a, s, tune.quads_d64);
  else if (group_ok && d == 64 && tune.lanes_d64 == 8) launch_group<8, 2>(a, s, 1);
  else if (group_ok && d == 64) launch_group<16, 1>(a, s, tune.quads_d64);
  else if (group_ok && d == 128 && tune.lanes_d128 == 16) launch_group<16, 2>(a, s, 1);
  else if (group_ok && d == 128) launch_group<32, 1>(a, s, tune.quads_d128);
  else if (group_ok && d == 256) launch_group<32, 2>(a, s, 1);
  else if (d <= 16) spmm_rows_kernel<4, 1><<<(unsigned)blocks, threads, 0, s>>>(a);
  else if (d <= 32) spmm_rows_kernel<8, 1><<<(unsigned)blocks, threads, 0, s>>>(a);
  else if (d <= 64) spmm_rows_kernel<16, 1><<<(unsigned)blocks, threads, 0, s>>>(a);
  else if (d <= 128) spmm_rows_kernel<32, 1><<<(unsigned)blocks, threads, 0, s>>>(a);
  else if (d <= 256) spmm_rows_kernel<32, 2><<<(unsigned)blocks, threads, 0, s>>>(a);
  else if (d <= 512) spmm_rows_kernel<32, 4><<<(unsigned)blocks, threads, 0, s>>>(a);
  else TGCN_REQUIRE(false, "embedding width %d > 512 is not supported", d);
  TGCN_CHECK_LAUNCH();
  if (record) handle_guard_exit(g, s);
  return 0;
}

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static int check_common(const tgcn_graph* g, int64_t d, int32_t n_layers) {
  TGCN_REQUIRE(g != nullptr, "graph is NULL");
  TGCN_REQUIRE(d > 0 && d % 4 == 0 && d <= 512, "embedding width d=%lld must be a multiple of 4 and <= 512", (long long)d);
  TGCN_REQUIRE(n_layers >= 0 && n_layers <= TGCN_MAX_LAYERS, "n_layers=%d out of range [0, %d]", n_layers, TGCN_MAX_LAYERS);
  return 0;
}

static void base_args(const tgcn_graph* g, int64_t d, SpmmArgs& a) {
  a.rowptr = g->rowptr;
  a.col = g->col;
  a.val = g->val;
  a.qptr = g->qptr;
  a.qcol = g->qcol;
  a.qval = g->qval;
  a.keep = nullptr;
  a.tperm = nullptr;
  a.keep_div = 1.f;
  a.keep_scale = 1.f;
  a.n_rows = (int)g->n_rows;
  a.zero = 0u;
  a.dbg_rows = g->n_rows;
  a.dbg_nnz = g->nnz;
  a.dbg_x_rows = g->is_block ? (1ll << 31) : g->n_users + g->n_items;  // a row block gathers from a caller-sized table
  a.row_lo = 0;
  a.order = g->order;  // NULL when the handle was created with the row order switched off (TGCN_ROW_ORDER=0 at creation)
  a.n_units = g->order ? g->n_ordered : (int)g->n_rows;
  a.d = (int)d;
  a.segments = g->segments;
  a.n_segments = g->n_segments;
  a.split_rows = g->split_rows;
  a.split_counters = g->split_counters;
  a.ep.n_add = 0;
  a.ep.add_split = g->is_block ? 0x7fffffff : (int)g->n_users;
  a.ep.divisor = 1.f;
  a.ep.accumulate = 0;
  a.ep.n_peers = 0;
  a.x_split = g->is_block ? 0x7fffffff : (int)g->n_users;
}

struct RowRange {  // a launch restricted to rows [row_lo, row_hi) (natural order) + the long-row segments [seg_lo, seg_hi)
  int row_lo, row_hi, seg_lo, seg_hi;
};

struct ScatterSpec {  // destination of the last pass in feature-sliced mode (see Epilogue)
  int n_peers, users_per_rank, d_full, col_off;
  int64_t user_row0;  // global id of the user stored at local row 0 of peer 0
  float* const* peer_user;
  float* const* peer_item;
};

}  // namespace tgcn

using namespace tgcn;

static int spmm_ex_impl(const tgcn_graph_t* g, int64_t d, const float* d_x_user, const float* d_x_item,
                        const uint8_t* d_keep, float dropout, int32_t transposed, int32_t n_add,
                        const float* const* h_add_user, const float* const* h_add_item, float divisor,
                        int32_t accumulate, float* d_y, void* d_workspace, int64_t workspace_bytes, tgcn_stream_t stream,
                        const ScatterSpec* sc, const RowRange* range = nullptr);

extern "C" {

int64_t tgcn_propagate_workspace_bytes(const tgcn_graph_t* g, int64_t d, int32_t n_layers) {
  if (!g || d <= 0) return -1;
  const int64_t bufs = n_layers > 1 ? n_layers - 1 : 0;  // fwd keeps E_1..E_{L-1}; bwd ping-pongs within them
  const int64_t layer = align_up(g->n_rows * d * (int64_t)sizeof(float), 256);
  // layer buffers | long-row partial sums | the transposed dropout mask of propagate_bwd (one byte per nnz)
  return bufs * layer + align_up((int64_t)g->n_segments * d * sizeof(float), 256) + align_up(g->nnz, 256) + 256;
}

// Operator-level building block: one SpMM pass with the fused epilogue.  Used by the multi-GPU host
// code (one call per hop between all-gathers) and by the entry points below.
int tgcn_spmm_ex(const tgcn_graph_t* g, int64_t d, const float* d_x_user, const float* d_x_item,
                 const uint8_t* d_keep, float dropout, int32_t transposed, int32_t n_add,
                 const float* const* h_add_user, const float* const* h_add_item, float divisor,
                 int32_t accumulate, float* d_y, void* d_workspace, int64_t workspace_bytes, tgcn_stream_t stream) {
  return spmm_ex_impl(g, d, d_x_user, d_x_item, d_keep, dropout, transposed, n_add, h_add_user, h_add_item, divisor,
                      accumulate, d_y, d_workspace, workspace_bytes, stream, nullptr);
}

}  // extern "C"

static int spmm_ex_impl(const tgcn_graph_t* g, int64_t d, const float* d_x_user, const float* d_x_item,
                        const uint8_t* d_keep, float dropout, int32_t transposed, int32_t n_add,
                        const float* const* h_add_user, const float* const* h_add_item, float divisor,
                        int32_t accumulate, float* d_y, void* d_workspace, int64_t workspace_bytes, tgcn_stream_t stream,
                        const ScatterSpec* sc, const RowRange* range) {
  if (int rc = check_common(g, d, 1)) return rc;
  TGCN_REQUIRE(d_x_user && (d_y || sc), "NULL x or y");
  TGCN_REQUIRE(n_add >= 0 && n_add <= kMaxAddends, "n_add=%d out of range", n_add);
  const int64_t need = align_up((int64_t)g->n_segments * d * sizeof(float), 256);
  TGCN_REQUIRE(g->n_segments == 0 || (d_workspace && workspace_bytes >= need), "workspace too small: need %lld bytes", (long long)need);
  SpmmArgs a;
  base_args(g, d, a);
  a.x_user = d_x_user;
  a.x_item = d_x_item ? d_x_item : d_x_user + (size_t)g->n_users * d;
  if (d_keep) {
    TGCN_REQUIRE(dropout >= 0.f && dropout < 1.f, "dropout=%f out of [0,1)", dropout);
    a.keep = d_keep;
    a.keep_div = (float)(1.0 - (double)dropout);
    a.keep_scale = (float)(1.0 / (1.0 - (double)dropout));
    if (transposed) {
      TGCN_REQUIRE(g->tperm != nullptr, "transpose permutation not built: call tgcn_graph_build_transpose_perm");
      a.tperm = g->tperm;
    }
  }
  for (int t = 0; t < n_add; ++t) {
    a.ep.add_user[t] = h_add_user[t];
    a.ep.add_item[t] = (h_add_item && h_add_item[t]) ? h_add_item[t] : h_add_user[t] + (size_t)g->n_users * d;
  }
  a.ep.n_add = n_add;
  a.ep.divisor = divisor;
  a.ep.accumulate = accumulate;
  a.partial = (float*)d_workspace;
  a.y = d_y;
  if (sc) {
    TGCN_REQUIRE(!accumulate, "feature-sliced scatter cannot accumulate");
    TGCN_REQUIRE(sc->n_peers >= 1 && sc->n_peers <= TGCN_MAX_PEERS, "n_peers=%d out of range [1, %d]", sc->n_peers, TGCN_MAX_PEERS);
    {  // every user row of the handle must land inside one of the peer tables
      const int64_t first = g->row_begin - sc->user_row0;
      const int64_t user_rows = g->is_block ? (g->row_begin < g->n_users ? g->n_rows : 0) : g->n_users;
      TGCN_REQUIRE(sc->users_per_rank > 0 && (user_rows == 0 || (first >= 0 && first + user_rows <= (int64_t)sc->users_per_rank * sc->n_peers)),
                   "users_per_rank x n_peers does not cover the user rows of this handle");
    }
    TGCN_REQUIRE(sc->d_full % 4 == 0 && sc->col_off % 4 == 0 && sc->col_off + d <= sc->d_full, "bad column slice");
    a.ep.n_peers = sc->n_peers;
    a.ep.row_base = (int)g->row_begin;
    a.ep.user_row0 = (int)sc->user_row0;
    a.ep.users_per_rank = sc->users_per_rank;
    a.ep.n_users = (int)g->n_users;
    a.ep.d_full = sc->d_full;
    a.ep.col_off = sc->col_off;
    for (int q = 0; q < sc->n_peers; ++q) {
      TGCN_REQUIRE(sc->peer_user[q] && sc->peer_item[q], "NULL peer table %d", q);
      a.ep.peer_user[q] = sc->peer_user[q];
      a.ep.peer_item[q] = sc->peer_item[q];
    }
  }
  if (range) {  // the partial-sum slots and split counters stay indexed globally; only the unit list shrinks
    TGCN_REQUIRE(range->row_lo >= 0 && range->row_lo <= range->row_hi && range->row_hi <= g->n_rows && range->seg_lo >= 0 &&
                     range->seg_lo <= range->seg_hi && range->seg_hi <= g->n_segments, "bad row range");
    a.order = nullptr;
    a.row_lo = range->row_lo;
    a.n_rows = a.n_units = range->row_hi - range->row_lo;
    a.segments = g->segments + range->seg_lo;
    a.n_segments = range->seg_hi - range->seg_lo;
    if (a.n_rows + a.n_segments == 0) return 0;
  }
  return launch_spmm(g, a, (cudaStream_t)stream);
}

extern "C" {

int tgcn_layer_mean(int64_t n, int32_t n_add, const float* const* h_add, float divisor, float* d_out, tgcn_stream_t stream) {
  TGCN_REQUIRE(n > 0 && n % 4 == 0, "n=%lld must be a positive multiple of 4", (long long)n);
  TGCN_REQUIRE(n_add >= 1 && n_add <= kMaxAddends && h_add && d_out, "bad addends");
  MeanArgs m;
  m.n_add = n_add;
  for (int t = 0; t < n_add; ++t) {
    TGCN_REQUIRE(h_add[t] != nullptr, "NULL addend %d", t);
    m.add[t] = h_add[t];
  }
  m.divisor = divisor;
  const int64_t n4 = n / 4;
  layer_mean_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(m, n4, d_out);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int tgcn_layer_mean_scatter(int64_t n_rows, int64_t d_slice, int32_t n_add, const float* const* h_add, float divisor,
                            int64_t d_full, int64_t col_off, int64_t row0, int32_t n_dst, float* const* h_dst,
                            tgcn_stream_t stream) {
  TGCN_REQUIRE(n_rows > 0 && d_slice > 0 && d_slice % 4 == 0, "bad source shape");
  TGCN_REQUIRE(n_add >= 1 && n_add <= kMaxAddends && h_add, "bad addends");
  TGCN_REQUIRE(d_full % 4 == 0 && col_off % 4 == 0 && col_off >= 0 && col_off + d_slice <= d_full && row0 >= 0, "bad destination geometry");
  TGCN_REQUIRE(n_dst >= 1 && n_dst <= TGCN_MAX_PEERS && h_dst, "n_dst=%d out of range [1, %d]", n_dst, TGCN_MAX_PEERS);
  MeanScatterArgs a;
  a.m.n_add = n_add;
  for (int t = 0; t < n_add; ++t) {
    TGCN_REQUIRE(h_add[t] != nullptr, "NULL addend %d", t);
    a.m.add[t] = h_add[t];
  }
  a.m.divisor = divisor;
  a.ds4 = (int)(d_slice / 4);
  a.d_full = (int)d_full;
  a.col_off = (int)col_off;
  a.row0 = row0;
  a.n_dst = n_dst;
  for (int q = 0; q < n_dst; ++q) {
    TGCN_REQUIRE(h_dst[q] != nullptr, "NULL destination %d", q);
    a.dst[q] = h_dst[q];
  }
  const int64_t n4 = n_rows * (d_slice / 4);
  layer_mean_scatter_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, n4);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int tgcn_spmm_scatter(const tgcn_graph_t* g, int64_t d_slice, const float* d_x_user, const float* d_x_item, int32_t n_add,
                      const float* const* h_add_user, const float* const* h_add_item, float divisor, int64_t d_full,
                      int64_t col_off, int32_t n_peers, int64_t users_per_rank, int64_t user_row0,
                      float* const* h_peer_user_out, float* const* h_peer_item_out, void* d_workspace,
                      int64_t workspace_bytes, tgcn_stream_t stream) {
  TGCN_REQUIRE(h_peer_user_out && h_peer_item_out, "NULL peer table list");
  TGCN_REQUIRE(users_per_rank > 0 && users_per_rank < (1ll << 31) && d_full > 0 && d_full <= 4096 && user_row0 >= 0, "bad slice geometry");
  ScatterSpec sc{n_peers, (int)users_per_rank, (int)d_full, (int)col_off, user_row0, h_peer_user_out, h_peer_item_out};
  return spmm_ex_impl(g, d_slice, d_x_user, d_x_item, nullptr, 0.f, 0, n_add, h_add_user, h_add_item, divisor, 0, nullptr,
                      d_workspace, workspace_bytes, stream, &sc);
}

int tgcn_spmm_fwd(const tgcn_graph_t* g, int64_t d, const float* d_x, float* d_y, void* d_workspace,
                  int64_t workspace_bytes, tgcn_stream_t stream) {
  return tgcn_spmm_ex(g, d, d_x, nullptr, nullptr, 0.f, 0, 0, nullptr, nullptr, 1.f, 0, d_y, d_workspace, workspace_bytes, stream);
}

}  // extern "C"

static int propagate_fwd_impl(const tgcn_graph_t* g, int64_t d, int32_t n_layers, int32_t single, const float* d_user_w,
                              const float* d_item_w, const uint8_t* d_keep, float dropout, float* d_out, void* d_workspace,
                              int64_t workspace_bytes, tgcn_stream_t stream, const ScatterSpec* sc) {
  if (int rc = check_common(g, d, n_layers)) return rc;
  TGCN_REQUIRE(!g->is_block, "propagate_fwd needs a whole-graph handle; drive row blocks with tgcn_spmm_ex");
  TGCN_REQUIRE(d_user_w && d_item_w && (d_out || sc), "NULL table pointer");
  TGCN_REQUIRE(!sc || n_layers >= 1, "feature-sliced propagation needs n_layers >= 1");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t N = g->n_rows;
  if (n_layers == 0) {
    TGCN_CHECK_CUDA(cudaMemcpyAsync(d_out, d_user_w, sizeof(float) * g->n_users * d, cudaMemcpyDeviceToDevice, s));
    TGCN_CHECK_CUDA(cudaMemcpyAsync(d_out + g->n_users * d, d_item_w, sizeof(float) * g->n_items * d, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  const int64_t need = tgcn_propagate_workspace_bytes(g, d, n_layers);
  TGCN_REQUIRE(workspace_bytes >= need && (d_workspace || need == 0), "workspace too small: need %lld bytes, got %lld", (long long)need, (long long)workspace_bytes);
  const int64_t layer_bytes = align_up(N * d * (int64_t)sizeof(float), 256);
  char* ws = (char*)d_workspace;
  const int n_bufs = n_layers - 1;
  float* partial = (float*)(ws + (int64_t)n_bufs * layer_bytes);
  auto buf = [&](int i) { return (float*)(ws + (int64_t)i * layer_bytes); };
  const float* add_u[kMaxAddends];
  const float* add_i[kMaxAddends];
  for (int l = 1; l <= n_layers; ++l) {
    const bool last = l == n_layers;
    const float* xu = l == 1 ? d_user_w : buf(l - 2);
    const float* xi = l == 1 ? d_item_w : nullptr;
    int n_add = 0;
    float divisor = 1.f;
    if (last && !single) {
      add_u[0] = d_user_w;
      add_i[0] = d_item_w;
      for (int t = 1; t < n_layers; ++t) {
        add_u[t] = buf(t - 1);
        add_i[t] = nullptr;
      }
      n_add = n_layers;
      divisor = (float)(n_layers + 1);
    }
    float* y = last ? d_out : buf(l - 1);
    if (int rc = spmm_ex_impl(g, d, xu, xi, d_keep, dropout, 0, n_add, add_u, add_i, divisor, 0, y, partial,
                              workspace_bytes - ((char*)partial - ws), stream, last ? sc : nullptr))
      return rc;
  }
  return 0;
}

extern "C" {

int tgcn_propagate_fwd(const tgcn_graph_t* g, int64_t d, int32_t n_layers, int32_t single, const float* d_user_w,
                       const float* d_item_w, const uint8_t* d_keep, float dropout, float* d_out, void* d_workspace,
                       int64_t workspace_bytes, tgcn_stream_t stream) {
  return propagate_fwd_impl(g, d, n_layers, single, d_user_w, d_item_w, d_keep, dropout, d_out, d_workspace, workspace_bytes,
                            stream, nullptr);
}

int tgcn_propagate_sliced(const tgcn_graph_t* g, int64_t d_slice, int32_t n_layers, int32_t single,
                          const float* d_user_slice, const float* d_item_slice, const uint8_t* d_keep, float dropout,
                          int64_t d_full, int64_t col_off, int32_t n_peers, int64_t users_per_rank,
                          float* const* h_peer_user_out, float* const* h_peer_item_out, void* d_workspace,
                          int64_t workspace_bytes, tgcn_stream_t stream) {
  TGCN_REQUIRE(h_peer_user_out && h_peer_item_out, "NULL peer table list");
  TGCN_REQUIRE(users_per_rank > 0 && users_per_rank < (1ll << 31) && d_full > 0 && d_full <= 4096, "bad slice geometry");
  ScatterSpec sc{n_peers, (int)users_per_rank, (int)d_full, (int)col_off, 0, h_peer_user_out, h_peer_item_out};
  return propagate_fwd_impl(g, d_slice, n_layers, single, d_user_slice, d_item_slice, d_keep, dropout, nullptr, d_workspace,
                            workspace_bytes, stream, &sc);
}

int tgcn_propagate_bwd(tgcn_graph_t* g, int64_t d, int32_t n_layers, int32_t single, const float* d_grad_out,
                       const uint8_t* d_keep, float dropout, int32_t accumulate, float* d_grad_in, void* d_workspace,
                       int64_t workspace_bytes, tgcn_stream_t stream) {
  if (int rc = check_common(g, d, n_layers)) return rc;
  TGCN_REQUIRE(!g->is_block, "propagate_bwd needs a whole-graph handle");
  TGCN_REQUIRE(d_grad_out && d_grad_in, "NULL gradient pointer");
  TGCN_REQUIRE(n_layers >= 1, "propagate_bwd needs n_layers >= 1");
  const int64_t N = g->n_rows;
  if (d_keep && !g->tperm) {
    if (int rc = tgcn_graph_build_transpose_perm(g, stream)) return rc;
  }
  const int64_t need = tgcn_propagate_workspace_bytes(g, d, n_layers);
  TGCN_REQUIRE(workspace_bytes >= need && (d_workspace || need == 0), "workspace too small: need %lld bytes, got %lld", (long long)need, (long long)workspace_bytes);
  const int64_t layer_bytes = align_up(N * d * (int64_t)sizeof(float), 256);
  char* ws = (char*)d_workspace;
  const int n_bufs = n_layers - 1;
  float* partial = (float*)(ws + (int64_t)n_bufs * layer_bytes);
  auto buf = [&](int i) { return (float*)(ws + (int64_t)(i & 1) * layer_bytes); };
  const int64_t partial_bytes = align_up((int64_t)g->n_segments * d * sizeof(float), 256);
  const uint8_t* keep = d_keep;
  int transposed = 1;
  if (d_keep && n_layers >= 2) {  // gather the mask into transposed order once; the passes then run like forward ones
    uint8_t* keep_t = (uint8_t*)partial + partial_bytes;
    permute_mask_kernel<<<(unsigned)((g->nnz + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_keep, g->tperm, g->nnz, keep_t);
    TGCN_CHECK_LAUNCH();
    keep = keep_t;
    transposed = 0;
  }
  // Horner: H_0 = G;  H_l = G + Â_dropᵀ·H_{l-1};  dE0 = H_L / (L+1).   single: dE0 = (Âᵀ)^L·G.
  const float* add_u[1] = {d_grad_out};
  const float* h = d_grad_out;
  for (int l = 1; l <= n_layers; ++l) {
    const bool last = l == n_layers;
    float* y = last ? d_grad_in : buf(l - 1);
    const int n_add = single ? 0 : 1;
    const float divisor = (last && !single) ? (float)(n_layers + 1) : 1.f;
    if (int rc = tgcn_spmm_ex(g, d, h, nullptr, keep, dropout, transposed, n_add, add_u, nullptr, divisor,
                              last ? accumulate : 0, y, partial, partial_bytes, stream))
      return rc;
    h = y;
  }
  return 0;
}

// Host-buffer form of `representation`, pipelined over PCIe (full duplex) and three streams:
//   copy-in stream : item table (the small one) first, then the user table in row chunks;
//   `stream`       : layer 1's USER-row pass starts as soon as the item table has landed (user rows only gather item
//                    rows — Â is bipartite) and overlaps the upload of the user table; the item-row pass waits for it;
//                    layers 2..L-1 as usual; the LAST layer runs its item-row pass first, then the user rows in row
//                    chunks (natural row order), each chunk signalling an event;
//   copy-out stream: the item rows of the result, then every user-row chunk as soon as its launch has finished, so the
//                    download overlaps the rest of the last layer.
// `stream` finally waits for the last download, so the call keeps its stream-ordered contract.  The floor is the PCIe
// time of the two tables (they cannot overlap each other: the result needs every layer): see DESIGN.md §4.
constexpr int kHostChunks = 8;

int tgcn_propagate_host(const tgcn_graph_t* cg, int64_t d, int32_t n_layers, int32_t single, const float* h_user_w,
                        const float* h_item_w, float* h_out, float* d_stage, void* d_workspace, int64_t workspace_bytes,
                        tgcn_stream_t stream) {
  if (int rc = check_common(cg, d, n_layers)) return rc;
  TGCN_REQUIRE(h_user_w && h_item_w && h_out && d_stage, "NULL buffer");
  TGCN_REQUIRE(!cg->is_block, "propagate_host needs a whole-graph handle");
  tgcn_graph* g = const_cast<tgcn_graph*>(cg);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t NU = g->n_users, NI = g->n_items, N = NU + NI;
  const int64_t nu = NU * d, ni = NI * d;
  float* e0_u = d_stage;
  float* e0_i = d_stage + nu;
  float* d_out = d_stage + nu + ni;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  TGCN_CHECK_CUDA(cudaStreamIsCapturing(s, &cap));
  if (n_layers == 0 || !g->bipartite || cap != cudaStreamCaptureStatusNone) {  // nothing to overlap / not splittable: serial form
    TGCN_CHECK_CUDA(cudaMemcpyAsync(e0_u, h_user_w, sizeof(float) * nu, cudaMemcpyHostToDevice, s));
    TGCN_CHECK_CUDA(cudaMemcpyAsync(e0_i, h_item_w, sizeof(float) * ni, cudaMemcpyHostToDevice, s));
    if (int rc = tgcn_propagate_fwd(g, d, n_layers, single, e0_u, e0_i, nullptr, 0.f, d_out, d_workspace, workspace_bytes, stream)) return rc;
    TGCN_CHECK_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(float) * (nu + ni), cudaMemcpyDeviceToHost, s));
    return 0;
  }
  if (!g->host_in) {
    TGCN_CHECK_CUDA(cudaStreamCreateWithFlags(&g->host_in, cudaStreamNonBlocking));
    TGCN_CHECK_CUDA(cudaStreamCreateWithFlags(&g->host_out, cudaStreamNonBlocking));
    for (int i = 0; i < kHostChunks + 5; ++i) TGCN_CHECK_CUDA(cudaEventCreateWithFlags(&g->host_ev[i], cudaEventDisableTiming));
  }
  cudaEvent_t ev_start = g->host_ev[kHostChunks], ev_items = g->host_ev[kHostChunks + 1], ev_users = g->host_ev[kHostChunks + 2],
              ev_done = g->host_ev[kHostChunks + 3], ev_items_out = g->host_ev[kHostChunks + 4];
  // ---- upload: ordered behind whatever `stream` did before (the staging buffers may still be in use) ----
  TGCN_CHECK_CUDA(cudaEventRecord(ev_start, s));
  TGCN_CHECK_CUDA(cudaStreamWaitEvent(g->host_in, ev_start, 0));
  TGCN_CHECK_CUDA(cudaMemcpyAsync(e0_i, h_item_w, sizeof(float) * ni, cudaMemcpyHostToDevice, g->host_in));
  TGCN_CHECK_CUDA(cudaEventRecord(ev_items, g->host_in));
  for (int c = 0; c < kHostChunks; ++c) {  // chunked so that the copy engine never holds one multi-GB descriptor
    const int64_t r0 = NU * c / kHostChunks, r1 = NU * (c + 1) / kHostChunks;
    if (r1 > r0) TGCN_CHECK_CUDA(cudaMemcpyAsync(e0_u + r0 * d, h_user_w + r0 * d, sizeof(float) * (r1 - r0) * d, cudaMemcpyHostToDevice, g->host_in));
  }
  TGCN_CHECK_CUDA(cudaEventRecord(ev_users, g->host_in));
  // ---- layers ----
  const int64_t need = tgcn_propagate_workspace_bytes(g, d, n_layers);
  TGCN_REQUIRE(workspace_bytes >= need && (d_workspace || need == 0), "workspace too small: need %lld bytes, got %lld", (long long)need, (long long)workspace_bytes);
  const int64_t layer_bytes = align_up(N * d * (int64_t)sizeof(float), 256);
  char* ws = (char*)d_workspace;
  float* partial = (float*)(ws + (int64_t)(n_layers - 1) * layer_bytes);
  const int64_t partial_bytes = workspace_bytes - ((char*)partial - ws);
  auto buf = [&](int i) { return (float*)(ws + (int64_t)i * layer_bytes); };
  const RowRange users{0, (int)NU, 0, g->n_user_segments}, items{(int)NU, (int)N, g->n_user_segments, g->n_segments};
  const float* add_u[kMaxAddends];
  const float* add_i[kMaxAddends];
  for (int l = 1; l <= n_layers; ++l) {
    const bool last = l == n_layers;
    const float* xu = l == 1 ? e0_u : buf(l - 2);
    const float* xi = l == 1 ? e0_i : nullptr;
    int n_add = 0;
    float divisor = 1.f;
    if (last && !single) {
      add_u[0] = e0_u;
      add_i[0] = e0_i;
      for (int t = 1; t < n_layers; ++t) {
        add_u[t] = buf(t - 1);
        add_i[t] = nullptr;
      }
      n_add = n_layers;
      divisor = (float)(n_layers + 1);
    }
    float* y = last ? d_out : buf(l - 1);
    auto pass = [&](const RowRange& r) {
      return spmm_ex_impl(g, d, xu, xi, nullptr, 0.f, 0, n_add, add_u, add_i, divisor, 0, y, partial, partial_bytes, stream, nullptr, &r);
    };
    if (l == 1) TGCN_CHECK_CUDA(cudaStreamWaitEvent(s, ev_items, 0));
    if (!last) {
      if (int rc = pass(users)) return rc;                       // layer 1: overlaps the user-table upload
      if (l == 1) TGCN_CHECK_CUDA(cudaStreamWaitEvent(s, ev_users, 0));
      if (int rc = pass(items)) return rc;
      continue;
    }
    // last layer: item rows first (their download starts at once), then the user rows chunk by chunk
    if (l == 1) TGCN_CHECK_CUDA(cudaStreamWaitEvent(s, ev_users, 0));
    if (int rc = pass(items)) return rc;
    TGCN_CHECK_CUDA(cudaEventRecord(ev_items_out, s));
    TGCN_CHECK_CUDA(cudaStreamWaitEvent(g->host_out, ev_items_out, 0));
    TGCN_CHECK_CUDA(cudaMemcpyAsync(h_out + nu, d_out + nu, sizeof(float) * ni, cudaMemcpyDeviceToHost, g->host_out));
    for (int c = 0; c < kHostChunks; ++c) {
      const int64_t r0 = NU * c / kHostChunks, r1 = NU * (c + 1) / kHostChunks;
      // the long user rows (segments) ride with chunk 0; they write rows of ANY chunk, all of which are downloaded after
      // chunk 0's launch has finished (same stream, later events)
      const RowRange chunk{(int)r0, (int)r1, c == 0 ? 0 : g->n_user_segments, g->n_user_segments};
      if (int rc = pass(chunk)) return rc;
      TGCN_CHECK_CUDA(cudaEventRecord(g->host_ev[c], s));
      TGCN_CHECK_CUDA(cudaStreamWaitEvent(g->host_out, g->host_ev[c], 0));
      if (r1 > r0) TGCN_CHECK_CUDA(cudaMemcpyAsync(h_out + r0 * d, d_out + r0 * d, sizeof(float) * (r1 - r0) * d, cudaMemcpyDeviceToHost, g->host_out));
    }
  }
  TGCN_CHECK_CUDA(cudaEventRecord(ev_done, g->host_out));
  TGCN_CHECK_CUDA(cudaStreamWaitEvent(s, ev_done, 0));
  return 0;
}

}  // extern "C"
