// Graph handle: borrowed CSR of Â + owned load-balancing metadata and transpose permutation.
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace tgcn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// tperm[p] = q such that entry q is (c, r) when entry p is (r, c).  One thread per nnz.
__global__ void transpose_perm_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int n_rows,
                                      int64_t nnz, int* __restrict__ tperm, int* __restrict__ missing) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  // row of p: largest r with rowptr[r] <= p
  int lo = 0, hi = n_rows;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if ((int64_t)__ldg(rowptr + mid) <= p) lo = mid;
    else hi = mid;
  }
  const int r = lo;
  const int c = __ldg(col + p);
  int a = __ldg(rowptr + c), b = __ldg(rowptr + c + 1);
  const int end = b;
  while (a < b) {
    int mid = (a + b) >> 1;
    if (__ldg(col + mid) < r) a = mid + 1;
    else b = mid;
  }
  if (a < end && __ldg(col + a) == r) {
    tperm[p] = a;
  } else {
    tperm[p] = -1;
    atomicAdd(missing, 1);
  }
}

// Packed copy of the CSR: one thread per quad; its row by binary search in qptr.
__global__ void pack_quads_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ val,
                                  const int* __restrict__ qptr, int n_rows, int n_quads, int4* __restrict__ qcol, float4* __restrict__ qval) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_quads) return;
  int lo = 0, hi = n_rows;  // largest r with qptr[r] <= q
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(qptr + mid) <= q) lo = mid;
    else hi = mid;
  }
  const int p = __ldg(rowptr + lo) + 4 * (q - __ldg(qptr + lo)), end = __ldg(rowptr + lo + 1);
  int c[4];
  float v[4];
  const int last = __ldg(col + end - 1);  // padding repeats the row's last column with weight 0: gathers stay unconditional
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c[i] = p + i < end ? __ldg(col + p + i) : last;
    v[i] = p + i < end ? __ldg(val + p + i) : 0.f;
  }
  qcol[q] = make_int4(c[0], c[1], c[2], c[3]);
  qval[q] = make_float4(v[0], v[1], v[2], v[3]);
}

// Per row: columns strictly increasing (the binary searches rely on it) and, for a whole graph, on the other
// side of the user/item boundary.  flags[0] |= unsorted, flags[1] |= not bipartite.
__global__ void check_rows_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int n_rows, int n_users,
                                  int* __restrict__ flags) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int b = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
  bool unsorted = false;
  for (int p = b + 1; p < e; ++p) unsorted |= __ldg(col + p) <= __ldg(col + p - 1);
  if (unsorted) flags[0] = 1;
  if (e > b) {
    const bool bad = r < n_users ? __ldg(col + b) < n_users : __ldg(col + e - 1) >= n_users;
    if (bad) flags[1] = 1;
  }
}

static int check_rows(tgcn_graph* g, cudaStream_t stream) {
  int* flags = nullptr;
  int h[2] = {0, 0};
  TGCN_CHECK_CUDA(cudaMalloc(&flags, 2 * sizeof(int)));
  cudaError_t e = cudaMemsetAsync(flags, 0, 2 * sizeof(int), stream);
  if (e == cudaSuccess) {
    check_rows_kernel<<<(unsigned)((g->n_rows + 255) / 256), 256, 0, stream>>>(g->rowptr, g->col, (int)g->n_rows, (int)g->n_users, flags);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(h, flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(flags);
  TGCN_REQUIRE(e == cudaSuccess, "check_rows_kernel failed: %s", cudaGetErrorString(e));
  TGCN_REQUIRE(h[0] == 0, "CSR columns must be strictly increasing within every row (coalesced, sorted COO order)");
  g->bipartite = g->is_block ? 0 : (h[1] == 0);
  return 0;
}

static int build_segments(tgcn_graph* g, cudaStream_t stream) {
  std::vector<int> rowptr(g->n_rows + 1);
  TGCN_CHECK_CUDA(cudaMemcpyAsync(rowptr.data(), g->rowptr, sizeof(int) * (g->n_rows + 1), cudaMemcpyDeviceToHost, stream));
  TGCN_CHECK_CUDA(cudaStreamSynchronize(stream));
  TGCN_REQUIRE(rowptr[0] == 0 && (int64_t)rowptr[g->n_rows] == g->nnz, "rowptr does not span [0, nnz]: rowptr[0]=%d rowptr[n]=%d nnz=%lld",
               rowptr[0], rowptr[g->n_rows], (long long)g->nnz);
  std::vector<Segment> segs;
  std::vector<SplitRow> splits;
  std::vector<int> qptr(g->n_rows + 1);
  int64_t n_quads = 0;
  int max_deg = 0;
  for (int64_t r = 0; r < g->n_rows; ++r) {
    const int deg = rowptr[r + 1] - rowptr[r];
    TGCN_REQUIRE(deg >= 0, "rowptr is not non-decreasing at row %lld", (long long)r);
    if (deg > max_deg) max_deg = deg;
    qptr[r] = (int)n_quads;
    if (deg > kSplitThreshold) {
      SplitRow sr{(int)r, (int)segs.size(), 0, 0};
      for (int b = rowptr[r]; b < rowptr[r + 1]; b += kSegmentLen) {
        int e = b + kSegmentLen < rowptr[r + 1] ? b + kSegmentLen : rowptr[r + 1];
        const int qb = (int)n_quads + (b - rowptr[r]) / 4, qe = (int)n_quads + (e - rowptr[r] + 3) / 4;  // kSegmentLen % 4 == 0
        segs.push_back(Segment{(int)r, b, e, (int)segs.size(), (int)splits.size(), qb, qe, 0});
        sr.n_parts++;
      }
      splits.push_back(sr);
    }
    n_quads += (deg + 3) / 4;
    TGCN_REQUIRE(n_quads < (1ll << 31), "packed layout exceeds int32 indexing");
  }
  qptr[g->n_rows] = (int)n_quads;
  {
    // Packed copy for the SpMM kernels (A/B switch, read once per handle: TGCN_SPMM_PACKED=0 keeps the 32-bit col/val loads).
    // Failing to allocate it is not an error: the kernels then read the borrowed CSR arrays.
    const char* sw = getenv("TGCN_SPMM_PACKED");
    if (n_quads > 0 && !(sw && atoi(sw) == 0)) {
      bool ok = cudaMalloc(&g->qptr, sizeof(int) * qptr.size()) == cudaSuccess && cudaMalloc(&g->qcol, sizeof(int4) * n_quads) == cudaSuccess &&
                cudaMalloc(&g->qval, sizeof(float4) * n_quads) == cudaSuccess;
      if (ok) {
        ok = cudaMemcpyAsync(g->qptr, qptr.data(), sizeof(int) * qptr.size(), cudaMemcpyHostToDevice, stream) == cudaSuccess;
        if (ok) {
          pack_quads_kernel<<<(unsigned)((n_quads + 255) / 256), 256, 0, stream>>>(g->rowptr, g->col, g->val, g->qptr, (int)g->n_rows,
                                                                                   (int)n_quads, g->qcol, g->qval);
          ok = cudaGetLastError() == cudaSuccess && cudaStreamSynchronize(stream) == cudaSuccess;  // qptr (host vector) dies with this scope
        }
      }
      if (!ok) {
        (void)cudaGetLastError();
        if (g->qptr) cudaFree(g->qptr);
        if (g->qcol) cudaFree(g->qcol);
        if (g->qval) cudaFree(g->qval);
        g->qptr = nullptr;
        g->qcol = nullptr;
        g->qval = nullptr;
      }
    }
  }
  // processing order of the short rows: stable counting sort by (user/item phase, steps descending)
  {
    constexpr int kStep = 4;  // kUnroll of the SpMM kernels
    const int n_keys = kSplitThreshold / kStep + 1;
    const int64_t phase_split = g->is_block ? g->n_rows : g->n_users;
    std::vector<int> count(2 * n_keys + 1, 0);
    auto key_of = [&](int64_t r, int deg) { return (r < phase_split ? 0 : n_keys) + (n_keys - 1 - (deg + kStep - 1) / kStep); };
    for (int64_t r = 0; r < g->n_rows; ++r) {
      const int deg = rowptr[r + 1] - rowptr[r];
      if (deg <= kSplitThreshold) count[key_of(r, deg) + 1]++;
    }
    for (size_t k = 1; k < count.size(); ++k) count[k] += count[k - 1];
    std::vector<int> order(count.back());
    for (int64_t r = 0; r < g->n_rows; ++r) {
      const int deg = rowptr[r + 1] - rowptr[r];
      if (deg <= kSplitThreshold) order[count[key_of(r, deg)]++] = (int)r;
    }
    g->n_ordered = (int)order.size();
    const char* sw = getenv("TGCN_ROW_ORDER");  // A/B switch, read ONCE per handle: 0 = walk the rows in natural order
    if (!order.empty() && !(sw && atoi(sw) == 0)) {
      TGCN_CHECK_CUDA(cudaMalloc(&g->order, sizeof(int) * order.size()));
      TGCN_CHECK_CUDA(cudaMemcpyAsync(g->order, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice, stream));
      TGCN_CHECK_CUDA(cudaStreamSynchronize(stream));
    }
  }
  g->max_degree = max_deg;
  g->n_segments = (int)segs.size();
  g->n_user_segments = 0;  // segments are emitted in row order: those of rows < n_users come first
  for (const Segment& sg : segs)
    if (g->is_block || sg.row < g->n_users) g->n_user_segments++;
  g->n_split_rows = (int)splits.size();
  if (!segs.empty()) {
    TGCN_CHECK_CUDA(cudaMalloc(&g->segments, sizeof(Segment) * segs.size()));
    TGCN_CHECK_CUDA(cudaMalloc(&g->split_rows, sizeof(SplitRow) * splits.size()));
    TGCN_CHECK_CUDA(cudaMalloc(&g->split_counters, sizeof(int) * splits.size()));
    TGCN_CHECK_CUDA(cudaMemsetAsync(g->split_counters, 0, sizeof(int) * splits.size(), stream));
    TGCN_CHECK_CUDA(cudaMemcpyAsync(g->segments, segs.data(), sizeof(Segment) * segs.size(), cudaMemcpyHostToDevice, stream));
    TGCN_CHECK_CUDA(cudaMemcpyAsync(g->split_rows, splits.data(), sizeof(SplitRow) * splits.size(), cudaMemcpyHostToDevice, stream));
    TGCN_CHECK_CUDA(cudaStreamSynchronize(stream));
  }
  return 0;
}

}  // namespace tgcn

using namespace tgcn;

extern "C" {

int tgcn_abi_version(void) { return TGCN_ABI_VERSION; }
const char* tgcn_last_error(void) { return tgcn::g_err; }

static int create_common(tgcn_graph_t** out, int64_t n_users, int64_t n_items, int64_t row_begin, int64_t n_rows,
                         int64_t nnz, const int32_t* d_rowptr, const int32_t* d_col, const float* d_val, int is_block,
                         tgcn_stream_t stream) {
  TGCN_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  TGCN_REQUIRE(n_users > 0 && n_items > 0 && n_rows > 0 && nnz >= 0, "bad sizes: n_users=%lld n_items=%lld n_rows=%lld nnz=%lld",
               (long long)n_users, (long long)n_items, (long long)n_rows, (long long)nnz);
  TGCN_REQUIRE(n_users + n_items < (1ll << 31) && nnz < (1ll << 31), "graph exceeds int32 indexing");
  TGCN_REQUIRE(d_rowptr && d_col && d_val, "NULL CSR array");
  int dev_count = 0;
  TGCN_CHECK_CUDA(cudaGetDeviceCount(&dev_count));
  TGCN_REQUIRE(dev_count > 0, "no CUDA device: libtgcn_b200 has no CPU fallback");
  tgcn_graph* g = new tgcn_graph();
  g->n_users = n_users;
  g->n_items = n_items;
  g->row_begin = row_begin;
  g->n_rows = n_rows;
  g->nnz = nnz;
  g->is_block = is_block;
  g->rowptr = d_rowptr;
  g->col = d_col;
  g->val = d_val;
  g->tperm = nullptr;
  g->qptr = nullptr;
  g->qcol = nullptr;
  g->qval = nullptr;
  g->order = nullptr;
  g->n_ordered = 0;
  g->segments = nullptr;
  g->split_rows = nullptr;
  g->split_counters = nullptr;
  g->n_segments = g->n_split_rows = 0;
  g->bipartite = 0;
  g->mask_col_off = (int)n_users;
  g->busy_event = nullptr;
  g->busy_stream = nullptr;
  g->busy_valid = 0;
  g->n_user_segments = 0;
  g->host_in = g->host_out = nullptr;
  for (auto& e : g->host_ev) e = nullptr;
  int rc = build_segments(g, (cudaStream_t)stream);
  if (rc == 0 && nnz > 0) rc = check_rows(g, (cudaStream_t)stream);
  if (rc != 0) {
    tgcn_graph_destroy(g);
    return rc;
  }
  *out = g;
  return 0;
}

int tgcn_graph_create(tgcn_graph_t** out, int64_t n_users, int64_t n_items, int64_t nnz, const int32_t* d_rowptr,
                      const int32_t* d_col, const float* d_val, tgcn_stream_t stream) {
  return create_common(out, n_users, n_items, 0, n_users + n_items, nnz, d_rowptr, d_col, d_val, 0, stream);
}

int tgcn_graph_create_block(tgcn_graph_t** out, int64_t n_users, int64_t n_items, int64_t row_begin,
                            int64_t n_local_rows, int64_t nnz_local, const int32_t* d_rowptr, const int32_t* d_col,
                            const float* d_val, tgcn_stream_t stream) {
  return create_common(out, n_users, n_items, row_begin, n_local_rows, nnz_local, d_rowptr, d_col, d_val, 1, stream);
}

int tgcn_graph_build_transpose_perm(tgcn_graph_t* g, tgcn_stream_t stream) {
  TGCN_REQUIRE(g != nullptr, "graph is NULL");
  if (g->tperm) return 0;
  TGCN_REQUIRE(!g->is_block, "transpose permutation is only defined for a whole-graph handle");
  if (g->nnz == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  int* tperm = nullptr;
  int* missing = nullptr;
  TGCN_CHECK_CUDA(cudaMalloc(&tperm, sizeof(int) * g->nnz));
  TGCN_CHECK_CUDA(cudaMalloc(&missing, sizeof(int)));
  TGCN_CHECK_CUDA(cudaMemsetAsync(missing, 0, sizeof(int), s));
  const int threads = 256;
  const int64_t blocks = (g->nnz + threads - 1) / threads;
  transpose_perm_kernel<<<(unsigned)blocks, threads, 0, s>>>(g->rowptr, g->col, (int)g->n_rows, g->nnz, tperm, missing);
  int h_missing = 0;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h_missing, missing, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(missing);
  if (e != cudaSuccess || h_missing != 0) {
    cudaFree(tperm);
    if (e != cudaSuccess) set_error("transpose_perm_kernel failed: %s", cudaGetErrorString(e));
    else set_error("adjacency is not structurally symmetric: %d entries have no transpose partner", h_missing);
    return 3;
  }
  g->tperm = tperm;
  return 0;
}

void tgcn_graph_destroy(tgcn_graph_t* g) {
  if (!g) return;
  if (g->tperm) cudaFree(g->tperm);
  if (g->qptr) cudaFree(g->qptr);
  if (g->qcol) cudaFree(g->qcol);
  if (g->qval) cudaFree(g->qval);
  if (g->order) cudaFree(g->order);
  if (g->segments) cudaFree(g->segments);
  if (g->split_rows) cudaFree(g->split_rows);
  if (g->split_counters) cudaFree(g->split_counters);
  if (g->busy_event) cudaEventDestroy(g->busy_event);
  if (g->host_in) cudaStreamDestroy(g->host_in);
  if (g->host_out) cudaStreamDestroy(g->host_out);
  for (auto e : g->host_ev)
    if (e) cudaEventDestroy(e);
  delete g;
}

int64_t tgcn_graph_num_segments(const tgcn_graph_t* g) { return g ? g->n_segments : -1; }

int tgcn_graph_set_mask_col_offset(tgcn_graph_t* g, int64_t col_offset) {
  TGCN_REQUIRE(g != nullptr, "graph is NULL");
  TGCN_REQUIRE(col_offset >= 0 && col_offset < (1ll << 31), "bad column offset");
  g->mask_col_off = (int)col_offset;
  return 0;
}

}  // extern "C"
