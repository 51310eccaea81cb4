// Kernel-backed bodies for the reference's MATRIX-RETURNING scoring methods — the callers that want the dense
// result rather than the fused top-k:
//   score_batchwise            base_model.py:173-179      (B, d)·(n_items, d)ᵀ -> (B, n_items)
//   get_features_batchwise     ltr_models.py:131-146      5 such products written as the planes of (B, n_items, 5)
//   score_batchwise_ltr        ltr_models.py:200-204, :227-232   one product of width d + 2D with per-row / per-column bias
//   score_pairwise_adv         advanced_sampling.py:37-44  (B, d) x (B, C, d) -> (B, C)
//   get_features_pairwise      ltr_models.py:148-166      five row-wise dot products -> (B, 5)
// `predict` / `get_loss` never come through here (eval_tc.cu / eval_topk.cu / adv_ltr.cu fuse these products with the
// mask + top-k / the selection so that the intermediates never reach HBM); these exist so that every §8(b) name is a
// callable with the reference's signature and fp32 (FMA) arithmetic, as torch.matmul computes it with TF32 off (G11).
//
// Roofline: dense_nt_kernel is bound by the fp32 FMA pipe for K >= 64 (2·K flops per 4 output bytes) and by the
// HBM write of the (M, N) result below that; the two row-wise kernels are HBM gathers (B·C·4d and B·(2·4d + 4·4D) bytes).
#include "common.cuh"

namespace tgcn {

constexpr int DM = 128, DN = 128, DK = 16, DPAD = 4;
constexpr int kDenseThreads = 256;

struct DenseArgs {
  const float* a;  // (M, K) rows at lda
  int64_t lda;
  int M;
  const float* b;  // (N, K) rows at ldb
  int64_t ldb;
  int N;
  int K;
  const float* row_bias;  // (M) or NULL
  const float* col_bias;  // (N) or NULL
  float* out;             // element (m, n) at out[m·ldo_row + n·ldo_col]
  int64_t ldo_row, ldo_col;
  int vec_store;  // ldo_col == 1 and 16-byte aligned rows: 128-bit stores
};

// C = A·Bᵀ, 128x128x16 register-tiled SGEMM (8x8 outputs per thread, exact fp32 FMA), global -> register -> smem staging
// with the next K-chunk's loads in flight during the FMAs.
__global__ void __launch_bounds__(kDenseThreads, 2) dense_nt_kernel(const DenseArgs a) {
  __shared__ __align__(16) float As[DK][DM + DPAD];
  __shared__ __align__(16) float Bs[DK][DN + DPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * DN, m0 = blockIdx.y * DM;
  const int ld_row = tid >> 2, ld_k = (tid & 3) * 4;
  const float* a_ptr[2];
  const float* b_ptr[2];
  bool a_ok[2], b_ok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int m = m0 + ld_row + 64 * h, n = n0 + ld_row + 64 * h;
    a_ok[h] = m < a.M;
    b_ok[h] = n < a.N;
    a_ptr[h] = a.a + (size_t)(a_ok[h] ? m : 0) * a.lda;
    b_ptr[h] = a.b + (size_t)(b_ok[h] ? n : 0) * a.ldb;
  }
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float4 ra[2], rb[2];
  auto fetch = [&](int kc) {
    const int kk = kc * DK + ld_k;
    const bool kin = kk < a.K;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      ra[h] = (a_ok[h] && kin) ? ldg4(a_ptr[h] + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[h] = (b_ok[h] && kin) ? ldg4(b_ptr[h] + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  const int n_kchunks = (a.K + DK - 1) / DK;
  fetch(0);
  for (int kc = 0; kc < n_kchunks; ++kc) {
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = ld_row + 64 * h;
      As[ld_k + 0][r] = ra[h].x;
      As[ld_k + 1][r] = ra[h].y;
      As[ld_k + 2][r] = ra[h].z;
      As[ld_k + 3][r] = ra[h].w;
      Bs[ld_k + 0][r] = rb[h].x;
      Bs[ld_k + 1][r] = rb[h].y;
      Bs[ld_k + 2][r] = rb[h].z;
      Bs[ld_k + 3][r] = rb[h].w;
    }
    __syncthreads();
    if (kc + 1 < n_kchunks) fetch(kc + 1);
#pragma unroll
    for (int kk = 0; kk < DK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  float cb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = n0 + (j < 4 ? 0 : 64) + tx * 4 + (j & 3);
    cb[j] = (a.col_bias && n < a.N) ? __ldg(a.col_bias + n) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? 0 : 64) + ty * 4 + (i & 3);
    if (m >= a.M) continue;
    const float rbias = a.row_bias ? __ldg(a.row_bias + m) : 0.f;
    float* orow = a.out + (size_t)m * a.ldo_row;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + 64 * h + tx * 4;
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = acc[i][h * 4 + q] + (rbias + cb[h * 4 + q]);
      if (a.vec_store && n + 3 < a.N) {
        *reinterpret_cast<float4*>(orow + n) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (n + q < a.N) orow[(size_t)(n + q) * a.ldo_col] = v[q];
      }
    }
  }
}

// out[b, c] = <u[b, :], items[b, c, :]>: one warp per (b, c) pair, 128-bit loads.
__global__ void __launch_bounds__(256) pairwise_adv_kernel(int64_t n_pairs, int n_cand, int d, const float* __restrict__ u, int64_t ldu,
                                                           const float* __restrict__ items, float* __restrict__ out) {
  const int64_t pair = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pair >= n_pairs) return;
  const int64_t b = pair / n_cand;
  const float* up = u + b * ldu;
  const float* ip = items + pair * d;
  float part = 0.f;
  for (int c = lane; c < (d >> 2); c += 32) part += dot4(ldg4(up + c * 4), ldg4(ip + c * 4));
  part = warp_sum(part);
  if (lane == 0) out[pair] = part;
}

struct FeatRowsArgs {
  int batch, d, D;
  const float *ue, *ie, *ur, *ud, *ir, *id;
  int64_t ld_ue, ld_ie, ld_ur, ld_ud, ld_ir, ld_id;
  float* out;
  int64_t ldo;
};

// out[b, 0..4] = [ue·ie, ur·ir, ud·id, ur·id, ud·ir] of ROW-ALIGNED vectors (the reference's feature order, ltr_models.py:154-163)
__global__ void __launch_bounds__(256) ltr_features_rows_kernel(const FeatRowsArgs a) {
  const int row = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= a.batch) return;
  const float* ue = a.ue + (size_t)row * a.ld_ue;
  const float* ie = a.ie + (size_t)row * a.ld_ie;
  const float* ur = a.ur + (size_t)row * a.ld_ur;
  const float* ud = a.ud + (size_t)row * a.ld_ud;
  const float* ir = a.ir + (size_t)row * a.ld_ir;
  const float* id = a.id + (size_t)row * a.ld_id;
  float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f, f4 = 0.f;
  for (int c = lane; c < (a.d >> 2); c += 32) f0 += dot4(ldg4(ue + c * 4), ldg4(ie + c * 4));
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
    float* o = a.out + (size_t)row * a.ldo;
    o[0] = f0;
    o[1] = f1;
    o[2] = f2;
    o[3] = f3;
    o[4] = f4;
  }
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace tgcn

using namespace tgcn;

extern "C" {

int tgcn_score_batchwise(int64_t n_rows, int64_t n_cols, int64_t K, const float* d_a, int64_t lda, const float* d_b, int64_t ldb,
                         const float* d_row_bias, const float* d_col_bias, float* d_out, int64_t ldo_row, int64_t ldo_col,
                         tgcn_stream_t stream) {
  TGCN_REQUIRE(n_rows > 0 && n_cols > 0 && n_rows < (1ll << 31) && n_cols < (1ll << 31), "bad shape: %lld x %lld", (long long)n_rows, (long long)n_cols);
  TGCN_REQUIRE(K > 0 && K % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && lda >= K && ldb >= K, "K=%lld, lda=%lld, ldb=%lld must be multiples of 4 with ld >= K",
               (long long)K, (long long)lda, (long long)ldb);
  TGCN_REQUIRE(d_a && d_b && d_out && aligned16(d_a) && aligned16(d_b), "operands must be non-NULL and 16-byte aligned");
  TGCN_REQUIRE(ldo_row > 0 && ldo_col > 0, "bad output strides");
  const int64_t m_tiles = (n_rows + DM - 1) / DM, n_tiles = (n_cols + DN - 1) / DN;
  TGCN_REQUIRE(m_tiles <= 65535, "too many row tiles (%lld): rank at most 8.3M rows per call", (long long)m_tiles);
  DenseArgs a;
  a.a = d_a;
  a.lda = lda;
  a.M = (int)n_rows;
  a.b = d_b;
  a.ldb = ldb;
  a.N = (int)n_cols;
  a.K = (int)K;
  a.row_bias = d_row_bias;
  a.col_bias = d_col_bias;
  a.out = d_out;
  a.ldo_row = ldo_row;
  a.ldo_col = ldo_col;
  a.vec_store = ldo_col == 1 && ldo_row % 4 == 0 && aligned16(d_out);
  dense_nt_kernel<<<dim3((unsigned)n_tiles, (unsigned)m_tiles), kDenseThreads, 0, (cudaStream_t)stream>>>(a);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int tgcn_score_pairwise_adv(int64_t batch, int32_t n_cand, int64_t d, const float* d_users_emb, int64_t ldu, const float* d_items_emb,
                            float* d_out, tgcn_stream_t stream) {
  TGCN_REQUIRE(batch > 0 && n_cand > 0 && d > 0 && d % 4 == 0 && ldu % 4 == 0 && ldu >= d, "bad sizes: batch=%lld n_cand=%d d=%lld ldu=%lld",
               (long long)batch, n_cand, (long long)d, (long long)ldu);
  TGCN_REQUIRE(d_users_emb && d_items_emb && d_out && aligned16(d_users_emb) && aligned16(d_items_emb), "operands must be non-NULL and 16-byte aligned");
  const int64_t n_pairs = batch * n_cand;
  const int64_t blocks = (n_pairs * 32 + 255) / 256;
  TGCN_REQUIRE(blocks < (1ll << 31), "grid too large");
  pairwise_adv_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n_pairs, n_cand, (int)d, d_users_emb, ldu, d_items_emb, d_out);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int tgcn_ltr_features_rows(int64_t batch, int64_t d, int64_t D, const float* d_ue, int64_t ld_ue, const float* d_ie, int64_t ld_ie,
                           const float* d_ur, int64_t ld_ur, const float* d_ud, int64_t ld_ud, const float* d_ir, int64_t ld_ir,
                           const float* d_id, int64_t ld_id, float* d_out, int64_t ldo, tgcn_stream_t stream) {
  TGCN_REQUIRE(batch > 0 && batch < (1ll << 26) && d > 0 && d % 4 == 0 && D > 0 && D % 4 == 0, "bad sizes: batch=%lld d=%lld D=%lld", (long long)batch,
               (long long)d, (long long)D);
  TGCN_REQUIRE(d_ue && d_ie && d_ur && d_ud && d_ir && d_id && d_out && ldo >= 5, "NULL argument or ldo < 5");
  TGCN_REQUIRE(ld_ue % 4 == 0 && ld_ie % 4 == 0 && ld_ur % 4 == 0 && ld_ud % 4 == 0 && ld_ir % 4 == 0 && ld_id % 4 == 0, "row strides must be multiples of 4");
  TGCN_REQUIRE(aligned16(d_ue) && aligned16(d_ie) && aligned16(d_ur) && aligned16(d_ud) && aligned16(d_ir) && aligned16(d_id), "operands must be 16-byte aligned");
  FeatRowsArgs a{(int)batch, (int)d, (int)D, d_ue, d_ie, d_ur, d_ud, d_ir, d_id, ld_ue, ld_ie, ld_ur, ld_ud, ld_ir, ld_id, d_out, ldo};
  ltr_features_rows_kernel<<<(unsigned)((batch * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  TGCN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
