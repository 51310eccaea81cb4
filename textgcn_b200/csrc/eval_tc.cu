// Full-ranking evaluation on the 5th-generation tensor cores: the user x item score GEMM runs on tcgen05.mma (fp32 accumulation
// in TMEM) fused with train-item masking and a streaming per-user top-k, so the score matrix exists only in tensor memory
// (base_model.py:255-261).  Two arithmetic schemes share one kernel template:
//
//   * 3xTF32 (short item sweeps, k up to 64, bias terms, K > 128, and the second pass of the screened scheme): operands are pre-split
//     once per call into [hi | lo] TF32 halves (hi = rna_tf32(x), lo = rna_tf32(x - hi); both have zero low mantissa bits, so the
//     tensor core's input truncation is exact) and hi·hi + lo·hi + hi·lo is accumulated.  Dropping lo·lo bounds the relative product
//     error by ~2^-21; scores agree with the fp32 SGEMM to ~1e-6 norm-wise (the parity budget is 1e-5) and are bit-exact whenever the
//     inputs are TF32-representable (the tie fixture).
//   * SCREEN (long item sweeps, k <= 24; resident user tile for K <= 128 without bias terms, streamed otherwise): ONE TF32 product per
//     score from the raw user rows and a rounded copy of the item rows finds 40 candidates per user; those that can matter are re-scored in exact fp32 FMA and a per-row certificate decides
//     whether the row is provably the exact top-k or must be ranked again by the 3xTF32 scheme (device-gated second pass).  See the
//     SCREEN / INS template comments and screen_finalize below; DESIGN.md §4 has the derivation and the measurements.
//
// Common structure:
//   * One CTA = 128 users x the whole item range (or one split of it); CTA pairs (cta_group::2) share every 256-row item tile between
//     two SMs on long sweeps.  The user tile is loaded ONCE by TMA and stays resident in shared memory (K <= 128; wider contractions
//     stream it with the item tile); item tiles of BN rows stream through an mbarrier ring of 128-byte-swizzled K-chunks (32 fp32 = one
//     swizzle row).  Warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warp 2 = TMEM allocator, warps 4..4+4·EW-1 =
//     epilogue (EW = 2: eight warps), and for SCREEN four more "inserter" warps that own the lists.  Item tiles are 256 rows whenever
//     three ring stages still fit: a 128x128x8 TF32 MMA measured ~117 cycles against 64 ideal.
//   * Accumulators: 2 x BN TMEM columns (double buffered): the epilogue of tile t overlaps the MMAs of tile t+1.
//   * Epilogue: TMEM lane = user row.  The EW warps of a lane quarter split every tile's COLUMNS.  A thread reads 32 scores with
//     tcgen05.ld (the load of the next 32 in flight meanwhile), takes their MAXIMUM (one FMNMX3 per two scores, four independent chains)
//     and compares it with the row's k-th best; only when that fires — ~k·(1 + ln(n/k)) times per sweep — does it build the 32-bit hit
//     mask and handle the candidates.  3xTF32: each of the row's EW threads keeps a private sorted list in REGISTERS over its share of
//     the items, consults a 128-bit register Bloom filter of the user's train items (exact binary search in the user row of Â only on
//     a Bloom hit) and shift-inserts; the lists are merged (lexicographic insert) once, after the sweep, through the then-idle item
//     ring.  Items arrive in increasing id order within a thread, so a strict '>' keeps the canonical (score desc, id asc) order.
//     SCREEN: the candidates go through a per-row ring in shared memory to the row's inserter thread (INS).
//     History (ncu, c2): row-split warps that each rebuilt a full hit mask for 32 rows but owned 16 ran the tensor pipe at 45 % —
//     0.37 warp instructions per score, 2 warps per scheduler with a tcgen05.wait::ld stall per group.
//
// Roofline: tensor pipe.  3xTF32: 3 x 2·K TF32 flops per score (MMA floor 128 cycles per 128x256x8 instruction); measured (ncu) tensor
// pipe 93.6-95 % active at 2 M items / d = 128 (~860 TFLOP/s TF32 under the power cap), 49 % at 63 k items / d = 64 (bound by the list
// updates of the sweep's opening).  SCREEN: 2·K flops per score; tensor pipe 78-85 % active at 2 M items / d = 128, 2.5 x the 3xTF32
// scheme's users/s.
#include <cuda.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace tgcn {

constexpr int TC_BM = 128;
constexpr int TC_CHUNK = 32;            // fp32 per 128-byte swizzle row
constexpr int TC_A_CHUNK_BYTES = TC_BM * 128;
constexpr int TC_MAX_STAGES = 8;
constexpr int kInsRing = 16;           // INS variant: candidates a row can have in flight between its scanner threads and its inserter
constexpr int kInsBytes = 2 * kInsRing * 128 * 4 + 3 * 128 * 4 + 64;  // rings (score, id) + tail / head / threshold per row + done counters
constexpr int kBarBlockBytes = 256;   // mbarriers + the TMEM address slot ((10 + 2·TC_MAX_STAGES)·8 + 16 bytes, rounded up)

struct TcArgs {
  int n_rank, K, n_range, item_begin, k;
  int tiles_per_split, n_stages;
  const int* users;
  const int* mrowptr;
  const int* mcol;
  int mrow_begin, mcol_off;
  int* part_ids;
  float* part_scores;
  int direct;    // single split: write the final table here (no merge pass), completing short lists when `finalize`
  int finalize;
  int* out_ids;
  float* out_scores;
  // device-gated re-run (the screened path's exact second pass): rank only the first *n_rank_dev rows, write row m to out_rows[m]
  const int* n_rank_dev;
  const int* out_rows;
  int chunk_major;    // streamed screened variant: operands stored K-chunk plane by plane ([KC][rows][32]); TMA row = kc·rows + r
  int split_fastest;  // rasterisation: consecutive CTAs (pairs) take the SPLITS of one user tile (pair) instead of consecutive user tiles
  // screened variant: raw fp32 operands (1xTF32 scores), exact fp32 re-scoring of the candidates that matter, certificate
  int Kr;                         // real contraction width (K is padded to whole 32-wide chunks; TMA zero-fills the pad)
  const float* ivec;              // item table (row = item id) and its leading dimension, for the exact re-scoring
  int64_t ldi;
  float eps_c;                    // |s_tf32 - s| <= eps_c * |u| * |i|
  const unsigned* max_inorm2;     // bits of max_i |i|^2 over the ranked item range (written by screen_prep_items_kernel)
  const float* ug;                // streamed screened variant: the raw user operand (n_rank, K) in the workspace — the exact rows for re-scoring
  const float* ibias;             // item bias table (by item id) or NULL; the user bias sits in the operand's bias chunk
  int has_bias;
  float eps_ub, eps_ib;           // |error| <= ... + eps_ub·|ub| + eps_ib·max|ib| (bias chunk: [ub, 1] x [1, ib])
  const unsigned* max_ib;         // bits of max |ib| over the ranked item range
  int* fb_mark;                   // per rank row: 1 once the row is queued for the exact second pass
  int* fb_rows;                   // queue of rank rows whose certificate failed
  int* fb_count;
};

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// The suspend-time hint lets the hardware park the waiting thread instead of having it poll: ncu counted 1.06 G of the
// kernel's 4.5 G issued instructions in the single-thread TMA / MMA wait loops, on the schedulers the epilogue warps need.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(1000000u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU (traps after ~2 s instead).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- CTA pair (cta_group::2): two SMs of one TPC execute ONE MMA of M = 256; each CTA supplies its own 128 rows of A and HALF of the
// B tile from its own shared memory and keeps its 128 accumulator rows in its own tensor memory.  Only the leader (cluster rank 0)
// issues MMAs; both CTAs issue TMA loads whose transaction bytes land on the LEADER's barrier (peer bit of the address cleared).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {  // arrives on `bar` (same offset) in BOTH CTAs once the MMAs retire
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar, uint32_t cta) {  // arrive on `bar` of CTA `cta` of the cluster
  asm volatile(
      "{\n"
      ".reg .b32 rem;\n"
      "mapa.shared::cluster.u32 rem, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [rem];\n"
      "}\n" ::"r"(bar),
      "r"(cta)
      : "memory");
}
// K-major, 128-byte swizzle: 8-row atoms of 1024 B, SBO = 1024 B, LBO unused, descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// v[j] for a run-time j without spilling v to local memory: a 5-level multiplexer of selects (rare path only)
__device__ __forceinline__ float pick32(const uint32_t (&v)[32], int j) {
  uint32_t a[16], b[8], c[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 16) ? v[i + 16] : v[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 8) ? a[i + 8] : a[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[i + 4] : b[i];
  const uint32_t d0 = (j & 2) ? c[2] : c[0], d1 = (j & 2) ? c[3] : c[1];
  return __uint_as_float((j & 1) ? d1 : d0);
}

// ---- per-thread top-k list held in REGISTERS (KL entries, sorted best-first) ------------------------------------
// Insert (s, id): entry j takes its upper neighbour when s beats that neighbour, s itself when s beats only entry j.
// Every entry is decided from the ORIGINAL values (walk bottom-up), so the KL updates are independent (high ILP,
// ~5 instructions each, no memory).  Candidates arrive in increasing id order, so a strict '>' keeps the earlier
// entry ahead on equal scores (canonical order).
template <int KL>
__device__ __forceinline__ void reg_list_insert(float (&ls)[KL], int (&li)[KL], float s, int id) {
#pragma unroll
  for (int j = KL - 1; j >= 1; --j) {
    const bool up = s > ls[j - 1];
    const bool here = s > ls[j];
    ls[j] = up ? ls[j - 1] : (here ? s : ls[j]);
    li[j] = up ? li[j - 1] : (here ? id : li[j]);
  }
  const bool top = s > ls[0];
  ls[0] = top ? s : ls[0];
  li[0] = top ? id : li[0];
}

// Same for a candidate that may carry a LOWER id than an equal-score entry (merging the lists of the column slices).
template <int KL>
__device__ __forceinline__ void reg_list_insert_lex(float (&ls)[KL], int (&li)[KL], float s, int id) {
#pragma unroll
  for (int j = KL - 1; j >= 1; --j) {
    const bool up = ranks_before(s, id, ls[j - 1], li[j - 1]);
    const bool here = ranks_before(s, id, ls[j], li[j]);
    ls[j] = up ? ls[j - 1] : (here ? s : ls[j]);
    li[j] = up ? li[j - 1] : (here ? id : li[j]);
  }
  const bool top = ranks_before(s, id, ls[0], li[0]);
  ls[0] = top ? s : ls[0];
  li[0] = top ? id : li[0];
}

// ---- screened variant: what happens to a row's list after the sweep (called by all EW threads of the row) -------------------
struct ScreenFin {
  float* ms;            // [kl] approximate scores of the row's merged list, sorted best-first (in the idle item ring) ...
  int* mi;              // [kl] ... their item ids (INT_MAX = never filled) ...
  float* me;            // [kl] ... and room for their exact scores
  int* mp;              // one int per row
  const uint8_t* urow;  // the row in the resident user tile: chunk c at + c·16 KB, 16-byte unit q at ((q ^ (t & 7)) << 4)
  const float* ug;      // streamed variant (no resident tile): the row's first chunk in the raw user operand (chunk-major: + ug_plane per chunk), else NULL
  int64_t ug_plane;     // floats between two K-chunk planes of that operand
  const float* ibias;   // item bias table (by item id) or NULL
  float ub;             // user bias (0 without)
  float eps_bias;       // what the bias terms add to the error bound
  int t, sub, n_sub, bar_id, kl, m, mlo, mhi, split;
  bool valid;
};

__device__ __forceinline__ float4 screen_user_unit(const ScreenFin& f, int g) {  // 16-byte unit g of the row's exact user vector
  if (f.ug) return __ldg(reinterpret_cast<const float4*>(f.ug + (int64_t)(g >> 3) * f.ug_plane) + (g & 7));
  return *reinterpret_cast<const float4*>(f.urow + (size_t)(g >> 3) * TC_A_CHUNK_BYTES + (((g & 7) ^ (f.t & 7)) << 4));
}

// exact fp32 score of list entries [lo, hi) — strided over the row's n_sub threads — from the fp32 tables
__device__ __forceinline__ void screen_rescore(const ScreenFin& f, int lo, int hi, int k4, const float* __restrict__ ivec, int64_t ldi) {
  for (int j = lo + f.sub; j < hi; j += f.n_sub) {
    const int id = f.mi[j];
    const float4* ip = reinterpret_cast<const float4*>(ivec + (size_t)id * ldi);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int g0 = 0; g0 < k4; g0 += 8) {
      float4 iv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) iv[q] = g0 + q < k4 ? __ldg(ip + g0 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 x = g0 + q < k4 ? screen_user_unit(f, g0 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[0] = fmaf(x.x, iv[q].x, acc[0]);
        acc[1] = fmaf(x.y, iv[q].y, acc[1]);
        acc[2] = fmaf(x.z, iv[q].z, acc[2]);
        acc[3] = fmaf(x.w, iv[q].w, acc[3]);
      }
    }
    f.me[j] = (((acc[0] + acc[1]) + (acc[2] + acc[3])) + f.ub) + (f.ibias ? __ldg(f.ibias + id) : 0.f);
  }
}

// With eps = eps_c·|u|·max|i| >= |approximate - exact|:  (1) the k best entries by approximate score are re-scored; E = the smallest
// of their exact scores, so k items score at least E;  (2) an entry whose approximate score + eps is below E cannot be among the k
// best, nor can any entry after it (the list is sorted) or any item that is not in the full list (it scores at most the last entry):
// entries are re-scored up to the first such one;  (3) the re-scored entries, sorted on (exact score desc, id asc), start with the
// exact top-k.  If no entry of a FULL list can be ruled out, an item outside the list might belong to the k best: the row is queued
// for the second pass.
__device__ __noinline__ void screen_finalize(const ScreenFin f, int Kr, int k, const float* __restrict__ ivec, int64_t ldi, float eps_c,
                                             const unsigned* __restrict__ max_inorm2, int* fb_mark, int* fb_rows, int* fb_count, int direct,
                                             int finalize, const int* __restrict__ mcol, int mcol_off, int n_rank_stride,
                                             const int* __restrict__ out_rows, int* out_ids, float* out_scores, int* part_ids,
                                             float* part_scores) {
  const int k4 = Kr >> 2;
  int n_valid = 0;
  float eps = 0.f;
  if (f.sub == 0) {
    float un2 = 0.f;
    for (int g = 0; g < k4; ++g) {
      const float4 x = screen_user_unit(f, g);
      un2 = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, un2))));
    }
    eps = eps_c * sqrtf(un2) * sqrtf(__uint_as_float(__ldg(max_inorm2))) + f.eps_bias;
    while (n_valid < f.kl && f.mi[n_valid] != INT_MAX) ++n_valid;
    *f.mp = n_valid < k ? n_valid : k;
  }
  asm volatile("bar.sync %0, %1;" ::"r"(f.bar_id), "r"(32 * f.n_sub) : "memory");
  const int P1 = *f.mp;
  screen_rescore(f, 0, P1, k4, ivec, ldi);
  asm volatile("bar.sync %0, %1;" ::"r"(f.bar_id), "r"(32 * f.n_sub) : "memory");
  bool flagged = false;
  if (f.sub == 0) {
    int P = P1;
    if (n_valid >= k) {
      float E = INFINITY;
      for (int j = 0; j < k; ++j) E = fminf(E, f.me[j]);
      while (P < n_valid && f.ms[P] + eps >= E) ++P;
      flagged = f.valid && ((P == f.kl) || !(eps < INFINITY));
    }
    *f.mp = P;
  }
  asm volatile("bar.sync %0, %1;" ::"r"(f.bar_id), "r"(32 * f.n_sub) : "memory");
  const int P = *f.mp;
  screen_rescore(f, P1, P, k4, ivec, ldi);
  asm volatile("bar.sync %0, %1;" ::"r"(f.bar_id), "r"(32 * f.n_sub) : "memory");
  if (f.sub != 0 || !f.valid) return;
  for (int j = 1; j < P; ++j) {  // insertion sort on (exact score desc, id asc)
    const float s = f.me[j];
    const int id = f.mi[j];
    int q = j;
    while (q > 0 && ranks_before(s, id, f.me[q - 1], f.mi[q - 1])) {
      f.me[q] = f.me[q - 1];
      f.mi[q] = f.mi[q - 1];
      --q;
    }
    f.me[q] = s;
    f.mi[q] = id;
  }
  if (flagged && atomicExch(fb_mark + f.m, 1) == 0) fb_rows[atomicAdd(fb_count, 1)] = f.m;
  const int real = P < k ? P : k;  // P < k only when fewer than k items could be ranked at all
  if (direct) {
    const size_t o = (size_t)(out_rows ? __ldg(out_rows + f.m) : f.m) * k;
    for (int j = 0; j < k; ++j) {
      int id = j < real ? f.mi[j] : -1;
      float s = j < real ? f.me[j] : -INFINITY;
      if (j >= real) {
        if (finalize) {  // fewer than k rankable items: complete with train items, lowest id first (G9)
          const int tt = j - real;
          id = tt < f.mhi - f.mlo ? __ldg(mcol + f.mlo + tt) - mcol_off : -1;
        } else {
          id = INT_MAX;
        }
      }
      out_ids[o + j] = id;
      out_scores[o + j] = s;
    }
  } else {
    const size_t o = ((size_t)f.split * n_rank_stride + f.m) * k;
    for (int j = 0; j < k; ++j) {
      part_ids[o + j] = j < real ? f.mi[j] : INT_MAX;
      part_scores[o + j] = j < real ? f.me[j] : -INFINITY;
    }
  }
}

// BN = item rows per tile, KL = list capacity (>= k), EW = epilogue warps per TMEM lane quarter: warp `sub` of a quarter
// scans columns [sub·BN/EW, (sub+1)·BN/EW) of every tile for the quarter's 32 user rows.
// STREAM = false: the user tile [hi | lo] stays resident in shared memory for the whole sweep (K <= 128).
// STREAM = true : wide contractions (the LTR score, K = d + 2D + bias chunk): user and item K-chunks travel together
//                 through the ring, one stage = {U_hi, U_lo, I_hi, I_lo} of one 32-wide K-chunk (12 MMAs per stage).
// CTAS = 2 (non-streamed only): CTA pairs (cluster of 2 along x) — every MMA covers 256 users x BN items, the item tile is split
//                 between the two CTAs' rings (BN / 2 rows each), so the shared-memory operand fetch per SM and the number of MMA
//                 instructions per score both halve.  Each CTA keeps its own users' accumulator rows, lists and epilogue.
// SCREEN (non-streamed only): the operands are the RAW fp32 tables (no hi/lo split) and every score is ONE TF32 product — a third of
//                 the tensor work.  The tensor core reads the upper 19 bits of each fp32, so |s_tf32 - s| <= eps = eps_c·|u|·|i|
//                 (2·2^-10 from the two truncations + accumulation, Cauchy-Schwarz over the products).  The sweep keeps the best
//                 KL > k items by the APPROXIMATE score; afterwards every kept item within 2·eps of the approximate k-th best is
//                 re-scored in exact fp32 FMA from the tables and the list is re-sorted on the exact scores.  An item outside
//                 that band has an exact score below at least k re-scored ones, so the result is the exact top-k PROVIDED the band
//                 ends inside the list; a row whose band reaches the list's last entry (an excluded item might belong to it) is
//                 queued in fb_rows and ranked again by the 3xTF32 variant (eval_topk_tc, second pass gated on the device).
// NACC = accumulator stages in tensor memory (NACC·BN <= 512 columns): 2 lets the epilogue of one tile overlap the MMAs of the next;
//                 4 (BN = 128) lets a warp that updates lists in one tile catch up over the next three before the MMA issuer, which
//                 waits for the slowest epilogue warp of the stage it is about to overwrite, has to wait for it — measured on the
//                 screened pair variant and NOT adopted: c5 163.9 ms against 129.3 ms with two stages of 256 columns (the 128-column
//                 MMA costs nearly as much per instruction as the 256-column one; profiles/r02/README.md).
// INS (screened variant): four more warps own the lists.  The epilogue ("scanner") threads only take the maximum of their scores and push
//                 the few that beat the row's threshold into a small per-row ring in shared memory; inserter thread t keeps row t's ONE
//                 sorted list in registers, pops the ring, applies the train-item mask and publishes the new threshold.  The list updates
//                 — hundreds of instructions for one lane while its warp waits — leave the path that drains tensor memory, whose pace the
//                 MMA issuer depends on (a stage is released by the slowest scanner warp), and the two column slices of a row share one
//                 list and one threshold instead of keeping two.
template <int BN, int KL, int EW, bool STREAM, int CTAS = 1, bool SCREEN = false, int NACC = 2, bool INS = false>
__global__ void __launch_bounds__(128 + 128 * EW + (INS ? 128 : 0), 1)
eval_topk_tc_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_i, const TcArgs a) {
  static_assert(CTAS == 1 || CTAS == 2, "one CTA or a CTA pair per 128 / 256 users");
  static_assert(!SCREEN || INS, "the screened variant runs with inserter warps");
  static_assert(NACC * BN <= 512 && (NACC == 2 || NACC == 4), "tensor memory holds 512 accumulator columns");
  static_assert(!INS || SCREEN, "inserter warps are implemented for the screened variant");
  constexpr int PLANES = SCREEN ? 1 : 2;  // operand planes per K-chunk: the raw fp32 values, or their [hi | lo] TF32 halves
  constexpr bool PAIR = CTAS == 2;
  int n_rank = a.n_rank;
  if (a.n_rank_dev) n_rank = min(n_rank, __ldg(a.n_rank_dev));
  // Which user tile and which item split this CTA takes.  Default: blockIdx.x = user tile, blockIdx.y = split.  split_fastest: CTAs are
  // scheduled in linear order (x fastest), and the streamed variant wants the CTAs that run at the same time to share user tiles —
  // every CTA re-streams its 128 x 2K user tile once per item tile, and 148 different tiles of 1.7 MB (K = 1600) thrash the L2
  // (ncu: 61.9 GB of DRAM reads per call, nearly all of it user operand) — so consecutive CTA groups take the splits of ONE tile group.
  unsigned tile_x = blockIdx.x, split_y = blockIdx.y;
  if (a.split_fastest) {
    const unsigned lin = blockIdx.y * gridDim.x + blockIdx.x, grp = lin / CTAS;
    split_y = grp % gridDim.y;
    tile_x = (grp / gridDim.y) * CTAS + lin % CTAS;
  }
  if ((int)(tile_x / CTAS * CTAS) * TC_BM >= n_rank) return;  // (a whole pair leaves together)
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle-128B tiles need 1024-byte alignment
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const int KC = a.K / TC_CHUNK;
  const int n_a = STREAM ? 0 : (SCREEN ? KC : 2 * KC);  // resident user chunks
  constexpr int BROWS = BN / CTAS;  // item rows of a tile held by THIS CTA's ring
  constexpr int B_STAGE_BYTES = STREAM ? PLANES * (TC_A_CHUNK_BYTES + BROWS * 128) : BROWS * 128;
  const uint32_t sA = base;
  const uint32_t sB = sA + n_a * TC_A_CHUNK_BYTES;
  uint8_t* ring = gen_base + n_a * TC_A_CHUNK_BYTES;  // item ring; reused for the list merge once the sweep is over
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + a.n_stages * B_STAGE_BYTES);
  const uint32_t bar_a_full = smem_u32(bars + 0);
  const uint32_t bar_b_full = smem_u32(bars + 1);                      // [TC_MAX_STAGES]
  const uint32_t bar_b_empty = smem_u32(bars + 1 + TC_MAX_STAGES);     // [TC_MAX_STAGES]
  const uint32_t bar_t_full = smem_u32(bars + 1 + 2 * TC_MAX_STAGES);  // [4]
  const uint32_t bar_t_empty = smem_u32(bars + 5 + 2 * TC_MAX_STAGES); // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9 + 2 * TC_MAX_STAGES);

  float* rq_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBlockBytes);  // [kInsRing][128] (INS only)
  int* rq_i = reinterpret_cast<int*>(rq_s + kInsRing * 128);                                   // [kInsRing][128], -1 = empty slot
  int* rq_tail = rq_i + kInsRing * 128;                                                        // [128] pushes claimed
  int* rq_head = rq_tail + 128;                                                                // [128] pops done
  float* rq_thr = reinterpret_cast<float*>(rq_head + 128);                                     // [128] the row's current KL-th best
  int* rq_done = reinterpret_cast<int*>(rq_thr + 128);                                         // [4] scanner warps of the quarter that finished

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = tile_x * TC_BM;
  const int n_tiles = (a.n_range + BN - 1) / BN;
  const int tile_begin = split_y * a.tiles_per_split;
  const int tile_end = min(n_tiles, tile_begin + a.tiles_per_split);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_u) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_i) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar_a_full, 1);
    for (int s = 0; s < TC_MAX_STAGES; ++s) {
      mbar_init(bar_b_full + 8 * s, 1);
      mbar_init(bar_b_empty + 8 * s, 1);
    }
    for (int s = 0; s < NACC; ++s) {
      mbar_init(bar_t_full + 8 * s, 1);
      mbar_init(bar_t_empty + 8 * s, 4 * EW * CTAS);  // one arrival per epilogue warp (of both CTAs of a pair, on the leader's barrier)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (INS) {
    if (warp >= 4 + 4 * EW) {
      const int t = threadIdx.x - (128 + 128 * EW);
      for (int e = 0; e < kInsRing; ++e) rq_i[e * 128 + t] = -1;
      rq_tail[t] = 0;
      rq_head[t] = 0;
      rq_thr[t] = -INFINITY;
      if (t < 4) rq_done[t] = 0;
    }
  }
  if constexpr (PAIR) cluster_sync_all();  // both CTAs' barriers exist before anyone (TMA of the peer, remote arrives) touches them
  if (warp == 2) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(NACC * BN) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(NACC * BN) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer: resident user tile, then the item chunks of every tile ----
      int stage = 0;
      uint32_t phase = 0;
      if constexpr (STREAM) {
        auto load = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
          if constexpr (PAIR) tma_load_2d_pair(dst, map, bar, c0, c1);
          else tma_load_2d(dst, map, bar, c0, c1);
        };
        for (int tile = tile_begin; tile < tile_end; ++tile) {
          const int row0 = tile * BN + (int)cta_rank * BROWS;  // pair: this CTA streams its half of the item tile (and its own users)
          for (int kc = 0; kc < KC; ++kc) {
            const uint32_t st = sB + stage * B_STAGE_BYTES, full = bar_b_full + 8 * stage;
            mbar_wait(bar_b_empty + 8 * stage, phase ^ 1);
            if (leader) mbar_expect_tx(full, CTAS * B_STAGE_BYTES);
            if constexpr (SCREEN) {  // one raw chunk of each operand (chunk-major layout: plane kc, contiguous boxes)
              load(st, &map_u, full, 0, kc * a.n_rank + m0);
              load(st + TC_A_CHUNK_BYTES, &map_i, full, 0, kc * a.n_range + row0);
            } else {
              load(st, &map_u, full, kc * TC_CHUNK, m0);
              load(st + TC_A_CHUNK_BYTES, &map_u, full, a.K + kc * TC_CHUNK, m0);
              load(st + 2 * TC_A_CHUNK_BYTES, &map_i, full, kc * TC_CHUNK, row0);
              load(st + 2 * TC_A_CHUNK_BYTES + BROWS * 128, &map_i, full, a.K + kc * TC_CHUNK, row0);
            }
            if (++stage == a.n_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      } else {
      if (leader) mbar_expect_tx(bar_a_full, CTAS * n_a * TC_A_CHUNK_BYTES);  // pair: the leader's barrier counts both CTAs' user tiles
      for (int c = 0; c < n_a; ++c) {
        if constexpr (PAIR) tma_load_2d_pair(sA + c * TC_A_CHUNK_BYTES, &map_u, bar_a_full, c * TC_CHUNK, m0);
        else tma_load_2d(sA + c * TC_A_CHUNK_BYTES, &map_u, bar_a_full, c * TC_CHUNK, m0);
      }
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int row0 = tile * BN + (int)cta_rank * BROWS;  // pair: this CTA streams its half of the item tile
        for (int c = 0; c < n_a; ++c) {  // order: hi_0, lo_0, hi_1, lo_1, ...  (screened: the raw chunks in order)
          const int colk = SCREEN ? c * TC_CHUNK : ((c & 1) ? a.K + (c >> 1) * TC_CHUNK : (c >> 1) * TC_CHUNK);
          mbar_wait(bar_b_empty + 8 * stage, phase ^ 1);   // own barrier: the MMA commit is multicast to both CTAs
          if (leader) mbar_expect_tx(bar_b_full + 8 * stage, CTAS * B_STAGE_BYTES);
          if constexpr (PAIR) tma_load_2d_pair(sB + stage * B_STAGE_BYTES, &map_i, bar_b_full + 8 * stage, colk, row0);
          else tma_load_2d(sB + stage * B_STAGE_BYTES, &map_i, bar_b_full + 8 * stage, colk, row0);
          if (++stage == a.n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ---- MMA issuer: D[128·CTAS x BN] (+)= A[128·CTAS x 8] · B[BN x 8]ᵀ, kind::tf32, fp32 accumulate in TMEM ----
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((TC_BM * CTAS) >> 4) << 24);
      auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t acc) {
        if constexpr (PAIR) umma_tf32_pair(d, ad, bd, idesc, acc);
        else umma_tf32(d, ad, bd, idesc, acc);
      };
      auto commit = [&](uint32_t bar) {
        if constexpr (PAIR) umma_commit_pair(bar);
        else umma_commit(bar);
      };
      if constexpr (!STREAM) {
        mbar_wait(bar_a_full, 0);
        tc_fence_after();
      }
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        mbar_wait(bar_t_empty + 8 * as, aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        uint32_t accumulate = 0;
        if constexpr (STREAM) {
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait(bar_b_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t st = sB + stage * B_STAGE_BYTES;
            if constexpr (SCREEN) {
              const uint32_t a_raw = st, b_raw = st + TC_A_CHUNK_BYTES;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                mma(d_tmem, umma_desc(a_raw + kk * 32), umma_desc(b_raw + kk * 32), accumulate);
                accumulate = 1;
              }
            } else {
              const uint32_t a_hi = st, a_lo = st + TC_A_CHUNK_BYTES, b_hi = st + 2 * TC_A_CHUNK_BYTES, b_lo = b_hi + BROWS * 128;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                mma(d_tmem, umma_desc(a_hi + kk * 32), umma_desc(b_hi + kk * 32), accumulate);
                accumulate = 1;
              }
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) mma(d_tmem, umma_desc(a_lo + kk * 32), umma_desc(b_hi + kk * 32), 1);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) mma(d_tmem, umma_desc(a_hi + kk * 32), umma_desc(b_lo + kk * 32), 1);
            }
            commit(bar_b_empty + 8 * stage);
            if (++stage == a.n_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        for (int c = 0; c < n_a; ++c) {
          mbar_wait(bar_b_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t bb = sB + stage * B_STAGE_BYTES;
          const int kc = SCREEN ? c : c >> 1;
          const uint32_t a_hi = sA + kc * TC_A_CHUNK_BYTES, a_lo = sA + (KC + kc) * TC_A_CHUNK_BYTES;
          if constexpr (SCREEN) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              mma(d_tmem, umma_desc(a_hi + kk * 32), umma_desc(bb + kk * 32), accumulate);
              accumulate = 1;
            }
          } else if ((c & 1) == 0) {  // B = hi chunk: hi·hi and lo·hi
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              mma(d_tmem, umma_desc(a_hi + kk * 32), umma_desc(bb + kk * 32), accumulate);
              accumulate = 1;
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma(d_tmem, umma_desc(a_lo + kk * 32), umma_desc(bb + kk * 32), 1);
          } else {  // B = lo chunk: hi·lo
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma(d_tmem, umma_desc(a_hi + kk * 32), umma_desc(bb + kk * 32), 1);
          }
          commit(bar_b_empty + 8 * stage);  // frees the smem slot (of both CTAs of a pair) when these MMAs retire
          if (++stage == a.n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        commit(bar_t_full + 8 * as);  // accumulator of this tile complete (signalled in both CTAs of a pair)
        if (++as == NACC) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    if constexpr (INS) {
      constexpr int CW = BN / EW, NG = CW / 32;
      static_assert(NG >= 2 && NG % 2 == 0, "the TMEM loads are double buffered two groups at a time");
      const int q = warp & 3;  // TMEM lane quarter = row quarter
      const int t = q * 32 + lane;
      const int m = m0 + t;
      const bool valid = m < n_rank;
      float* ms = reinterpret_cast<float*>(ring) + (size_t)t * (3 * KL);  // the row's list after the sweep (item ring, idle by then)
      int* mi = reinterpret_cast<int*>(ms + KL);
      ScreenFin f;
      f.ms = ms;
      f.mi = mi;
      f.me = ms + 2 * KL;
      f.mp = reinterpret_cast<int*>(ring + (size_t)TC_BM * 3 * KL * 4) + t;
      f.urow = gen_base + (size_t)t * 128;
      f.ug = nullptr;
      f.ug_plane = 0;
      f.ibias = a.ibias;
      f.ub = 0.f;
      f.eps_bias = 0.f;
      if constexpr (STREAM) {  // no resident tile: the exact user row comes from the raw operand in global memory (bias chunk last)
        f.ug = a.ug + (size_t)(valid ? m : 0) * TC_CHUNK;
        f.ug_plane = (int64_t)a.n_rank * TC_CHUNK;
        if (a.has_bias) {
          f.ub = __ldg(f.ug + (int64_t)(a.K / TC_CHUNK - 1) * f.ug_plane);
          f.eps_bias = a.eps_ub * fabsf(f.ub) + a.eps_ib * __uint_as_float(__ldg(a.max_ib));
        }
      }
      f.t = t;
      f.n_sub = EW + 1;
      f.bar_id = 1 + q;
      f.kl = KL;
      f.m = m;
      f.valid = valid;
      f.mlo = 0;
      f.mhi = 0;
      f.split = split_y;
      if (warp < 4 + 4 * EW) {
        // ---- scanner: maximum of 32 scores against the row's published threshold; the rare hits go to the row's ring ----
        const int sub = (warp - 4) >> 2;
        volatile float* vthr = rq_thr + t;
        volatile int* vhead = rq_head + t;
        auto scan32 = [&](const uint32_t (&v)[32], int nbase) {
          float mx[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
#pragma unroll
          for (int j = 4; j < 32; ++j) mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(v[j]));
          const float thr = *vthr;
          if (!(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) > thr) || !valid) return;
          uint32_t h4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (__uint_as_float(v[j]) > thr) h4[j & 3] |= 1u << j;
          uint32_t hits = (h4[0] | h4[1]) | (h4[2] | h4[3]);
          const int rem = a.n_range - nbase;  // columns past the item range hold zero-filled rows
          if (rem < 32) hits &= rem > 0 ? (1u << rem) - 1u : 0u;
          while (hits) {
            const int j = __ffs(hits) - 1;
            hits &= hits - 1;
            const float s = pick32(v, j);
            if (!(s > *vthr)) continue;  // the inserter may have raised the threshold meanwhile
            const int slot = atomicAdd(rq_tail + t, 1);
            if (slot - *vhead >= kInsRing) {  // ring full: wait for the inserter (bounded)
              const long long t0 = clock64();
              while (slot - *vhead >= kInsRing)
                if (clock64() - t0 > 4000000000ll) __trap();
            }
            rq_s[(slot & (kInsRing - 1)) * 128 + t] = s;
            __threadfence_block();
            *reinterpret_cast<volatile int*>(rq_i + (slot & (kInsRing - 1)) * 128 + t) = a.item_begin + nbase + j;
          }
        };
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = tile_begin; tile < tile_end; ++tile) {
          mbar_wait(bar_t_full + 8 * as, aphase);
          tc_fence_after();
          const int n0 = tile * BN + sub * CW;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + sub * CW;
          uint32_t va[32], vb[32];
          tmem_ld32(taddr, va);
#pragma unroll 1
          for (int g = 0; g < NG; g += 2) {
            tmem_wait_ld();
            tmem_ld32(taddr + (g + 1) * 32, vb);
            scan32(va, n0 + g * 32);
            tmem_wait_ld();
            if (g + 2 < NG) tmem_ld32(taddr + (g + 2) * 32, va);
            scan32(vb, n0 + (g + 1) * 32);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cta(bar_t_empty + 8 * as, 0);
            else mbar_arrive(bar_t_empty + 8 * as);
          }
          if (++as == NACC) {
            as = 0;
            aphase ^= 1;
          }
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) atomicAdd(rq_done + q, 1);  // every push of this warp is in shared memory
        f.sub = 1 + sub;
      } else {
        // ---- inserter: row t's list (registers), Bloom filter and mask row; pops the row's ring until the quarter's scanners are done ----
        const int user = valid ? (a.users ? __ldg(a.users + m) : m) : 0;
        TGCN_DASSERT(!valid || !a.mrowptr || user >= a.mrow_begin);
        int mlo = 0, mhi = 0;
        if (valid && a.mrowptr) {
          mlo = __ldg(a.mrowptr + user - a.mrow_begin);
          mhi = __ldg(a.mrowptr + user - a.mrow_begin + 1);
        }
        uint32_t bloom[4] = {0u, 0u, 0u, 0u};
        for (int p = mlo; p < mhi; ++p) {
          const uint32_t b = (uint32_t)(__ldg(a.mcol + p) - a.mcol_off) & 127u;
          bloom[0] |= (b >> 5) == 0 ? 1u << (b & 31) : 0u;
          bloom[1] |= (b >> 5) == 1 ? 1u << (b & 31) : 0u;
          bloom[2] |= (b >> 5) == 2 ? 1u << (b & 31) : 0u;
          bloom[3] |= (b >> 5) == 3 ? 1u << (b & 31) : 0u;
        }
        float ls[KL];
        int li[KL];
#pragma unroll
        for (int j = 0; j < KL; ++j) {
          ls[j] = -INFINITY;
          li[j] = INT_MAX;
        }
        float thr = -INFINITY;
        int head = 0;
        volatile int* vtail = rq_tail + t;
        volatile int* vdone = rq_done + q;
        long long t0 = clock64();  // of the last progress (the spins below are bounded: a protocol bug traps instead of hanging)
        for (;;) {
          const bool fin = *vdone == EW;  // read BEFORE the tail: a push precedes its warp's done count
          const bool has = *vtail != head;
          if (has) {
            t0 = clock64();
            volatile int* slot_i = rq_i + (head & (kInsRing - 1)) * 128 + t;
            int item;
            while ((item = *slot_i) < 0)  // claimed but not yet written
              if (clock64() - t0 > 20000000000ll) __trap();
            __threadfence_block();
            const float s = *reinterpret_cast<volatile float*>(rq_s + (head & (kInsRing - 1)) * 128 + t);
            *slot_i = -1;
            __threadfence_block();
            ++head;
            *reinterpret_cast<volatile int*>(rq_head + t) = head;
            if (s > thr) {
              const uint32_t b = (uint32_t)item & 127u;
              const uint32_t word = (b >> 5) == 0 ? bloom[0] : (b >> 5) == 1 ? bloom[1] : (b >> 5) == 2 ? bloom[2] : bloom[3];
              if (!((word >> (b & 31)) & 1u) || !sorted_contains(a.mcol, mlo, mhi, item + a.mcol_off)) {
                reg_list_insert<KL>(ls, li, s, item);  // (arrival is not in id order: the exact re-sort restores the canonical order)
                thr = ls[KL - 1];
                *reinterpret_cast<volatile float*>(rq_thr + t) = thr;
              }
            }
          }
          if (!__any_sync(0xffffffffu, has)) {
            if (fin) break;
            __nanosleep(64);
            if (clock64() - t0 > 20000000000ll) __trap();
          }
        }
#pragma unroll
        for (int j = 0; j < KL; ++j) {
          ms[j] = ls[j];
          mi[j] = li[j];
        }
        f.sub = 0;
        f.mlo = mlo;
        f.mhi = mhi;
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * (EW + 1)) : "memory");  // the list is staged; scanners join the re-scoring
      screen_finalize(f, a.Kr, a.k, a.ivec, a.ldi, a.eps_c, a.max_inorm2, a.fb_mark, a.fb_rows, a.fb_count, a.direct, a.finalize, a.mcol, a.mcol_off,
                      a.n_rank, a.out_rows, a.out_ids, a.out_scores, a.part_ids, a.part_scores);
    } else {
    // ---- epilogue: EW threads per user row, each scanning BN / EW columns of every tile ----
    constexpr int CW = BN / EW;       // columns per warp and tile
    constexpr int NG = CW / 32;       // 32-column groups per warp and tile
    static_assert(NG >= 2 && NG % 2 == 0, "the TMEM loads are double buffered two groups at a time");
    const int ew = warp & 3;          // the TMEM lane quarter this warp may read (warp id % 4)
    const int sub = (warp - 4) >> 2;  // which column slice of every tile this warp scans
    const int t = ew * 32 + lane;
    const int m = m0 + t;
    const bool valid = m < n_rank;
    const int user = valid ? (a.users ? __ldg(a.users + m) : m) : 0;
    TGCN_DASSERT(!valid || !a.mrowptr || user >= a.mrow_begin);
    int mlo = 0, mhi = 0;
    if (valid && a.mrowptr) {
      mlo = __ldg(a.mrowptr + user - a.mrow_begin);
      mhi = __ldg(a.mrowptr + user - a.mrow_begin + 1);
    }
    // 128-bit Bloom filter of the user's train items (bit = item id mod 128): the exact membership test (a binary
    // search in global memory, ~1 us of dependent latency) only runs for candidates whose bit is set.
    uint32_t bloom[4] = {0u, 0u, 0u, 0u};
    for (int p = mlo; p < mhi; p += 4) {
      int it[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) it[q] = p + q < mhi ? __ldg(a.mcol + p + q) - a.mcol_off : -1;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (it[q] >= 0) {
          const uint32_t b = (uint32_t)it[q] & 127u;
          bloom[0] |= (b >> 5) == 0 ? 1u << (b & 31) : 0u;
          bloom[1] |= (b >> 5) == 1 ? 1u << (b & 31) : 0u;
          bloom[2] |= (b >> 5) == 2 ? 1u << (b & 31) : 0u;
          bloom[3] |= (b >> 5) == 3 ? 1u << (b & 31) : 0u;
        }
    }
    float ls[KL];
    int li[KL];
#pragma unroll
    for (int j = 0; j < KL; ++j) {
      ls[j] = -INFINITY;
      li[j] = INT_MAX;
    }
    float thr = -INFINITY;
    // 32 scores of this row: nothing to do unless their maximum beats the row's current k-th best
    auto scan32 = [&](const uint32_t (&v)[32], int nbase) {
      float mx[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
#pragma unroll
      for (int j = 4; j < 32; ++j) mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(v[j]));
      if (!(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) > thr) || !valid) return;
      // rare and divergent: ~k·(1 + ln(n/k)) candidates per list over the whole sweep
      uint32_t h4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (__uint_as_float(v[j]) > thr) h4[j & 3] |= 1u << j;
      uint32_t hits = (h4[0] | h4[1]) | (h4[2] | h4[3]);
      const int rem = a.n_range - nbase;  // columns past the item range hold zero-filled rows
      if (rem < 32) hits &= rem > 0 ? (1u << rem) - 1u : 0u;
      while (hits) {
        const int j = __ffs(hits) - 1;
        hits &= hits - 1;
        const float s = pick32(v, j);
        if (s > thr) {
          const int item = a.item_begin + nbase + j;
          const uint32_t b = (uint32_t)item & 127u;
          const uint32_t word = (b >> 5) == 0 ? bloom[0] : (b >> 5) == 1 ? bloom[1] : (b >> 5) == 2 ? bloom[2] : bloom[3];
          const bool maybe = (word >> (b & 31)) & 1u;
          TGCN_DASSERT(item >= a.item_begin && item < a.item_begin + a.n_range);
          if (!maybe || !sorted_contains(a.mcol, mlo, mhi, item + a.mcol_off)) {
            reg_list_insert<KL>(ls, li, s, item);
            thr = ls[KL - 1];
          }
        }
      }
    };
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      mbar_wait(bar_t_full + 8 * as, aphase);
      tc_fence_after();
      const int n0 = tile * BN + sub * CW;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + as * BN + sub * CW;
      uint32_t va[32], vb[32];
      tmem_ld32(taddr, va);
#pragma unroll 1
      for (int g = 0; g < NG; g += 2) {  // group g + 1 is in flight while group g is scanned
        tmem_wait_ld();
        tmem_ld32(taddr + (g + 1) * 32, vb);
        scan32(va, n0 + g * 32);
        tmem_wait_ld();
        if (g + 2 < NG) tmem_ld32(taddr + (g + 2) * 32, va);
        scan32(vb, n0 + (g + 1) * 32);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cta(bar_t_empty + 8 * as, 0);  // the leader's MMA issuer waits for both CTAs' epilogues
        else mbar_arrive(bar_t_empty + 8 * as);
      }
      if (++as == NACC) {
        as = 0;
        aphase ^= 1;
      }
    }
    if constexpr (EW > 1) {
      // merge the column slices' lists into warp `sub == 0`'s: the sweep is over (the last accumulator was complete, so
      // every TMA load has landed and every MMA has read its operands) and the item ring is free
      float* ms = reinterpret_cast<float*>(ring) + (size_t)t * (2 * KL);
      int* mi = reinterpret_cast<int*>(ms + KL);
      for (int r = 1; r < EW; ++r) {
        if (sub == r) {
#pragma unroll
          for (int j = 0; j < KL; ++j) {
            ms[j] = ls[j];
            mi[j] = li[j];
          }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + ew), "r"(32 * EW) : "memory");
        if (sub == 0) {
#pragma unroll 1
          for (int j = 0; j < KL; ++j) {
            const int id = mi[j];
            if (id != INT_MAX) reg_list_insert_lex<KL>(ls, li, ms[j], id);
          }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + ew), "r"(32 * EW) : "memory");
      }
    }
    {
    const bool writer = valid && sub == 0;
    if (writer && a.direct) {
      int real = 0;  // the list is sorted, so sentinels (never-filled slots) come last
#pragma unroll
      for (int j = 0; j < KL; ++j) real += li[j] != INT_MAX ? 1 : 0;
      const size_t o = (size_t)(a.out_rows ? __ldg(a.out_rows + m) : m) * a.k;
#pragma unroll
      for (int j = 0; j < KL; ++j)
        if (j < a.k) {
          int id = li[j];
          float s = ls[j];
          if (a.finalize && j >= real) {  // fewer than k rankable items: complete with train items, lowest id first (G9)
            const int tt = j - real;
            id = tt < mhi - mlo ? __ldg(a.mcol + mlo + tt) - a.mcol_off : -1;
            s = -INFINITY;
          }
          a.out_ids[o + j] = id;
          a.out_scores[o + j] = s;
        }
    } else if (writer) {
      const size_t o = ((size_t)split_y * a.n_rank + m) * a.k;
#pragma unroll
      for (int j = 0; j < KL; ++j)
        if (j < a.k) {
          a.part_ids[o + j] = li[j];
          a.part_scores[o + j] = ls[j];
        }
    }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // the leader's MMAs read the peer's shared memory: nobody leaves before everything retired
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NACC * BN) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NACC * BN) : "memory");
  }
}

// out row r = [hi(x), hi(bias chunk) | lo(x), lo(bias chunk)] of source row (rows ? rows[r] : row_begin + r); each half is
// Kp = K (+ 32 when bias terms exist) floats.  The optional bias chunk folds `score += user_bias[u] + item_bias[i]` into
// the contraction: the user side carries [ub, 1, 0, ...] and the item side [1, ib, 0, ...], so no epilogue work is needed.
__global__ void __launch_bounds__(256) tf32_split_kernel(const float* __restrict__ src, int64_t ld, const int* __restrict__ rows,
                                                         int64_t row_begin, int64_t n_rows, int K, int Kp,
                                                         const float* __restrict__ bias, const int* __restrict__ bias_rows,
                                                         int64_t bias_begin, int item_side, int bias_chunk, float* __restrict__ out,
                                                         const int* __restrict__ gate, int gate_rows) {
  const int k4 = Kp >> 2;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_rows * k4) return;
  const int64_t r = t / k4;
  if (gate) {  // device-gated second pass: only the first *gate rows (gate_rows) / nothing at all when *gate == 0
    const int c = __ldg(gate);
    if (gate_rows ? r >= c : c == 0) return;
  }
  const int c = (int)(t % k4) * 4;
  float xs[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < K) {
    const int64_t sr = rows ? (int64_t)__ldg(rows + r) : row_begin + r;
    const float4 x = ldg4(src + sr * ld + c);
    xs[0] = x.x;
    xs[1] = x.y;
    xs[2] = x.z;
    xs[3] = x.w;
  } else if (c == Kp - TC_CHUNK && bias_chunk) {
    const float b = bias ? __ldg(bias + (bias_rows ? (int64_t)__ldg(bias_rows + r) : bias_begin + r)) : 0.f;
    xs[0] = item_side ? 1.f : b;
    xs[1] = item_side ? b : 1.f;
  }
  float hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(xs[i]));
    hi[i] = __uint_as_float(h);
    const float rem = xs[i] - hi[i];
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(rem));
    lo[i] = __uint_as_float(l);
  }
  *reinterpret_cast<float4*>(out + r * 2 * Kp + c) = make_float4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<float4*>(out + r * 2 * Kp + Kp + c) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

// Screened variant, item side: out (n_rows, Kp) = rna_tf32([src row | zero pad | bias chunk [1, ib, 0 ...] when bias terms exist]) — the
// tensor core then reads exactly these values, so the item operand is off by at most 2^-11 relative instead of the 2^-10 of a 19-bit
// truncation — plus the max |row|^2 of the exact rows and the max |ib|.
__global__ void __launch_bounds__(256) screen_prep_items_kernel(const float* __restrict__ src, int64_t ld, int64_t row_begin, int64_t n_rows,
                                                                int K, int Kp, const float* __restrict__ ibias, int bias_chunk,
                                                                int chunk_major, float* __restrict__ out, unsigned* __restrict__ max_norm2,
                                                                unsigned* __restrict__ max_ib) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float best = 0.f, best_b = 0.f;
  for (int64_t r = warp; r < n_rows; r += n_warps) {
    const float* row = src + (row_begin + r) * ld;
    float acc = 0.f;
    for (int c = lane * 4; c < Kp; c += 128) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < K) {
        x = ldg4(row + c);
        acc = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, acc))));
      } else if (bias_chunk && c == Kp - TC_CHUNK) {
        const float b = ibias ? __ldg(ibias + row_begin + r) : 0.f;
        x.x = 1.f;
        x.y = b;
        best_b = fmaxf(best_b, fabsf(b));
      }
      uint32_t h[4];
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h[0]) : "f"(x.x));
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h[1]) : "f"(x.y));
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h[2]) : "f"(x.z));
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h[3]) : "f"(x.w));
      // chunk_major: plane kc holds the kc-th 32-wide K-chunk of every row, so that a TMA box of 128 rows is one contiguous 16 KB block
      float* dst = chunk_major ? out + ((int64_t)(c >> 5) * n_rows + r) * TC_CHUNK + (c & 31) : out + r * Kp + c;
      *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    best = fmaxf(best, acc);
  }
  if (lane == 0 && best > 0.f) atomicMax(max_norm2, __float_as_uint(best * 1.0001f));  // (any summation order stays below the bound)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best_b = fmaxf(best_b, __shfl_xor_sync(0xffffffffu, best_b, o));
  if (lane == 0 && best_b > 0.f) atomicMax(max_ib, __float_as_uint(best_b));
}

// Streamed screened variant, user side: out (n_rows, Kp) = [src row (rows ? rows[r] : r) | zero pad | bias chunk [ub, 1, 0 ...]], raw fp32
// (the exact rows the re-scoring reads; the tensor core truncates them itself).
__global__ void __launch_bounds__(256) screen_prep_users_kernel(const float* __restrict__ src, int64_t ld, const int* __restrict__ rows,
                                                                int64_t n_rows, int K, int Kp, const float* __restrict__ ubias,
                                                                int bias_chunk, int chunk_major, float* __restrict__ out) {
  const int k4 = Kp >> 2;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_rows * k4) return;
  const int64_t r = t / k4;
  const int c = (int)(t % k4) * 4;
  const int64_t sr = rows ? (int64_t)__ldg(rows + r) : r;
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < K) {
    x = ldg4(src + sr * ld + c);
  } else if (bias_chunk && c == Kp - TC_CHUNK) {
    x.x = ubias ? __ldg(ubias + sr) : 0.f;
    x.y = 1.f;
  }
  float* dst = chunk_major ? out + ((int64_t)(c >> 5) * n_rows + r) * TC_CHUNK + (c & 31) : out + r * Kp + c;
  *reinterpret_cast<float4*>(dst) = x;
}

// out (n_rows, K) = src[rows[r], :K]
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, int64_t ld, const int* __restrict__ rows, int64_t n_rows,
                                                          int K, float* __restrict__ out) {
  const int k4 = K >> 2;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_rows * k4) return;
  const int64_t r = t / k4;
  const int c = (int)(t % k4) * 4;
  *reinterpret_cast<float4*>(out + r * K + c) = ldg4(src + (int64_t)__ldg(rows + r) * ld + c);
}

// the queue of rank rows the screen could not certify -> the user ids (mask rows) and source rows of the second pass
__global__ void __launch_bounds__(256) fb_index_kernel(const int* __restrict__ fb_rows, const int* __restrict__ fb_count,
                                                       const int* __restrict__ users, int by_pos, int n_rank, int* __restrict__ fb_users,
                                                       int* __restrict__ fb_src) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= min(n_rank, __ldg(fb_count))) return;
  const int pos = __ldg(fb_rows + j);
  const int user = users ? __ldg(users + pos) : pos;
  fb_users[j] = user;
  fb_src[j] = (by_pos || !users) ? pos : user;
}

// Screened path, after the second pass: the rows it ranked carry 3xTF32 scores; give them the same exact fp32 scores (same FMA order as
// screen_rescore) and order as every other row, so that a row's output does not depend on which pass produced it.  One warp per queue
// entry, lane e = list position e (k <= 24).
__global__ void __launch_bounds__(256) fb_rescore_kernel(const int* __restrict__ fb_rows, const int* __restrict__ fb_count, int n_rank,
                                                         const int* __restrict__ users, int by_pos, const float* __restrict__ uvec, int64_t ldu,
                                                         const float* __restrict__ ivec, int64_t ldi, int K, int k,
                                                         const float* __restrict__ ubias, const float* __restrict__ ibias,
                                                         int* __restrict__ out_ids, float* __restrict__ out_scores) {
  const int lane = threadIdx.x & 31;
  const int j = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (j >= min(n_rank, __ldg(fb_count))) return;
  const int m = __ldg(fb_rows + j);
  const int64_t src = (by_pos || !users) ? m : __ldg(users + m);
  const size_t o = (size_t)m * k;
  int id = INT_MAX;
  float s = -INFINITY;
  bool real = false;
  if (lane < k) {
    id = out_ids[o + lane];
    s = out_scores[o + lane];
    real = s != -INFINITY;  // short lists end with -inf entries (train items / sentinels): they keep their places
    if (real) {
      const float4* up = reinterpret_cast<const float4*>(uvec + src * ldu);
      const float4* ip = reinterpret_cast<const float4*>(ivec + (size_t)id * ldi);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int g = 0; g < (K >> 2); ++g) {
        const float4 x = __ldg(up + g), y = __ldg(ip + g);
        acc[0] = fmaf(x.x, y.x, acc[0]);
        acc[1] = fmaf(x.y, y.y, acc[1]);
        acc[2] = fmaf(x.z, y.z, acc[2]);
        acc[3] = fmaf(x.w, y.w, acc[3]);
      }
      s = (((acc[0] + acc[1]) + (acc[2] + acc[3])) + (ubias ? __ldg(ubias + src) : 0.f)) + (ibias ? __ldg(ibias + id) : 0.f);
    }
  }
  int rank = 0;  // position among the real entries under (score desc, id asc)
  for (int e = 0; e < k; ++e) {
    const float se = __shfl_sync(0xffffffffu, s, e);
    const int ie = __shfl_sync(0xffffffffu, id, e);
    const bool re = __shfl_sync(0xffffffffu, real ? 1 : 0, e) != 0;
    if (re && e != lane && ranks_before(se, ie, s, id)) ++rank;
  }
  if (lane < k && real) {
    out_ids[o + rank] = id;
    out_scores[o + rank] = s;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int box_rows, int64_t ld = 0) {
  EncodeTiledFn fn = encode_fn();
  TGCN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(ld > 0 ? ld : cols) * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_CHUNK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TGCN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)r);
  return 0;
}

static inline int64_t al256(int64_t x) { return (x + 255) / 256 * 256; }

// Shared-memory plan for the tensor-core kernel (Kp = contraction width incl. the bias chunk); 0 stages = does not fit.
// CTA pairs (cta_group::2) for the resident-user-tile variant: on for long item sweeps, where the halved operand fetch per SM
// pays (c5, 2 M items: 561 k -> 569 k users/s on the same box; the kernel runs against the power cap, ~1.64 GHz, either way) and
// off for short ones, where the two cluster barriers and pair scheduling cost more (c2, 63 k items: 22.03 M -> 21.54 M users/s).
// TGCN_EVAL_PAIR=0 / 1 (read once per process) forces it off / on.  profiles/r02/README.md.
constexpr int64_t kPairMinItems = 262144;
static int pair_mode() {
  static const int mode = [] {
    const char* e = getenv("TGCN_EVAL_PAIR");
    return e ? (atoi(e) != 0 ? 1 : 0) : -1;
  }();
  return mode;
}
static bool pair_wanted(int64_t n_range, int k) {
  if (k > 40) return false;  // the pair variant is instantiated for the two register-list sizes with EW = 2
  const int m = pair_mode();
  return m < 0 ? n_range >= kPairMinItems : m == 1;
}

static void tc_plan(int Kp, int* bn, int* n_stages, size_t* smem, bool* stream, bool want_pair = false, bool* pair = nullptr) {
  *stream = Kp > 128;
  const bool use_pair = want_pair;
  if (pair) *pair = use_pair;
  const size_t a_bytes = *stream ? 0 : (size_t)(2 * (Kp / TC_CHUNK)) * TC_A_CHUNK_BYTES;
  const size_t fixed = 1024 /*align slack*/ + a_bytes + kBarBlockBytes;
  const size_t budget = 227 * 1024;
  // 256-row item tiles whenever three ring stages still fit beside the resident user tile: a 128x128x8 TF32 MMA was
  // measured at ~117 cycles against 64 ideal (issue overhead per instruction), so wide tiles matter more than ring depth
  // (c5, d = 128: 357 k -> 544 k users/s, 836 TFLOP/s).  The streamed variant has two 96 KB stages at 256 rows and still gains
  // (LTR, K = 1600: 875 k -> 1003 k users/s).
  *bn = 128;
  if (*stream || (fixed < budget && (budget - fixed) / ((size_t)256 * 128) >= 3)) *bn = 256;
  static const bool force_bn128 = [] {  // A/B switch, read once per process
    const char* e = getenv("TGCN_EVAL_BN");
    return e && atoi(e) == 128;
  }();
  if (force_bn128) *bn = 128;
  if (use_pair) *bn = 256;  // each CTA's ring holds half of the 256-row tile
  const size_t stage = *stream ? (size_t)2 * TC_A_CHUNK_BYTES + 2 * (size_t)*bn * 128 : (size_t)*bn * 128 / (use_pair ? 2 : 1);
  int s = fixed < budget ? (int)((budget - fixed) / stage) : 0;
  if (s > TC_MAX_STAGES) s = TC_MAX_STAGES;
  *n_stages = s;
  *smem = fixed + (size_t)s * stage;
}

// contraction width as the kernel sees it: K rounded up to whole 32-wide chunks (zero padded) + one chunk for the bias terms
static inline int tc_padded_k(int64_t K, bool has_bias) { return (int)((K + TC_CHUNK - 1) / TC_CHUNK) * TC_CHUNK + (has_bias ? TC_CHUNK : 0); }

bool eval_tc_eligible(int64_t K, int32_t k, bool has_bias) {
  if (K % 4 != 0 || K <= 0 || K > 8192 || k > 64) return false;  // register-resident lists: k <= 64
  int bn, st;
  size_t sm;
  bool stream;
  tc_plan(tc_padded_k(K, has_bias), &bn, &st, &sm, &stream);
  return st >= 2;
}

int eval_tc_tile_n(int64_t K, bool has_bias) {
  int bn, st;
  size_t sm;
  bool stream;
  tc_plan(tc_padded_k(K, has_bias), &bn, &st, &sm, &stream);
  return bn;
}

// ---- workspace layout (one place) --------------------------------------------------------------------------------
// [u2 | i2 | ctrl (count, max |i|^2) + fb_mark | fb_rows | fb_users | fb_src | part_ids | part_scores]
// The partial tables come last and are sized for the larger of the caller's split count and the second pass's (tc_fb_split_plan),
// so the two passes of a screened call agree on where everything before them lives.
constexpr int kFbSplits = 8;
void eval_split_plan(int64_t n_rank, int64_t n_items_range, int bn, int* n_splits, int* tiles_per_split);

// Item splits of the device-gated second pass: it usually ranks a handful of rows, so a lone CTA would sweep the whole range.
static void tc_fb_split_plan(int64_t n_rank, int64_t n_range, int bn, int* n_splits, int* tps, int64_t Kp = 128, int32_t k = 20) {
  eval_split_plan(n_rank, n_range, bn, n_splits, tps);
  const int64_t n_tiles = (n_range + bn - 1) / bn;
  // about 128 K-chunk steps of 256 items per split (32 tiles at K = 128, 2-3 at the LTR width: one row queued there used to cost 2 ms),
  // at most 32 splits, and partial tables of at most 256 MB (8 splits are always allowed)
  const int64_t kc = (Kp + TC_CHUNK - 1) / TC_CHUNK;
  int64_t want = n_tiles * kc / 128;
  int64_t cap = (256ll << 20) / (n_rank * k * 8 > 0 ? n_rank * k * 8 : 1);
  if (cap < kFbSplits) cap = kFbSplits;
  if (cap > 32) cap = 32;
  if (want > cap) want = cap;
  if (want > *n_splits) {
    *tps = (int)((n_tiles + want - 1) / want);
    *n_splits = (int)((n_tiles + *tps - 1) / *tps);
  }
}

// Item splits of the STREAMED variant: every CTA re-streams its 128 x 2Kp user tile once per item tile, so the CTAs that run at the
// same time should share user tiles (split_fastest rasterisation) and be few enough tiles that those stay in L2: enough splits that
// 148 concurrent CTAs hold at most ~32 MB of user tiles, at least 16 item tiles per split.  TGCN_EVAL_STREAM_RASTER=0 (read once): off.
static bool stream_raster_enabled() {
  static const bool on = [] {
    const char* e = getenv("TGCN_EVAL_STREAM_RASTER");
    return !(e && atoi(e) == 0);
  }();
  return on;
}
static void tc_stream_split_plan(int64_t n_rank, int64_t n_range, int64_t Kp, int bn, int* n_splits, int* tps, int planes = 2) {
  eval_split_plan(n_rank, n_range, bn, n_splits, tps);
  if (!stream_raster_enabled()) return;
  const int64_t n_tiles = (n_range + bn - 1) / bn;
  const int64_t tile_bytes = (int64_t)TC_BM * planes * Kp * 4;
  static const int64_t target_env = [] {  // experiment switch (read once): TGCN_EVAL_STREAM_L2_MB
    const char* e = getenv("TGCN_EVAL_STREAM_L2_MB");
    const int v = e ? atoi(e) : 0;
    return (int64_t)(v > 0 ? v : 0);
  }();
  // measured at the LTR shape: 3xTF32 14.6 ms at 32 MB (8 splits), 15.0 at 64, 17.3 at 16; the screened form, whose every split pays
  // the list updates of a sweep's opening again, 12.4 ms at 2 splits against 17.3 at 4 and 26.3 at 8
  const int64_t target_mb = target_env > 0 ? target_env : (planes == 1 ? 64 : 32);
  int64_t want = (148 * tile_bytes + (target_mb << 20) - 1) / (target_mb << 20);
  if (want > n_tiles / 16) want = n_tiles / 16;
  if (want > 16) want = 16;
  if (want > *n_splits) {
    *tps = (int)((n_tiles + want - 1) / want);
    *n_splits = (int)((n_tiles + *tps - 1) / *tps);
  }
}

struct TcWorkspace {
  float *u2, *i2;
  int* part_ids;
  float* part_scores;
  int* ctrl;  // [0] = queue length, [1] = bits of max |i|^2; fb_mark follows at +64 ints (one memset clears both)
  int *fb_mark, *fb_rows, *fb_users, *fb_src;
  int64_t bytes;
};

static TcWorkspace tc_workspace(void* base, int64_t n_rank, int64_t n_range, int64_t Kp, int32_t k, int n_splits) {
  int fs, fs2, ftps;
  tc_fb_split_plan(n_rank, n_range, 256, &fs, &ftps, Kp, k);
  tc_fb_split_plan(n_rank, n_range, 128, &fs2, &ftps, Kp, k);
  int64_t ns = std::max<int64_t>(n_splits, std::max(fs, fs2));
  if (Kp > 128) {
    int ss, stps;
    tc_stream_split_plan(n_rank, n_range, Kp, 256, &ss, &stps);
    ns = std::max<int64_t>(ns, ss);
  }
  TcWorkspace w;
  char* p = (char*)base;
  auto take = [&](int64_t bytes) {
    char* r = p;
    p += al256(bytes);
    return r;
  };
  w.u2 = (float*)take(n_rank * 2 * Kp * 4);
  w.i2 = (float*)take(n_range * 2 * Kp * 4);
  w.ctrl = (int*)take(256 + n_rank * 4);
  w.fb_mark = w.ctrl + 64;
  w.fb_rows = (int*)take(n_rank * 4);
  w.fb_users = (int*)take(n_rank * 4);
  w.fb_src = (int*)take(n_rank * 4);
  w.part_ids = (int*)take(ns * n_rank * k * 4);
  w.part_scores = (float*)take(ns * n_rank * k * 4);
  w.bytes = (int64_t)(p - (char*)base) + 256;
  return w;
}

int64_t eval_tc_workspace_bytes(int64_t n_rank, int64_t n_range, int64_t K, int32_t k, bool has_bias, int n_splits) {
  return tc_workspace(nullptr, n_rank, n_range, tc_padded_k(K, has_bias), k, n_splits).bytes;
}


template <int BN, int KL, int EW, bool ST, int CTAS, bool SCREEN, int NACC = 2, bool INS = false>
static int tc_launch(dim3 grid, size_t smem, cudaStream_t s, const CUtensorMap& map_u, const CUtensorMap& map_i, const TcArgs& a) {
  auto kern = eval_topk_tc_kernel<BN, KL, EW, ST, CTAS, SCREEN, NACC, INS>;
  TGCN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(128 + 128 * EW + (INS ? 128 : 0));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CTAS > 1 ? 1 : 0;
  TGCN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, map_u, map_i, a));
  return 0;
}

static void tc_args_common(TcArgs* a, int64_t n_rank, int Kp, int64_t K, int64_t n_range, int64_t item_begin, int32_t k, int tps, int n_stages,
                           const int* users, const int* mrowptr, const int* mcol, int mrow_begin, int mcol_off, const TcWorkspace& w,
                           int n_splits, int finalize, int* d_out_ids, float* d_out_scores) {
  *a = TcArgs{};
  a->n_rank = (int)n_rank;
  a->K = Kp;
  a->Kr = (int)K;
  a->n_range = (int)n_range;
  a->item_begin = (int)item_begin;
  a->k = k;
  a->tiles_per_split = tps;
  a->n_stages = n_stages;
  a->users = users;
  a->mrowptr = mrowptr;
  a->mcol = mcol;
  a->mrow_begin = mrow_begin;
  a->mcol_off = mcol_off;
  a->part_ids = w.part_ids;
  a->part_scores = w.part_scores;
  a->direct = n_splits == 1;
  a->finalize = finalize;
  a->out_ids = d_out_ids;
  a->out_scores = d_out_scores;
}

// The 3xTF32 variant.  gate != nullptr: the second pass of the screened path — every kernel reads the queue length on the device.
int eval_topk_tc(const int* mrowptr, const int* mcol, int mrow_begin, int mcol_off, int64_t n_rank, const int32_t* d_users,
                 int by_pos, const float* d_user_vecs, int64_t ldu, const float* d_item_vecs, int64_t ldi, int64_t K, int64_t item_begin,
                 int64_t item_end, const float* d_user_bias, const float* d_item_bias, int32_t k, int finalize, int* d_out_ids,
                 float* d_out_scores, void* d_workspace, int64_t workspace_bytes, int* n_splits_out, int** part_ids_out,
                 float** part_scores_out, cudaStream_t s, const TcGate* gate) {
  const bool has_bias = d_user_bias != nullptr || d_item_bias != nullptr;
  const int Kp = tc_padded_k(K, has_bias);
  int bn, n_stages;
  size_t smem;
  bool stream, pair;
  // CTA pairs: long sweeps of the resident variant (pair_wanted), and the streamed variant whenever there are two user tiles — it
  // re-streams the whole item operand per user tile and is bound by L2 -> SM traffic, of which the pair halves the item share
  const bool want_pair = Kp > 128 ? (k <= 40 && n_rank > TC_BM && pair_mode() != 0) : pair_wanted(item_end - item_begin, k);
  tc_plan(Kp, &bn, &n_stages, &smem, &stream, want_pair, &pair);
  TGCN_REQUIRE(n_stages >= 2, "3xTF32 path does not fit in shared memory for K=%lld", (long long)K);
  {  // the list merge at the end of a sweep borrows the item ring: 128 rows x (score, id) x list capacity
    const size_t brows = (size_t)bn / (pair ? 2 : 1);
    const size_t stage_bytes = stream ? (size_t)2 * TC_A_CHUNK_BYTES + 2 * brows * 128 : brows * 128;
    TGCN_REQUIRE((size_t)n_stages * stage_bytes >= (size_t)TC_BM * 2 * 40 * 4, "item ring too small for the list merge");
  }
  const int64_t n_range = item_end - item_begin;
  int n_splits, tps;
  if (gate) tc_fb_split_plan(n_rank, n_range, bn, &n_splits, &tps, Kp, k);
  else if (stream) tc_stream_split_plan(n_rank, n_range, Kp, bn, &n_splits, &tps);
  else eval_split_plan(n_rank, n_range, bn, &n_splits, &tps);
  const TcWorkspace w = tc_workspace(d_workspace, n_rank, n_range, Kp, k, n_splits);
  TGCN_REQUIRE(d_workspace && workspace_bytes >= w.bytes, "workspace too small: need %lld bytes", (long long)w.bytes);
  float *u2 = w.u2, *i2 = w.i2;
  const int k4 = Kp / 4;
  // user operand: rows gathered by user id unless already packed in list order; user bias follows the same indexing
  const int* urows = gate ? gate->src : (by_pos ? nullptr : d_users);
  tf32_split_kernel<<<(unsigned)((n_rank * k4 + 255) / 256), 256, 0, s>>>(d_user_vecs, ldu, urows, 0, n_rank, (int)K, Kp, d_user_bias, urows, 0, 0,
                                                                         has_bias ? 1 : 0, u2, gate ? gate->count : nullptr, 1);
  TGCN_CHECK_LAUNCH();
  tf32_split_kernel<<<(unsigned)((n_range * k4 + 255) / 256), 256, 0, s>>>(d_item_vecs, ldi, nullptr, item_begin, n_range, (int)K, Kp, d_item_bias,
                                                                          nullptr, item_begin, 1, has_bias ? 1 : 0, i2,
                                                                          gate ? gate->count : nullptr, 0);
  TGCN_CHECK_LAUNCH();
  CUtensorMap map_u, map_i;
  if (int rc = make_map(&map_u, u2, n_rank, 2 * (int64_t)Kp, TC_BM)) return rc;
  if (int rc = make_map(&map_i, i2, n_range, 2 * (int64_t)Kp, pair ? bn / 2 : bn)) return rc;
  TcArgs a;
  tc_args_common(&a, n_rank, Kp, K, n_range, item_begin, k, tps, n_stages, gate ? gate->users : d_users, mrowptr, mcol, mrow_begin, mcol_off, w,
                 n_splits, finalize, d_out_ids, d_out_scores);
  if (gate) {
    a.n_rank_dev = gate->count;
    a.out_rows = gate->rows;
  }
  a.split_fastest = (stream && n_splits > 1 && stream_raster_enabled()) ? 1 : 0;
  dim3 grid((unsigned)((n_rank + TC_BM - 1) / TC_BM), (unsigned)n_splits);
  if (pair) grid.x = (grid.x + 1) / 2 * 2;  // whole CTA pairs (a trailing CTA without users only lends its half of the item tile)
  int rc;
  if (pair && stream) rc = k <= 20 ? tc_launch<256, 20, 2, true, 2, false>(grid, smem, s, map_u, map_i, a) : tc_launch<256, 40, 2, true, 2, false>(grid, smem, s, map_u, map_i, a);
  else if (pair) rc = k <= 20 ? tc_launch<256, 20, 2, false, 2, false>(grid, smem, s, map_u, map_i, a) : tc_launch<256, 40, 2, false, 2, false>(grid, smem, s, map_u, map_i, a);
  else if (stream && bn == 256)
    rc = k <= 20   ? tc_launch<256, 20, 2, true, 1, false>(grid, smem, s, map_u, map_i, a)
         : k <= 40 ? tc_launch<256, 40, 2, true, 1, false>(grid, smem, s, map_u, map_i, a)
                   : tc_launch<256, 64, 1, true, 1, false>(grid, smem, s, map_u, map_i, a);
  else if (stream)
    rc = k <= 20   ? tc_launch<128, 20, 2, true, 1, false>(grid, smem, s, map_u, map_i, a)
         : k <= 40 ? tc_launch<128, 40, 2, true, 1, false>(grid, smem, s, map_u, map_i, a)
                   : tc_launch<128, 64, 1, true, 1, false>(grid, smem, s, map_u, map_i, a);
  else if (bn == 256)
    rc = k <= 20   ? tc_launch<256, 20, 2, false, 1, false>(grid, smem, s, map_u, map_i, a)
         : k <= 40 ? tc_launch<256, 40, 2, false, 1, false>(grid, smem, s, map_u, map_i, a)
                   : tc_launch<256, 64, 1, false, 1, false>(grid, smem, s, map_u, map_i, a);
  else
    rc = k <= 20   ? tc_launch<128, 20, 2, false, 1, false>(grid, smem, s, map_u, map_i, a)
         : k <= 40 ? tc_launch<128, 40, 2, false, 1, false>(grid, smem, s, map_u, map_i, a)
                   : tc_launch<128, 64, 1, false, 1, false>(grid, smem, s, map_u, map_i, a);
  if (rc) return rc;
  TGCN_CHECK_LAUNCH();
  *n_splits_out = n_splits;
  *part_ids_out = w.part_ids;
  *part_scores_out = w.part_scores;
  return 0;
}

// ---- screened variant -------------------------------------------------------------------------------------------------
constexpr int kScreenKL = 40;       // list capacity of the screen
constexpr int kScreenMaxK = 24;     // ... which leaves at least 16 entries of margin below the k-th best
// |s_tf32 - s| <= eps_c |u||i|: the user operand (raw fp32, so that the resident tile also serves the exact re-scoring) loses at most
// 2^-10 of each element to the tensor core's 19-bit read, the item operand (rounded to nearest beforehand) 2^-11: 1.47e-3 on a product
// with the cross term; the fp32 accumulation of the exact products adds < 4e-5 at K <= 128 (the 3xTF32 variant, same accumulation,
// agrees with fp64 to ~1e-6); 1.6e-3 keeps 6 % in hand, and wide contractions (K = 1600) get 1.7e-3.  Bias terms ride in one extra
// chunk as [ub, 1] x [1, ib]: ub·1 loses at most 2^-10 |ub| (1.1e-3 with slack), 1·ib at most 2^-11 |ib| (5.5e-4).
// TGCN_EVAL_SCREEN_EPS overrides eps_c (experiments: a value too small makes the parity tests fail, a large one sends every row to the
// second pass).
static float screen_eps_c(int64_t K) {
  static const float c = [] {
    const char* e = getenv("TGCN_EVAL_SCREEN_EPS");
    const float v = e ? (float)atof(e) : 0.f;
    return v > 0.f ? v : 0.f;
  }();
  return c > 0.f ? c : (K <= 128 ? 1.6e-3f : 1.7e-3f);
}

bool eval_tc_screen_eligible(int64_t K, int32_t k, bool has_bias) {
  (void)has_bias;
  return K % 4 == 0 && K > 0 && K <= 8192 && k <= kScreenMaxK;
}

// The streamed form (bias terms or K > 128) against the resident one
static inline bool screen_streams(int64_t K, bool has_bias) { return has_bias || tc_padded_k(K, false) > 128; }

// precision 0 takes the screened path for long item sweeps only: a short sweep is dominated by its opening, where every item beats
// an empty list, and the 40-entry list costs more there than the two TF32 products saved (c2, 63 k items, K = 64: 10.2 ms screened,
// 8.6 ms 3xTF32; c5, 2 M items, K = 128: 107.7 ms against 267.2 ms).  The 3xTF32 sweep grows with K, the screened one much less, so the
// break-even moves: measured (75 776 users, random embeddings, tools/screen_crossover.py) at ~50 k items for K = 128 and ~100 k for
// K = 64.  The STREAMED screened form (bias terms or K > 128: raw K-chunks of both operands travel through the ring, chunk-major
// copies, exact re-scoring from global user rows) is built, tested and available as precision 3, but precision 0 does not take it: at
// the LTR shape (K = 1600 + bias chunk, 63 k items, 18 944 users, random tables) it runs 11.1 ms against 14.4 ms for 3xTF32, a thin
// margin that a full second pass would more than eat, and near-duplicate items (same text, slightly different graph embedding) sit
// inside an error band that the wide text part of the operand makes large.  TGCN_EVAL_SCREEN = 0 / 1 (read once) forces the
// screened path off / on where eligible.
static int64_t screen_min_items(int64_t K, bool has_bias) {
  if (screen_streams(K, has_bias)) return INT64_MAX;
  const int Kp = tc_padded_k(K, false);
  return Kp >= 128 ? 65536 : Kp >= 96 ? 98304 : 131072;
}
bool eval_tc_screen_auto(int64_t n_range, int64_t K, int32_t k, bool has_bias) {
  static const int mode = [] {
    const char* e = getenv("TGCN_EVAL_SCREEN");
    return e ? (atoi(e) != 0 ? 1 : 0) : -1;
  }();
  if (!eval_tc_screen_eligible(K, k, has_bias) || mode == 0) return false;
  return mode == 1 || n_range >= screen_min_items(K, has_bias);
}

int eval_tc_screen_rescore(const TcGate* gate, int64_t n_rank, const int32_t* d_users, int by_pos, const float* d_user_vecs, int64_t ldu,
                           const float* d_item_vecs, int64_t ldi, int64_t K, int32_t k, const float* d_user_bias, const float* d_item_bias,
                           int* d_out_ids, float* d_out_scores, cudaStream_t s) {
  fb_rescore_kernel<<<(unsigned)((n_rank * 32 + 255) / 256), 256, 0, s>>>(gate->rows, gate->count, (int)n_rank, d_users, by_pos, d_user_vecs, ldu,
                                                                        d_item_vecs, ldi, (int)K, k, d_user_bias, d_item_bias, d_out_ids,
                                                                        d_out_scores);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int64_t eval_tc_screen_queue_offset(int64_t n_rank, int64_t n_range, int64_t K, int32_t k, bool has_bias) {
  if (!eval_tc_screen_eligible(K, k, has_bias)) return -1;
  const TcWorkspace w = tc_workspace(nullptr, n_rank, n_range, tc_padded_k(K, has_bias), k, 1);
  return (int64_t)((char*)w.ctrl - (char*)nullptr);
}

// Screen pass: approximate sweep + exact re-scoring + certificate; rows that fail are queued (gate_out describes the queue).
int eval_topk_screen(const int* mrowptr, const int* mcol, int mrow_begin, int mcol_off, int64_t n_rank, const int32_t* d_users, int by_pos,
                     const float* d_user_vecs, int64_t ldu, const float* d_item_vecs, int64_t ldi, int64_t K, int64_t item_begin,
                     int64_t item_end, const float* d_user_bias, const float* d_item_bias, int32_t k, int finalize, int* d_out_ids,
                     float* d_out_scores, void* d_workspace, int64_t workspace_bytes, int* n_splits_out, int** part_ids_out,
                     float** part_scores_out, TcGate* gate_out, cudaStream_t s) {
  const bool has_bias = d_user_bias != nullptr || d_item_bias != nullptr;
  const bool stream = screen_streams(K, has_bias);
  const int Kp = tc_padded_k(K, has_bias);  // (the bias chunk only exists in the streamed form)
  const int KC = Kp / TC_CHUNK;
  const int64_t n_range = item_end - item_begin;
  const int pm = pair_mode();
  const bool pair = pm < 0 ? n_rank > TC_BM : pm == 1;
  const int bn = 256;
  const size_t brows = (size_t)bn / (pair ? 2 : 1);
  const size_t a_bytes = stream ? 0 : (size_t)KC * TC_A_CHUNK_BYTES;
  const size_t fixed = 1024 + a_bytes + kBarBlockBytes + kInsBytes;
  const size_t stage = stream ? (size_t)TC_A_CHUNK_BYTES + brows * 128 : brows * 128;
  int n_stages = (int)((227 * 1024 - fixed) / stage);
  if (n_stages > TC_MAX_STAGES) n_stages = TC_MAX_STAGES;
  const size_t smem = fixed + (size_t)n_stages * stage;
  TGCN_REQUIRE(n_stages >= 2 && (size_t)n_stages * stage >= (size_t)TC_BM * 3 * kScreenKL * 4 + TC_BM * 4, "item ring too small");
  int n_splits, tps;
  if (stream) tc_stream_split_plan(n_rank, n_range, Kp, bn, &n_splits, &tps, 1);
  else eval_split_plan(n_rank, n_range, bn, &n_splits, &tps);
  const TcWorkspace w = tc_workspace(d_workspace, n_rank, n_range, Kp, k, n_splits);
  TGCN_REQUIRE(d_workspace && workspace_bytes >= w.bytes, "workspace too small: need %lld bytes", (long long)w.bytes);
  TGCN_CHECK_CUDA(cudaMemsetAsync(w.ctrl, 0, 256 + (size_t)n_rank * 4, s));
  screen_prep_items_kernel<<<148 * 8, 256, 0, s>>>(d_item_vecs, ldi, item_begin, n_range, (int)K, Kp, d_item_bias, has_bias ? 1 : 0, stream ? 1 : 0,
                                                  w.i2, (unsigned*)(w.ctrl + 1), (unsigned*)(w.ctrl + 2));
  TGCN_CHECK_LAUNCH();
  const float* uptr = d_user_vecs;
  int64_t uld = ldu, ucols = K;
  if (stream) {  // the raw operand with zero pad and bias chunk, rows in list order
    screen_prep_users_kernel<<<(unsigned)((n_rank * (Kp / 4) + 255) / 256), 256, 0, s>>>(d_user_vecs, ldu, (by_pos || !d_users) ? nullptr : d_users,
                                                                                        n_rank, (int)K, Kp, d_user_bias, has_bias ? 1 : 0, 1, w.u2);
    TGCN_CHECK_LAUNCH();
    uptr = w.u2;
    uld = ucols = Kp;
  } else if (!by_pos && d_users) {  // rows gathered by user id into list order
    gather_rows_kernel<<<(unsigned)((n_rank * (K / 4) + 255) / 256), 256, 0, s>>>(d_user_vecs, ldu, d_users, n_rank, (int)K, w.u2);
    TGCN_CHECK_LAUNCH();
    uptr = w.u2;
    uld = K;
  }
  CUtensorMap map_u, map_i;
  if (stream) {  // chunk-major operands: 2-D maps over [KC · rows][32]
    if (int rc = make_map(&map_u, w.u2, (int64_t)KC * n_rank, TC_CHUNK, TC_BM, TC_CHUNK)) return rc;
    if (int rc = make_map(&map_i, w.i2, (int64_t)KC * n_range, TC_CHUNK, (int)brows, TC_CHUNK)) return rc;
  } else {
    if (int rc = make_map(&map_u, uptr, n_rank, ucols, TC_BM, uld)) return rc;
    if (int rc = make_map(&map_i, w.i2, n_range, Kp, (int)brows, Kp)) return rc;
  }
  TcArgs a;
  tc_args_common(&a, n_rank, Kp, K, n_range, item_begin, k, tps, n_stages, d_users, mrowptr, mcol, mrow_begin, mcol_off, w, n_splits, finalize,
                 d_out_ids, d_out_scores);
  a.ivec = d_item_vecs;
  a.ldi = ldi;
  a.eps_c = screen_eps_c(K);
  a.max_inorm2 = (const unsigned*)(w.ctrl + 1);
  a.ug = stream ? w.u2 : nullptr;
  a.chunk_major = stream ? 1 : 0;
  a.ibias = d_item_bias;
  a.has_bias = has_bias ? 1 : 0;
  a.eps_ub = 1.1e-3f;
  a.eps_ib = 5.5e-4f;
  a.max_ib = (const unsigned*)(w.ctrl + 2);
  a.fb_mark = w.fb_mark;
  a.fb_rows = w.fb_rows;
  a.fb_count = w.ctrl;
  a.split_fastest = (stream && n_splits > 1 && stream_raster_enabled()) ? 1 : 0;
  dim3 grid((unsigned)((n_rank + TC_BM - 1) / TC_BM), (unsigned)n_splits);
  if (pair) grid.x = (grid.x + 1) / 2 * 2;
  int rc;
  if (stream) rc = pair ? tc_launch<256, kScreenKL, 2, true, 2, true, 2, true>(grid, smem, s, map_u, map_i, a) : tc_launch<256, kScreenKL, 2, true, 1, true, 2, true>(grid, smem, s, map_u, map_i, a);
  else rc = pair ? tc_launch<256, kScreenKL, 2, false, 2, true, 2, true>(grid, smem, s, map_u, map_i, a) : tc_launch<256, kScreenKL, 2, false, 1, true, 2, true>(grid, smem, s, map_u, map_i, a);
  if (rc) return rc;
  TGCN_CHECK_LAUNCH();
  fb_index_kernel<<<(unsigned)((n_rank + 255) / 256), 256, 0, s>>>(w.fb_rows, w.ctrl, d_users, by_pos, (int)n_rank, w.fb_users, w.fb_src);
  TGCN_CHECK_LAUNCH();
  gate_out->count = w.ctrl;
  gate_out->rows = w.fb_rows;
  gate_out->users = w.fb_users;
  gate_out->src = w.fb_src;
  *n_splits_out = n_splits;
  *part_ids_out = w.part_ids;
  *part_scores_out = w.part_scores;
  return 0;
}

}  // namespace tgcn
