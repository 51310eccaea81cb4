// (e) multi-GPU: the path's exchange steps as C-ABI helpers that take an ncclComm_t (SURVEY.md §8b "minimum export set").
//
// The reference has no distributed code; the B200-native design adds two exchanges:
//   * per hop, the all-reduce of the item-table slice inside a row group   -> tgcn_allreduce_sum_f32
//     (+ the all-gather of row blocks for the north_star's row-block scheme -> tgcn_allgather_f32),
//   * item-sharded eval: the all-to-all of the partial (U, k) top-k tables  -> tgcn_topk_exchange,
// plus communicator plumbing for a consumer that does not run torch.distributed (tgcn_comm_unique_id / _init_rank /
// _destroy: the 128-byte unique id travels over any transport).  The helpers accept ANY ncclComm_t of the process as a
// void*, e.g. one the caller created itself.  NCCL is resolved at run time from the library already loaded in the
// process (torch's bundled libnccl.so.2) or the system one — libtgcn_b200.so does not link against it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "common.cuh"

namespace tgcn {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};

static const NcclApi& nccl() {
  static const NcclApi api = [] {
    NcclApi a;
    memset(&a, 0, sizeof(a));
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy torch (or the caller) already loaded
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
#define TGCN_NCCL_SYM(field, name) *(void**)(&a.field) = dlsym(h, name)
    TGCN_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    TGCN_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    TGCN_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    TGCN_NCCL_SYM(AllReduce, "ncclAllReduce");
    TGCN_NCCL_SYM(AllGather, "ncclAllGather");
    TGCN_NCCL_SYM(Send, "ncclSend");
    TGCN_NCCL_SYM(Recv, "ncclRecv");
    TGCN_NCCL_SYM(GroupStart, "ncclGroupStart");
    TGCN_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    TGCN_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef TGCN_NCCL_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather && a.Send && a.Recv && a.GroupStart &&
           a.GroupEnd && a.GetErrorString;
    return a;
  }();
  return api;
}

#define TGCN_NCCL_READY() TGCN_REQUIRE(nccl().ok, "NCCL (libnccl.so.2) could not be resolved in this process")
#define TGCN_CHECK_NCCL(expr)                                                                          \
  do {                                                                                                 \
    ncclResult_t _r = (expr);                                                                          \
    if (_r != ncclSuccess) {                                                                           \
      ::tgcn::set_error("%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(_r), __FILE__, __LINE__); \
      return 4;                                                                                        \
    }                                                                                                  \
  } while (0)

}  // namespace tgcn

using namespace tgcn;

static_assert(sizeof(ncclUniqueId) == TGCN_COMM_ID_BYTES, "ncclUniqueId size");

extern "C" {

int tgcn_comm_unique_id(uint8_t* h_id) {
  TGCN_REQUIRE(h_id != nullptr, "NULL id buffer");
  TGCN_NCCL_READY();
  ncclUniqueId id;
  TGCN_CHECK_NCCL(nccl().GetUniqueId(&id));
  memcpy(h_id, &id, sizeof(id));
  return 0;
}

int tgcn_comm_init_rank(void** comm, int32_t n_ranks, int32_t rank, const uint8_t* h_id) {
  TGCN_REQUIRE(comm && h_id && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad communicator arguments");
  TGCN_NCCL_READY();
  ncclUniqueId id;
  memcpy(&id, h_id, sizeof(id));
  ncclComm_t c = nullptr;
  TGCN_CHECK_NCCL(nccl().CommInitRank(&c, n_ranks, id, rank));
  *comm = (void*)c;
  return 0;
}

int tgcn_comm_destroy(void* comm) {
  if (!comm) return 0;
  TGCN_NCCL_READY();
  TGCN_CHECK_NCCL(nccl().CommDestroy((ncclComm_t)comm));
  return 0;
}

int tgcn_allreduce_sum_f32(void* comm, float* d_buf, int64_t n, tgcn_stream_t stream) {
  TGCN_REQUIRE(comm && d_buf && n > 0, "bad all-reduce arguments");
  TGCN_NCCL_READY();
  TGCN_CHECK_NCCL(nccl().AllReduce(d_buf, d_buf, (size_t)n, ncclFloat32, ncclSum, (ncclComm_t)comm, (cudaStream_t)stream));
  return 0;
}

int tgcn_allgather_f32(void* comm, const float* d_send, float* d_recv, int64_t n_per_rank, tgcn_stream_t stream) {
  TGCN_REQUIRE(comm && d_send && d_recv && n_per_rank > 0, "bad all-gather arguments");
  TGCN_NCCL_READY();
  TGCN_CHECK_NCCL(nccl().AllGather(d_send, d_recv, (size_t)n_per_rank, ncclFloat32, (ncclComm_t)comm, (cudaStream_t)stream));
  return 0;
}

int tgcn_topk_exchange(void* comm, int32_t n_ranks, int64_t rows_per_rank, int32_t k, const int32_t* d_part_ids,
                       const float* d_part_scores, int32_t* d_recv_ids, float* d_recv_scores, tgcn_stream_t stream) {
  TGCN_REQUIRE(comm && n_ranks >= 1 && rows_per_rank > 0 && k > 0, "bad exchange sizes");
  TGCN_REQUIRE(d_part_ids && d_part_scores && d_recv_ids && d_recv_scores, "NULL table");
  TGCN_NCCL_READY();
  const size_t n = (size_t)rows_per_rank * k;  // block q of the send tables = this rank's candidates for user slice q
  cudaStream_t s = (cudaStream_t)stream;
  ncclComm_t c = (ncclComm_t)comm;
  TGCN_CHECK_NCCL(nccl().GroupStart());
  for (int q = 0; q < n_ranks; ++q) {
    TGCN_CHECK_NCCL(nccl().Send(d_part_ids + q * n, n, ncclInt32, q, c, s));
    TGCN_CHECK_NCCL(nccl().Recv(d_recv_ids + q * n, n, ncclInt32, q, c, s));
    TGCN_CHECK_NCCL(nccl().Send(d_part_scores + q * n, n, ncclFloat32, q, c, s));
    TGCN_CHECK_NCCL(nccl().Recv(d_recv_scores + q * n, n, ncclFloat32, q, c, s));
  }
  TGCN_CHECK_NCCL(nccl().GroupEnd());
  return 0;
}

}  // extern "C"
