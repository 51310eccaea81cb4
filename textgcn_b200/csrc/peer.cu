// Peer-mapped device memory for the feature-sliced multi-GPU propagation (one process per GPU): plain cudaMalloc +
// CUDA IPC handles.  The handles travel between ranks through the host code's own transport (torch.distributed).
#include <string.h>

#include "common.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == TGCN_PEER_HANDLE_BYTES, "IPC handle size");

extern "C" {

int tgcn_peer_alloc(int64_t bytes, void** d_ptr, uint8_t* h_handle) {
  TGCN_REQUIRE(bytes > 0 && d_ptr && h_handle, "bad peer allocation request");
  void* p = nullptr;
  TGCN_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    tgcn::set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return 1;
  }
  memcpy(h_handle, &h, sizeof(h));
  *d_ptr = p;
  return 0;
}

int tgcn_peer_open(const uint8_t* h_handle, void** d_ptr) {
  TGCN_REQUIRE(h_handle && d_ptr, "NULL handle");
  cudaIpcMemHandle_t h;
  memcpy(&h, h_handle, sizeof(h));
  void* p = nullptr;
  TGCN_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *d_ptr = p;
  return 0;
}

}  // extern "C"

namespace tgcn {

struct PeerFlags {
  long long* flags[TGCN_MAX_PEERS];  // flags[q] = rank q's flag array (peer-mapped), one int64 slot per rank
};

// Thread q publishes `epoch` into slot `rank` of peer q's array (release, system scope: everything this GPU stored before
// — the previous kernels' peer stores — is visible to whoever observes the flag), then waits until peer q has published it
// into OUR array.  Epochs only grow, so a rank that is already one barrier ahead does no harm.  The wait is bounded: a peer
// that never arrives traps the kernel after ~30 s instead of hanging the GPU.
__global__ void peer_barrier_kernel(const PeerFlags f, int n_peers, int rank, long long epoch) {
  const int q = threadIdx.x;
  if (q >= n_peers) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(f.flags[q] + rank), "l"(epoch) : "memory");
  const long long* mine = f.flags[rank] + q;
  const long long t0 = clock64();
  long long seen;
  do {
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
    if (seen < epoch && clock64() - t0 > 60000000000ll) __trap();
  } while (seen < epoch);
}

}  // namespace tgcn

extern "C" {

int tgcn_peer_barrier(int32_t n_peers, int32_t rank, int64_t* const* h_peer_flags, int64_t epoch, tgcn_stream_t stream) {
  TGCN_REQUIRE(n_peers >= 1 && n_peers <= TGCN_MAX_PEERS && rank >= 0 && rank < n_peers && h_peer_flags && epoch > 0, "bad barrier arguments");
  tgcn::PeerFlags f;
  for (int q = 0; q < n_peers; ++q) {
    TGCN_REQUIRE(h_peer_flags[q] != nullptr, "NULL flag array %d", q);
    f.flags[q] = (long long*)h_peer_flags[q];
  }
  tgcn::peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, n_peers, rank, (long long)epoch);
  TGCN_CHECK_LAUNCH();
  return 0;
}

int tgcn_peer_close(void* d_ptr) {
  if (d_ptr) TGCN_CHECK_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return 0;
}

int tgcn_peer_free(void* d_ptr) {
  if (d_ptr) TGCN_CHECK_CUDA(cudaFree(d_ptr));
  return 0;
}

}  // extern "C"
