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

int tgcn_peer_close(void* d_ptr) {
  if (d_ptr) TGCN_CHECK_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return 0;
}

int tgcn_peer_free(void* d_ptr) {
  if (d_ptr) TGCN_CHECK_CUDA(cudaFree(d_ptr));
  return 0;
}

}  // extern "C"
