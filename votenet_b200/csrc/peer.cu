// One-sided all-gather of the per-rank detection records over NVLink peer memory.
//
// The reference has no multi-GPU path at all; BASELINE.json adds "one all-gather of per-cloud
// detections".  With NCCL that is a rendezvous: a collective kernel on every rank holds SMs while it waits for the
// slowest peer, once per forward, and all forwards' collectives serialise on one communicator.  Here the PRODUCER
// pushes: when a rank's forward has written its record, one small kernel stores the record straight into every peer's
// inbox (P2P stores over NVLink/NVSwitch; the buffers are CUDA-IPC mappings of each rank's own allocation) and then
// raises a per-(slot, source) sequence flag with a system-scope release.  The consumer side is a one-warp kernel that
// waits for the `world` flags of its slot (acquire loads on LOCAL memory) in front of vnb_merge_detections.  No rank
// ever waits inside a communication kernel for data that has not been produced yet, forwards of different slots do not
// serialise against each other, and the transfer (world x 0.3 MB per forward) overlaps the other streams' compute.
#include "common.cuh"

#include <string.h>

namespace vnb {

constexpr int PEER_MAX = 16;
struct PeerPtrs {
  unsigned char* inbox[PEER_MAX];   // peer p: base of the slot's (world, nbytes) gather buffer in p's memory
  int* flags[PEER_MAX];             // peer p: base of the slot's (world) flag array in p's memory
};

__global__ void __launch_bounds__(1024) peer_push_kernel(PeerPtrs pp, int rank, const uint4* __restrict__ record,
                                                          size_t nvec, int seq) {
  const int p = blockIdx.x;   // destination rank (including this one: the local copy)
  uint4* dst = reinterpret_cast<uint4*>(pp.inbox[p]) + (size_t)rank * nvec;
  for (size_t i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = record[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();   // the record is visible system-wide before the flag
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(pp.flags[p] + rank), "r"(seq) : "memory");
  }
}

__global__ void peer_wait_kernel(int world, const int* __restrict__ flags, int seq) {
  const int r = threadIdx.x;
  if (r < world) {
    int v;
    unsigned spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
      if (v - seq >= 0) break;
      __nanosleep(200);
      if (++spins > (1u << 27)) trap_at(__LINE__);   // a peer died: trap instead of hanging the GPU (after ~30 s)
    } while (true);
  }
}

}  // namespace vnb

using namespace vnb;

extern "C" int vnb_peer_push_record(int world, int rank, const void* record, size_t nbytes, void* const* peer_inbox,
                                    int* const* peer_flags, int seq, void* stream) {
  VNB_REQUIRE(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world, "peer_push_record: 1 <= world <= 16, 0 <= rank < world");
  VNB_REQUIRE(nbytes % 16 == 0 && (reinterpret_cast<uintptr_t>(record) & 15) == 0, "peer_push_record: record must be 16-byte aligned and sized");
  PeerPtrs pp;
  for (int p = 0; p < world; ++p) {
    pp.inbox[p] = static_cast<unsigned char*>(peer_inbox[p]);
    pp.flags[p] = peer_flags[p];
  }
  peer_push_kernel<<<world, 1024, 0, as_stream(stream)>>>(pp, rank, static_cast<const uint4*>(record), nbytes / 16, seq);
  return check_launch("peer_push_record");
}

extern "C" int vnb_peer_alloc(size_t nbytes, void** dev_ptr, unsigned char* ipc_handle_64) {
  VNB_REQUIRE(nbytes > 0 && dev_ptr != nullptr && ipc_handle_64 != nullptr, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  VNB_CUDA(cudaMalloc(&p, nbytes));          // a dedicated allocation: the handle maps exactly this buffer, offset 0
  VNB_CUDA(cudaMemset(p, 0, nbytes));
  VNB_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  VNB_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(ipc_handle_64, &h, 64);
  *dev_ptr = p;
  return VNB_OK;
}

extern "C" int vnb_peer_open(const unsigned char* ipc_handle_64, void** dev_ptr) {
  VNB_REQUIRE(dev_ptr != nullptr && ipc_handle_64 != nullptr, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64, 64);
  // opened in the CURRENT device's context: CUDA sets up the peer mapping to the exporting GPU (NVLink) for this device
  VNB_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return VNB_OK;
}

extern "C" int vnb_peer_close(void* dev_ptr) {
  VNB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return VNB_OK;
}

extern "C" int vnb_peer_free(void* dev_ptr) {
  VNB_CUDA(cudaFree(dev_ptr));
  return VNB_OK;
}

extern "C" int vnb_peer_enable_access(int peer_device) {
  int cur = -1;
  VNB_CUDA(cudaGetDevice(&cur));
  if (peer_device == cur) return VNB_OK;
  int can = 0;
  VNB_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device));
  VNB_REQUIRE(can != 0, "peer_enable_access: device %d cannot access device %d over NVLink/PCIe peer-to-peer", cur, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return VNB_OK; }
  if (e != cudaSuccess) return set_err(VNB_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
  return VNB_OK;
}

extern "C" int vnb_peer_wait(int world, const int* flags, int seq, void* stream) {
  VNB_REQUIRE(world >= 1 && world <= PEER_MAX, "peer_wait: 1 <= world <= 16");
  peer_wait_kernel<<<1, 32, 0, as_stream(stream)>>>(world, flags, seq);
  return check_launch("peer_wait");
}
