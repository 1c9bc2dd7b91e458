// Peer mailboxes over NVLink for the chunk scheduler (lsvs_b200/scheduler.py: PeerTransport).
//
// Why not NCCL send/recv here: a point-to-point NCCL kernel stays resident (spinning) on the SMs of the posting GPU until
// the peer posts the matching call.  The encoder's GEMM / attention kernels are persistent, one CTA per SM with ~220 KB of
// shared memory and a static tile schedule, so a communication CTA that holds an SM for milliseconds makes every such
// kernel wait for a second wave: the owner ranks ended up synchronised to the alignment rank once per round.
// Here the payload moves with the copy engines (cudaMemcpyAsync into an IPC-mapped buffer of the peer: no SM involved, no
// rendezvous), followed by a one-thread kernel that publishes a sequence number in the peer's memory; the consumer
// enqueues a one-warp, zero-shared-memory kernel that returns as soon as the number is there (already the case in steady
// state, because consumers run one or two rounds behind producers).
#include "host_common.h"
#include <cstring>

namespace {

// A mailbox header is two words: flag[0] = sequence number (messages published so far), flag[1] = poison (non-zero once the
// publishing rank has seen a failure).  `status` is the caller's health word (device or pinned host memory): 0 = healthy,
// 1 = a wait of this rank timed out, 2 = a peer published poison.
__global__ void peer_signal_kernel(unsigned* flag, unsigned value, const volatile unsigned* status) {
  __threadfence_system();
  if (status != nullptr && *status != 0u) {
    // this rank's inputs may be stale (an earlier wait gave up): tell the consumer instead of publishing results built on them
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag + 1), "r"(1u) : "memory");
    return;
  }
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Returns when (int)(flag[0] - value) >= 0.  Gives up — the stream keeps going, nothing hangs — when the producer has published
// poison (*status = 2), after `timeout_ns` (*status = 1), or at once if *status is already non-zero (a failed rank does not
// wait a full timeout per message again).  Every later lsvs_peer_signal of this rank then publishes poison, so one lost peer
// stops the whole group within a message instead of letting it compute on stale mailboxes.
__global__ void peer_wait_kernel(const unsigned* flag, unsigned value, volatile unsigned* status, unsigned long long timeout_ns) {
  if (threadIdx.x != 0) return;
  if (*status != 0u) return;
  const unsigned long long t0 = global_ns();
  unsigned backoff = 32;
  while ((int)(ld_acquire_sys(flag) - value) < 0) {
    if (ld_acquire_sys(flag + 1) != 0u) { *status = 2u; __threadfence_system(); return; }
    __nanosleep(backoff);
    if (backoff < 2048) backoff *= 2;
    if (global_ns() - t0 > timeout_ns) { *status = 1u; __threadfence_system(); return; }
  }
}

}  // namespace

extern "C" int lsvs_peer_alloc(size_t bytes, void** ptr) {
  LSVS_CHECK_ARG(ptr != nullptr && bytes > 0, "lsvs_peer_alloc: null ptr / zero size");
  void* p = nullptr;
  LSVS_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    return lsvs::fail(LSVS_ECUDA, "lsvs_peer_alloc: %s", cudaGetErrorString(e));
  }
  *ptr = p;
  return LSVS_OK;
}

extern "C" int lsvs_peer_free(void* ptr) {
  if (ptr) LSVS_CUDA(cudaFree(ptr));
  return LSVS_OK;
}

extern "C" int lsvs_peer_export(const void* ptr, unsigned char* handle64) {
  LSVS_CHECK_ARG(ptr != nullptr && handle64 != nullptr, "lsvs_peer_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == LSVS_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  LSVS_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle64, &h, sizeof(h));
  return LSVS_OK;
}

extern "C" int lsvs_peer_open(const unsigned char* handle64, void** ptr) {
  LSVS_CHECK_ARG(ptr != nullptr && handle64 != nullptr, "lsvs_peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  LSVS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return LSVS_OK;
}

extern "C" int lsvs_peer_close(void* ptr) {
  if (ptr) LSVS_CUDA(cudaIpcCloseMemHandle(ptr));
  return LSVS_OK;
}

extern "C" int lsvs_peer_put(void* dst, const void* src, size_t bytes, void* stream) {
  LSVS_CHECK_ARG(dst != nullptr && src != nullptr, "lsvs_peer_put: null pointer");
  if (bytes == 0) return LSVS_OK;
  LSVS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  return LSVS_OK;
}

extern "C" int lsvs_peer_signal(unsigned* flag, unsigned value, const unsigned* status, void* stream) {
  LSVS_CHECK_ARG(flag != nullptr, "lsvs_peer_signal: null flag");
  peer_signal_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(flag, value, status);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

extern "C" int lsvs_peer_wait(const unsigned* flag, unsigned value, unsigned* status, double timeout_s, void* stream) {
  LSVS_CHECK_ARG(flag != nullptr && status != nullptr, "lsvs_peer_wait: null flag / status");
  LSVS_CHECK_ARG(timeout_s > 0, "lsvs_peer_wait: timeout must be positive");
  peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(flag, value, status, (unsigned long long)(timeout_s * 1e9));
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
