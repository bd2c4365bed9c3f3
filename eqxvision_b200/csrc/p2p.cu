// C1: the one collective of the path: an all-gather of the fp32 logits across the GPUs of one box,
// issued only when the caller wants the gathered output (`vit.py:273` / `resnet.py:356` produce [B/N, classes] per
// rank). Hand-written over NVLink peer memory instead of a library call: every rank PUSHES its slice straight into
// every peer's gather buffer with 16-byte stores to CUDA-IPC mapped pointers, then raises a flag in each peer's flag
// array (st.release.sys) and waits for the N flags in its own (ld.acquire.sys). One kernel, no host round trip, no
// staging copy; 256 KB per rank in the BASELINE ViT config, so the cost is launch + one NVLink round trip.
//
// Memory of one rank ("window", allocated by eqxv_p2p_alloc, exported with eqxv_ipc_get_handle):
//   [0, 256)        uint32 flags[64]   flags[r] = last epoch whose slice from rank r has fully arrived here
//   [256, 512)      uint32 epoch, arrive (local bookkeeping: launches so far, CTAs finished in this launch)
//   [512, ...)      two gather buffers of `buf_bytes` each (epoch parity selects one: a peer may already push epoch
//                   e+1 while this rank still reads epoch e; it cannot push e+2 before this rank has signalled e+1,
//                   which is stream-ordered after its reads of epoch e)
#include <string.h>

#include "common.h"
#include "ptx.cuh"

namespace eqxv {

constexpr int kP2PHeader = 512;
constexpr int kP2PMaxWorld = 64;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct P2PPeers {
  uint8_t* window[kP2PMaxWorld];   // peer windows as mapped in THIS process (own window at index rank)
};

__global__ void allgather_push_kernel(const uint4* __restrict__ src, long long nvec, P2PPeers peers, int rank,
                                      int world, long long slot_bytes, long long buf_bytes) {
  griddep_wait();
  uint8_t* mine = peers.window[rank];
  uint32_t* book = reinterpret_cast<uint32_t*>(mine + 256);   // [0] epoch, [1] arrive
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(book) + 1;
  const long long off = kP2PHeader + (long long)(e & 1) * buf_bytes + (long long)rank * slot_bytes;
  for (int q = 0; q < world; ++q) {
    const int p = (rank + q) % world;       // start with the own copy, then walk the ring: spreads the link load
    uint4* dst = reinterpret_cast<uint4*>(peers.window[p] + off);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
         i += (long long)gridDim.x * blockDim.x)
      dst[i] = __ldg(src + i);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t done = atomicAdd(book + 1, 1u);
    if (done == gridDim.x - 1) {            // the last CTA: every store of this rank is visible system-wide
      __threadfence_system();
      book[1] = 0;
      for (int p = 0; p < world; ++p)
        st_release_sys(reinterpret_cast<uint32_t*>(peers.window[p]) + rank, e);
      const uint32_t* flags = reinterpret_cast<const uint32_t*>(mine);
      for (int p = 0; p < world; ++p)
        while ((int32_t)(ld_acquire_sys(flags + p) - e) < 0) __nanosleep(64);
      *reinterpret_cast<volatile uint32_t*>(book) = e;
      __threadfence();
    }
  }
}

}  // namespace eqxv

using namespace eqxv;

extern "C" int eqxv_p2p_window_bytes(int64_t buf_bytes, int64_t* total) {
  EQXV_CHECK_ARG(buf_bytes > 0 && buf_bytes % 16 == 0 && total, "p2p_window_bytes: buf_bytes must be a positive multiple of 16");
  *total = kP2PHeader + 2 * buf_bytes;
  return EQXV_OK;
}

extern "C" int eqxv_p2p_alloc(void** ptr, int64_t bytes) {
  EQXV_CHECK_ARG(ptr && bytes > 0, "p2p_alloc: bad arguments");
  void* p = nullptr;
  EQXV_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e, "cudaMemset");
  }
  EQXV_CUDA(cudaDeviceSynchronize());
  *ptr = p;
  return EQXV_OK;
}

extern "C" int eqxv_p2p_free(void* ptr) {
  EQXV_CUDA(cudaFree(ptr));
  return EQXV_OK;
}

extern "C" int eqxv_ipc_get_handle(const void* ptr, uint8_t handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  EQXV_CHECK_ARG(ptr && handle, "ipc_get_handle: bad arguments");
  cudaIpcMemHandle_t h;
  EQXV_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle, &h, 64);
  return EQXV_OK;
}

extern "C" int eqxv_ipc_open_handle(const uint8_t handle[64], void** ptr) {
  EQXV_CHECK_ARG(ptr && handle, "ipc_open_handle: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* p = nullptr;
  EQXV_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return EQXV_OK;
}

extern "C" int eqxv_ipc_close_handle(void* ptr) {
  EQXV_CUDA(cudaIpcCloseMemHandle(ptr));
  return EQXV_OK;
}

extern "C" int eqxv_allgather_push(const void* src, int64_t bytes, void* const* windows, int32_t rank, int32_t world,
                                   int64_t slot_bytes, int64_t buf_bytes, void* stream) {
  EQXV_CHECK_ARG(src && windows && world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world,
                 "allgather_push: bad arguments");
  EQXV_CHECK_ARG(bytes > 0 && bytes % 16 == 0 && slot_bytes % 16 == 0 && bytes <= slot_bytes &&
                     (int64_t)world * slot_bytes <= buf_bytes && ((uintptr_t)src & 15) == 0,
                 "allgather_push: slices must be 16-byte multiples that fit their slot");
  P2PPeers peers;
  for (int i = 0; i < world; ++i) {
    EQXV_CHECK_ARG(windows[i], "allgather_push: window %d is NULL", i);
    peers.window[i] = static_cast<uint8_t*>(windows[i]);
  }
  const long long nvec = bytes / 16;
  int blocks = (int)((nvec + 255) / 256);
  if (blocks > 32) blocks = 32;   // co-resident by a wide margin: the last CTA spins on the peers' flags
  EQXV_CUDA(launch_kernel(allgather_push_kernel, dim3(blocks), dim3(256), (size_t)0, (cudaStream_t)stream,
                          reinterpret_cast<const uint4*>(src), nvec, peers, rank, world, (long long)slot_bytes,
                          (long long)buf_bytes));
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}

// offset of gather buffer `parity` inside a window (the host reads the gathered rows from its OWN window)
extern "C" int eqxv_p2p_buffer_offset(int32_t parity, int64_t buf_bytes, int64_t* offset) {
  EQXV_CHECK_ARG(offset && (parity == 0 || parity == 1), "p2p_buffer_offset: bad arguments");
  *offset = kP2PHeader + (int64_t)parity * buf_bytes;
  return EQXV_OK;
}
