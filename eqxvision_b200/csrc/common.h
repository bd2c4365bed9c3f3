// Host-side helpers shared by the translation units of libeqxv_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/eqxv_b200.h"

namespace eqxv {

// thread-local last-error text, returned by eqxv_last_error()
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define EQXV_CHECK_ARG(cond, ...)          \
  do {                                     \
    if (!(cond)) {                         \
      ::eqxv::set_error(__VA_ARGS__);      \
      return EQXV_ERR_INVALID_ARGUMENT;    \
    }                                      \
  } while (0)

#define EQXV_CUDA(call)                                           \
  do {                                                            \
    cudaError_t _e = (call);                                      \
    if (_e != cudaSuccess) return ::eqxv::cuda_fail(_e, #call);   \
  } while (0)

// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency).
struct TmapSpec {
  void* base;
  CUtensorMapDataType dtype;
  uint32_t rank;
  uint64_t dims[5];
  uint64_t strides_bytes[4];  // strides of dims 1..rank-1
  uint32_t box[5];
  uint32_t estride[5];
  CUtensorMapSwizzle swizzle;
};
int encode_tmap(CUtensorMap* out, const TmapSpec& s);

int device_sm_count();

// Kernel launch with the programmatic-dependent-launch attribute (see ptx.cuh: griddep_wait). Works
// under stream capture (the edge becomes a programmatic graph dependency). EQXV_NO_PDL=1 switches the
// attribute off (A/B measurements); the griddepcontrol instructions are no-ops then.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace eqxv
