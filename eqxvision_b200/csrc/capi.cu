// Library context, error reporting, tensor-map encoding and stream/graph/event plumbing.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"

namespace eqxv {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return EQXV_ERR_CUDA;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_sm_count = 0;
static int g_device = -1;

int device_sm_count() { return g_sm_count; }

bool pdl_enabled() {
  static const bool on = getenv("EQXV_NO_PDL") == nullptr;
  return on;
}

int encode_tmap(CUtensorMap* out, const TmapSpec& s) {
  if (!g_encode) {
    set_error("eqxv_init() was not called (cuTensorMapEncodeTiled unresolved)");
    return EQXV_ERR_NO_DEVICE;
  }
  cuuint64_t dims[5];
  cuuint64_t strides[4];
  cuuint32_t box[5], es[5];
  for (uint32_t i = 0; i < s.rank; ++i) {
    dims[i] = s.dims[i];
    box[i] = s.box[i];
    es[i] = s.estride[i];
  }
  for (uint32_t i = 0; i + 1 < s.rank; ++i) strides[i] = s.strides_bytes[i];
  const CUresult r = g_encode(out, s.dtype, s.rank, s.base, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, s.swizzle,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error(
        "cuTensorMapEncodeTiled failed (%d): rank %u dims [%llu,%llu,%llu,%llu] strides "
        "[%llu,%llu,%llu] box [%u,%u,%u,%u] estride [%u,%u,%u,%u] base %p",
        (int)r, s.rank, (unsigned long long)s.dims[0], (unsigned long long)s.dims[1],
        (unsigned long long)s.dims[2], (unsigned long long)s.dims[3],
        (unsigned long long)s.strides_bytes[0], (unsigned long long)s.strides_bytes[1],
        (unsigned long long)s.strides_bytes[2], s.box[0], s.box[1], s.box[2], s.box[3], s.estride[0],
        s.estride[1], s.estride[2], s.estride[3], s.base);
    return EQXV_ERR_CUDA;
  }
  return EQXV_OK;
}

int igemm_init();
int attention_init();
int bottleneck_init();
int gemv_init();

}  // namespace eqxv

using namespace eqxv;

extern "C" const char* eqxv_version(void) { return "eqxv_b200 0.1 (sm_100a)"; }
extern "C" const char* eqxv_last_error(void) { return g_err; }
extern "C" int eqxv_sm_count(void) { return g_sm_count; }

extern "C" int eqxv_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("no CUDA device visible (%s): this library has no CPU fallback",
              cudaGetErrorString(e));
    return EQXV_ERR_NO_DEVICE;
  }
  EQXV_CHECK_ARG(device >= 0 && device < count, "eqxv_init: device %d out of range [0,%d)", device,
                 count);
  EQXV_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  EQXV_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
              prop.minor);
    return EQXV_ERR_UNSUPPORTED;
  }
  g_sm_count = prop.multiProcessorCount;
  g_device = device;
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    EQXV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return EQXV_ERR_UNSUPPORTED;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  int rc = igemm_init();
  if (rc) return rc;
  rc = bottleneck_init();
  if (rc) return rc;
  rc = gemv_init();
  if (rc) return rc;
  rc = attention_init();
  if (rc) return rc;
  return EQXV_OK;
}

// ---------------------------------------------------------------------------------------------
// plumbing
// ---------------------------------------------------------------------------------------------
extern "C" int eqxv_stream_create(void** stream) {
  cudaStream_t s;
  EQXV_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = s;
  return EQXV_OK;
}
extern "C" int eqxv_stream_destroy(void* stream) {
  EQXV_CUDA(cudaStreamDestroy((cudaStream_t)stream));
  return EQXV_OK;
}
extern "C" int eqxv_stream_sync(void* stream) {
  EQXV_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return EQXV_OK;
}
extern "C" int eqxv_graph_begin(void* stream) {
  EQXV_CUDA(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
  return EQXV_OK;
}
extern "C" int eqxv_graph_end(void* stream, void** graph_exec) {
  cudaGraph_t g = nullptr;
  EQXV_CUDA(cudaStreamEndCapture((cudaStream_t)stream, &g));
  cudaGraphExec_t ge = nullptr;
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate");
  *graph_exec = ge;
  return EQXV_OK;
}
extern "C" int eqxv_graph_launch(void* graph_exec, void* stream) {
  EQXV_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
  return EQXV_OK;
}
extern "C" int eqxv_graph_destroy(void* graph_exec) {
  EQXV_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  return EQXV_OK;
}
extern "C" int eqxv_event_create(void** ev) {
  cudaEvent_t e;
  EQXV_CUDA(cudaEventCreate(&e));
  *ev = e;
  return EQXV_OK;
}
extern "C" int eqxv_event_destroy(void* ev) {
  EQXV_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return EQXV_OK;
}
extern "C" int eqxv_event_record(void* ev, void* stream) {
  EQXV_CUDA(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream));
  return EQXV_OK;
}
extern "C" int eqxv_event_sync(void* ev) {
  EQXV_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
  return EQXV_OK;
}
extern "C" int eqxv_stream_wait_event(void* stream, void* ev) {
  EQXV_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0));
  return EQXV_OK;
}
extern "C" int eqxv_event_elapsed_ms(void* start, void* stop, float* ms) {
  EQXV_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return EQXV_OK;
}
extern "C" int eqxv_memcpy_h2d_async(void* dst, const void* src, int64_t bytes, void* stream) {
  EQXV_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return EQXV_OK;
}
extern "C" int eqxv_memcpy_d2h_async(void* dst, const void* src, int64_t bytes, void* stream) {
  EQXV_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return EQXV_OK;
}
extern "C" int eqxv_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream) {
  EQXV_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return EQXV_OK;
}
extern "C" int eqxv_memset_async(void* dst, int value, int64_t bytes, void* stream) {
  EQXV_CUDA(cudaMemsetAsync(dst, value, (size_t)bytes, (cudaStream_t)stream));
  return EQXV_OK;
}
