// api.cu — library-level entry points and error plumbing.
#include <cstdarg>
#include <cstdio>

#include <cuda.h>

#include "common.cuh"

namespace w2t {

static thread_local char g_last_error[512] = "";

void set_last_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof g_last_error, fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
  set_last_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return W2T_ERR_CUDA;
}

}  // namespace w2t

extern "C" const char *w2t_version(void) { return "w2t 0.1 (sm_100a)"; }

extern "C" const char *w2t_last_error(void) { return w2t::g_last_error; }

extern "C" void w2t_clear_error(void) { w2t::g_last_error[0] = '\0'; }

extern "C" int w2t_device_info(int *sm_count, int *cc_major, int *cc_minor) {
  int dev = 0;
  W2T_CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  W2T_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return W2T_OK;
}

extern "C" int w2t_stream_wait_value32(w2t_stream_t stream, const int32_t *addr, int32_t value) {
  // the driver entry point is looked up at run time, so the library does not link against libcuda
  typedef CUresult (*wait_fn_t)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
  static wait_fn_t fn = nullptr;
  if (fn == nullptr) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    W2T_CUDA_TRY(cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || p == nullptr) {
      w2t::set_last_error("w2t_stream_wait_value32: cuStreamWaitValue32 is not available");
      return W2T_ERR_CUDA;
    }
    fn = reinterpret_cast<wait_fn_t>(p);
  }
  const CUresult r = fn(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(addr), (cuuint32_t)value,
                        CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) {
    w2t::set_last_error("w2t_stream_wait_value32: cuStreamWaitValue32 failed with %d", (int)r);
    return W2T_ERR_CUDA;
  }
  return W2T_OK;
}
