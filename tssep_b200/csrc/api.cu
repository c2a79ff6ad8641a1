// Error reporting and device queries shared by every entry point of the C ABI
// declared in include/tssep_b200.h.
#include "../../include/tssep_b200.h"
#include "common.cuh"

namespace tssep {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;  // idempotent lookup; benign if raced
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

}  // namespace tssep

extern "C" {

const char* tssep_last_error(void) { return tssep::g_error; }

int tssep_abi_version(void) { return TSSEP_ABI_VERSION; }

int tssep_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  TSSEP_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  TSSEP_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}

}  // extern "C"
