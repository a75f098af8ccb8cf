// Error plumbing and device queries for the C ABI (include/teco.h).
#include <stdarg.h>
#include <string.h>
#include "teco_common.cuh"

static thread_local char g_err[512] = "";

void teco_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int teco_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

extern "C" {

const char* teco_last_error(void) { return g_err; }

int teco_version(void) { return 100; }

int teco_device_props(int device, int64_t* props) {
  TECO_CHECK_ARG(props != nullptr, "teco_device_props: props is NULL");
  cudaDeviceProp p;
  TECO_CUDA_CALL(cudaGetDeviceProperties(&p, device));
  props[0] = p.multiProcessorCount;
  props[1] = p.major;
  props[2] = p.minor;
  props[3] = (int64_t)p.sharedMemPerBlockOptin;
  props[4] = (int64_t)p.l2CacheSize;
  props[5] = props[6] = props[7] = 0;
  return TECO_OK;
}

}  // extern "C"
