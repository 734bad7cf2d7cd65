// Library-level state: error text, architecture gate, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace ragb {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int check_device(int dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceProperties(%d) failed: %s", dev, cudaGetErrorString(e));
    return RAGB_ECUDA;
  }
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libragb200 only runs on sm_100 (B200) and has no fallback", dev, prop.major,
              prop.minor);
    return RAGB_EARCH;
  }
  return RAGB_OK;
}

int require_b200() {
  static thread_local int cached_dev = -1;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice failed: %s (no CUDA device? libragb200 has no CPU fallback)", cudaGetErrorString(e));
    return RAGB_ECUDA;
  }
  if (dev == cached_dev) return RAGB_OK;
  int rc = check_device(dev);
  if (rc == RAGB_OK) cached_dev = dev;
  return rc;
}

int device_sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached;
}

}  // namespace ragb

extern "C" {

int ragb_abi_version(void) { return 1; }
const char* ragb_last_error(void) { return ragb::g_error; }
int64_t ragb_launch_count(void) { return ragb::g_launches.load(); }

int ragb_device_check(int device) { return ragb::check_device(device); }

}  // extern "C"
