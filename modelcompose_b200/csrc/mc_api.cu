// Library-wide pieces of the C ABI: version, thread-local error text, device attribute cache.
#include <cstring>
#include <mutex>

#include "mc_common.cuh"

namespace mc {

std::string& last_error() {
  static thread_local std::string s;
  return s;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

int& launch_mode() {
  static thread_local int mode = 0;
  return mode;
}

int sm_count() {
  static std::mutex mu;
  static int cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  std::lock_guard<std::mutex> lock(mu);
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cache[dev] = n;
  }
  return cache[dev];
}

}  // namespace mc

extern "C" int mc_abi_version(void) { return MC_ABI_VERSION; }

extern "C" const char* mc_last_error(void) { return mc::last_error().c_str(); }

extern "C" int mc_set_launch_mode(int flags) {
  const int old = mc::launch_mode();
  mc::launch_mode() = flags;
  return old;
}

extern "C" int mc_device_info(char* name, size_t cap, int* sms, int* cc) {
  int dev = 0;
  MC_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MC_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (name && cap) {
    strncpy(name, prop.name, cap - 1);
    name[cap - 1] = 0;
  }
  if (sms) *sms = prop.multiProcessorCount;
  if (cc) *cc = prop.major * 10 + prop.minor;
  return MC_OK;
}
