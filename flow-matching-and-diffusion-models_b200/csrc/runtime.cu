// Error plumbing, launch accounting and device checks shared by every entry point.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace fm {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return (int)e;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int g_dev_checked = -2;  // -2 unknown, -1 unusable, >=0 device ordinal validated
static int g_sm_count = 0;

int ensure_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); fmdm_b200 has no CPU path", cudaGetErrorString(e));
    return FM_ERR_NO_DEVICE;
  }
  if (dev == g_dev_checked) return 0;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaGetDeviceProperties failed (%s)", cudaGetErrorString(e));
    return FM_ERR_NO_DEVICE;
  }
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; fmdm_b200 kernels are built for sm_100a only", dev, prop.major, prop.minor);
    return FM_ERR_NO_DEVICE;
  }
  g_sm_count = prop.multiProcessorCount;
  g_dev_checked = dev;
  return 0;
}

int sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("FMDM_PDL");
    on = (e != nullptr && e[0] == '1') ? 1 : 0;  // opt-in: measured neutral-to-negative on the power-capped loops
  }
  return on == 1;
}

}  // namespace fm

extern "C" int fm_version(void) { return 100; }
extern "C" const char* fm_last_error(void) { return fm::g_err; }
extern "C" long long fm_launch_count(void) { return fm::g_launches.load(std::memory_order_relaxed); }
