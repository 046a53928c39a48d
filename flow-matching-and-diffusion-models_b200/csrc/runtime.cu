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

// -1: FMDM_PDL unset (off unless a caller switches it on for a capture, fm_set_pdl), 0 / 1: forced by the environment
static int pdl_env() {
  static int env = -2;
  if (env == -2) {
    const char* e = getenv("FMDM_PDL");
    env = e == nullptr ? -1 : (e[0] == '1' ? 1 : 0);
  }
  return env;
}
static int g_pdl = 0;

bool pdl_enabled() {
  const int env = pdl_env();
  return env >= 0 ? env == 1 : g_pdl == 1;
}

int set_pdl(int on) {
  const int was = g_pdl;
  g_pdl = on ? 1 : 0;
  return was;
}

}  // namespace fm

extern "C" int fm_version(void) { return 100; }
extern "C" int fm_set_pdl(int on) { return fm::set_pdl(on); }
extern "C" const char* fm_last_error(void) { return fm::g_err; }
extern "C" long long fm_launch_count(void) { return fm::g_launches.load(std::memory_order_relaxed); }
