// Library-wide state: last-error message, launch counter, device properties.
#include <stdarg.h>
#include <atomic>

#include "common.cuh"

namespace mip360 {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_options[OPT_COUNT] = {{1}, {1}, {1}, {1}, {1}};

bool option(int key) { return key >= 0 && key < OPT_COUNT && g_options[key].load(std::memory_order_relaxed) != 0; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) dev = 0;
  return dev;
}

int sm_count() {
  static std::atomic<int> cached[MAX_DEVICES] = {};
  const int dev = current_device();
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

}  // namespace mip360

extern "C" {
const char* mip360_last_error(void) { return mip360::g_err; }
int mip360_version(void) { return 100; }
long long mip360_launch_count(void) { return mip360::g_launches.load(); }
void mip360_reset_launch_count(void) { mip360::g_launches.store(0); }
int mip360_sm_count(void) { return mip360::sm_count(); }
int mip360_set_option(int key, int value) {
  if (key < 0 || key >= mip360::OPT_COUNT) {
    mip360::set_error("set_option: unknown key %d", key);
    return MIP360_ERR_ARG;
  }
  mip360::g_options[key].store(value != 0);
  return MIP360_OK;
}
}
