// Library runtime: error string, launch counter, device attributes.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "i3d_common.cuh"

namespace i3d {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      n = v;
    else
      n = 148;  // B200
  }
  return n;
}

static int g_pdl = -1;   // -1: not decided yet (I3D_PDL env, default on)
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("I3D_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}

}  // namespace i3d

extern "C" {

int i3d_set_pdl(int enabled) {
  const int old = i3d::pdl_enabled() ? 1 : 0;
  i3d::g_pdl = enabled ? 1 : 0;
  return old;
}

int i3d_version(void) { return 100; }

const char* i3d_last_error_string(void) { return i3d::g_err; }

int64_t i3d_launch_count(void) { return i3d::g_launches.load(std::memory_order_relaxed); }
}
