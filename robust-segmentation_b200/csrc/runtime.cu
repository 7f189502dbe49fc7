// Library-level plumbing: ABI version, thread-local error text, cached device properties.
#include <cstring>

#include "common.cuh"

namespace robseg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static thread_local cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;
cudaEvent_t take_profile_start() {
  cudaEvent_t e = g_prof_start;
  g_prof_start = nullptr;
  return e;
}
cudaEvent_t take_profile_stop() {
  cudaEvent_t e = g_prof_stop;
  g_prof_stop = nullptr;
  return e;
}

}  // namespace robseg

extern "C" int robseg_profile_next_kernel(void* start_event, void* stop_event) {
  robseg::g_prof_start = static_cast<cudaEvent_t>(start_event);
  robseg::g_prof_stop = static_cast<cudaEvent_t>(stop_event);
  return 0;
}

extern "C" int robseg_version(void) { return ROBSEG_ABI_VERSION; }

extern "C" const char* robseg_last_error(void) { return robseg::g_err; }
