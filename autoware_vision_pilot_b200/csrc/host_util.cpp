// host_util.cpp — error reporting shared by every entry point of libvp_b200.so.
#include <cstdarg>
#include <cstdio>
#include "../../include/vp_b200_ops.h"
#include "ops_internal.h"

static thread_local char g_err[1024] = "";

extern "C" const char* vpb_last_error(void) { return g_err; }

extern "C" void vpb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace vpb {
std::mutex& init_mutex() {
  static std::mutex m;
  return m;
}
bool* device_flag(InitSlot slot) {
  static bool flags[kInitSlots][64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 63;
  return &flags[slot][dev];
}
}  // namespace vpb
