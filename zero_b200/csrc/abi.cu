// abi.cu — process-wide pieces of the C ABI: version, thread-local error string, launch counter.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "zb_common.h"

namespace zb {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = getenv("ZB_NO_PDL") == nullptr;
  return on;
}

int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

}  // namespace zb

extern "C" int zb_abi_version(void) { return ZB_ABI_VERSION; }
extern "C" const char* zb_last_error_string(void) { return zb::g_err; }
extern "C" int64_t zb_launch_count(void) { return zb::g_launches.load(); }
