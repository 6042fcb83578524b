// zb_common.h — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/zero_b200.h"

namespace zb {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return ZB_ECUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ZB_OK;
}

int num_sms();

}  // namespace zb

#define ZB_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      zb::set_error(__VA_ARGS__);    \
      return ZB_EINVAL;              \
    }                                \
  } while (0)
