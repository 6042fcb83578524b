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
extern std::atomic<int64_t> g_path_launches[ZB_PATH_COUNT_];
inline void note_path(int which) { g_path_launches[which].fetch_add(1, std::memory_order_relaxed); }

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
int num_sms_compute();   // num_sms() minus the SMs reserved for a concurrent collective (zb_set_sm_reserve)

// Every kernel of the library starts with griddepcontrol.wait and is launched with programmatic stream
// serialisation, so its launch latency and prologue overlap the tail of its predecessor in the stream (or in the
// captured CUDA graph).  ZB_NO_PDL=1 turns the attribute off.
bool pdl_enabled();

// Every kernel of the library asks for the SAME shared-memory carveout (the maximum): consecutive kernels with
// different L1 / shared splits make the SM reconfigure between them.  ZB_CARVEOUT=0 leaves the driver's default.
bool carveout_enabled();
void note_kernel_for_carveout(const void* kern);

template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  if (carveout_enabled()) note_kernel_for_carveout(reinterpret_cast<const void*>(kern));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);  // errors surface in check_launch()
}

}  // namespace zb

#define ZB_LAUNCH(kern, grid, block, smem, stream, ...) \
  ::zb::launch_kernel(kern, dim3(grid), dim3(block), static_cast<size_t>(smem), stream, __VA_ARGS__)

#define ZB_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      zb::set_error(__VA_ARGS__);    \
      return ZB_EINVAL;              \
    }                                \
  } while (0)
