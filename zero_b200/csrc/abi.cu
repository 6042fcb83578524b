// abi.cu — process-wide pieces of the C ABI: version, thread-local error string, launch counter.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_set>

#include "zb_common.h"

namespace zb {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<int64_t> g_path_launches[ZB_PATH_COUNT_] = {};

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

bool carveout_enabled() {
  static const bool on = getenv("ZB_CARVEOUT") != nullptr && getenv("ZB_CARVEOUT")[0] == '1';
  return on;
}

void note_kernel_for_carveout(const void* kern) {
  static std::mutex mu;
  static std::unordered_set<const void*> done;
  std::lock_guard<std::mutex> lock(mu);
  if (done.insert(kern).second)
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
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

// SMs the persistent tensor-core kernels (GEMM, attention) size their grids for.  A data-parallel job runs NCCL's
// all-reduce kernels (one CTA per channel) next to the backward pass; a persistent grid of exactly num_sms() CTAs
// then cannot be co-resident and runs in two rounds (measured: the overlapped GEMMs take up to 2x).  Leaving a few
// SMs to the collective (zb_set_sm_reserve, set by the trainer when world_size > 1) costs reserve / num_sms instead.
static std::atomic<int> g_sm_reserve{0};
int num_sms_compute() {
  const int n = num_sms() - g_sm_reserve.load(std::memory_order_relaxed);
  return n < 8 ? 8 : (n & ~1);   // even: the CTA-pair GEMM launches whole pairs
}

}  // namespace zb

extern "C" int zb_set_sm_reserve(int32_t n) {
  if (n < 0 || n > 64) {
    zb::set_error("zb_set_sm_reserve: %d outside [0, 64]", (int)n);
    return ZB_EINVAL;
  }
  zb::g_sm_reserve.store(n, std::memory_order_relaxed);
  return ZB_OK;
}

extern "C" int zb_abi_version(void) { return ZB_ABI_VERSION; }
extern "C" const char* zb_last_error_string(void) { return zb::g_err; }
extern "C" int64_t zb_launch_count(void) { return zb::g_launches.load(); }
extern "C" int64_t zb_path_launch_count(int32_t which) {
  return which >= 0 && which < ZB_PATH_COUNT_ ? zb::g_path_launches[which].load() : -1;
}
extern "C" int64_t zb_abi_struct_size(int32_t which) {
  switch (which) {
    case 0: return sizeof(zb_gemm_args);
    case 1: return sizeof(zb_attention_args);
    case 2: return sizeof(zb_add_ln_args);
    case 3: return sizeof(zb_embed_args);
    case 4: return sizeof(zb_ce_args);
    case 5: return sizeof(zb_adam_args);
    case 6: return sizeof(zb_beam_args);
    case 7: return sizeof(zb_colsum_args);
    case 8: return sizeof(zb_shard_adam_args);
    case 9: return sizeof(zb_vocab_ce_args);
    case 10: return sizeof(zb_vocab_topk_args);
    default: return -1;
  }
}
