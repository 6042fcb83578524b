// attention.cu — C-ABI entry points of K2/K3: validation + dispatch between the tensor-core fast path
// (attention_mma.cu: plain softmax attention, dh = 64) and the feature-complete path (attention_generic.cu).
#include "zb_common.h"

namespace zb {
int attention_generic_fwd(const zb_attention_args* a, cudaStream_t st);
int attention_generic_bwd(const zb_attention_args* a, cudaStream_t st);
bool attention_decode_supported(const zb_attention_args* a);
int attention_decode_fwd(const zb_attention_args* a, cudaStream_t st);
bool attention_mma_supported(const zb_attention_args* a, bool bwd);
int attention_mma_fwd(const zb_attention_args* a, cudaStream_t st);
int attention_mma_bwd(const zb_attention_args* a, cudaStream_t st);
bool attention_tc_supported(const zb_attention_args* a, bool bwd);   // attention_tc.cu (tcgen05 / TMEM / TMA)
int attention_tc_fwd(const zb_attention_args* a, cudaStream_t st);
int attention_tc_bwd(const zb_attention_args* a, cudaStream_t st);
long long attention_tc_bwd_workspace_bytes(const zb_attention_args* a);

static int validate(const zb_attention_args* a, bool bwd) {
  ZB_REQUIRE(a && a->q && a->k && a->v && a->o, "zb_attention: null pointer");
  ZB_REQUIRE(a->batch >= 0 && a->heads > 0 && a->lq > 0 && a->lk > 0 && a->dh > 0, "zb_attention: bad shape");
  ZB_REQUIRE((a->rpr_k == nullptr) == (a->rpr_v == nullptr), "zb_attention: rpr_k and rpr_v go together");
  ZB_REQUIRE(!a->rpr_k || (a->max_rel > 0 && a->max_rel <= 64), "zb_attention: max_rel out of range");
  ZB_REQUIRE(!(a->relu_attn && a->rpr_k), "zb_attention: ReLA + rpr is not a reference configuration");
  if (bwd) {
    ZB_REQUIRE(a->kv_group <= 1, "zb_attention_bwd: kv_group is a decode-only feature");
    ZB_REQUIRE(a->d_o && a->dq && a->dk && a->dv && a->lse && a->delta, "zb_attention_bwd: null pointer");
    ZB_REQUIRE(!a->rpr_k || (a->d_rpr_k && a->d_rpr_v), "zb_attention_bwd: rpr gradients missing");
  }
  return ZB_OK;
}
}  // namespace zb

extern "C" int zb_attention_fwd(const zb_attention_args* a, zb_stream_t stream) {
  using namespace zb;
  int rc = validate(a, false);
  if (rc) return rc;
  if (a->batch == 0) return ZB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (attention_decode_supported(a)) return attention_decode_fwd(a, st);
  if (attention_tc_supported(a, false)) return attention_tc_fwd(a, st);
  if (attention_mma_supported(a, false)) return attention_mma_fwd(a, st);
  return attention_generic_fwd(a, st);
}

extern "C" int zb_attention_bwd(const zb_attention_args* a, zb_stream_t stream) {
  using namespace zb;
  int rc = validate(a, true);
  if (rc) return rc;
  if (a->batch == 0) return ZB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (attention_tc_supported(a, true)) return attention_tc_bwd(a, st);
  if (attention_mma_supported(a, true)) return attention_mma_bwd(a, st);
  return attention_generic_bwd(a, st);
}

extern "C" int64_t zb_attention_bwd_workspace_bytes(const zb_attention_args* a) {
  return zb::attention_tc_bwd_workspace_bytes(a);
}
