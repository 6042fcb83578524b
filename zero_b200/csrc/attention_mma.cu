// attention_mma.cu — tensor-core fast path for plain softmax attention (placeholder until enabled).
#include "zb_common.h"

namespace zb {
bool attention_mma_supported(const zb_attention_args* a, bool bwd) { return false; }
int attention_mma_fwd(const zb_attention_args* a, cudaStream_t st) { return ZB_EUNSUPPORTED; }
int attention_mma_bwd(const zb_attention_args* a, cudaStream_t st) { return ZB_EUNSUPPORTED; }
}  // namespace zb
