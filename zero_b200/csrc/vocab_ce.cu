// vocab_ce.cu — K6: the vocabulary projection fused with the label-smoothed cross-entropy.
//
// The reference computes logits = feature @ E^T ([tokens, V] fp32, models/transformer.py:186-196), builds a dense
// smoothed one-hot target of the same size (utils/util.py:88-103) and calls softmax_cross_entropy_with_logits_v2
// (models/transformer.py:198-211): at BASELINE configs[1] that is 524 MB of logits written and re-read, plus as much
// again for the soft labels, every step.  Here the [tokens, V] logits never exist in memory:
//
//   pass 1  the tcgen05 GEMM (gemm2_tcgen05.cu, ce_mode 1) reduces every 128 x 256 accumulator tile, straight out of
//           TMEM, to per-row {max, sum exp(x - max), sum x, x[gold]}: 16 bytes per (row, half tile) instead of 1 KB
//   merge   one warp per row folds the V / 128 partials into the row's log-sum-exp, its smoothed NLL
//             nll = lse - p x[gold] - q (sum x - x[gold]) - normaliser,   p = 1 - eps, q = eps / (V - 1)
//           and its weight d loss / d nll = mask / (len_b * batch) * loss_scale; then the per-sentence / batch means
//   pass 2  (training only) the same GEMM again (ce_mode 2): the epilogue turns the recomputed logits into
//           d_logits = (exp(x - lse) - soft_label) * weight and stores them as bf16 — the only [tokens, V] tensor
//           that is ever written, read once each by the two gradient GEMMs that follow.
// The recomputation costs one more 2 * tokens * V * d GEMM (0.1 ms at configs[1]) and removes 1.05 GB of fp32 logits
// traffic and the stand-alone cross-entropy kernel (0.17 ms).
#include <math.h>

#include "zb_common.h"
#include "zb_ptx.cuh"

namespace zb {

int gemm2_launch_ce(const zb_gemm_args* a, int mode, const float* aux, const int32_t* labels, float4* stats, float ce_p,
                    float ce_q, cudaStream_t st);  // gemm2_tcgen05.cu
int ce_reduce_launch(const float* nll, const int32_t* labels, int batch, int seq_len, float* per_sample, float* loss,
                     cudaStream_t st);  // ce.cu

// one warp per token row
__global__ void __launch_bounds__(256)
vocab_ce_merge_kernel(const float4* __restrict__ stats, int parts, long long rows, const int32_t* __restrict__ labels,
                      int batch, int seq_len, int vocab, float smooth, float loss_scale, float* __restrict__ nll,
                      float2* __restrict__ aux) {
  grid_dep_wait();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float m = -INFINITY, s = 0.f, tot = 0.f, gold_x = 0.f;
  for (int k = lane; k < parts; k += 32) {
    const float4 t = __ldg(stats + (long long)k * rows + row);
    if (t.x > -INFINITY) {   // an empty part (every column past V) holds {-inf, 0, 0, 0}
      const float nm = fmaxf(m, t.x);
      s = s * __expf(m - nm) + t.y * __expf(t.x - nm);
      m = nm;
    }
    tot += t.z;
    gold_x += t.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    if (m2 > -INFINITY) {
      if (m == -INFINITY) {
        m = m2;
        s = s2;
      } else {
        const float nm = fmaxf(m, m2);
        s = s * __expf(m - nm) + s2 * __expf(m2 - nm);
        m = nm;
      }
    }
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
    gold_x += __shfl_xor_sync(0xffffffffu, gold_x, o);
  }
  // tokens of this row's sentence (models/transformer.py:208-210)
  const long long b = row / seq_len;
  int len = 0;
  for (int t = lane; t < seq_len; t += 32) len += labels[b * seq_len + t] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(0xffffffffu, len, o);
  if (lane == 0) {
    const float lse = m + logf(s);
    const float lg = gold_x - lse;
    float val;
    if (smooth > 0.f && smooth < 1.f) {
      const float n = (float)(vocab - 1);
      const float p = 1.f - smooth, q = smooth / n;
      const float norm = -(p * logf(p) + n * q * logf(q + 1e-20f));
      const float sum_lsm = tot - (float)vocab * lse;
      val = -(p * lg + q * (sum_lsm - lg)) - norm;
    } else {
      val = -lg;
    }
    nll[row] = val;
    float w = 0.f;
    if (labels[row] != 0 && len > 0) w = loss_scale / ((float)len * (float)batch);
    if (aux) aux[row] = make_float2(lse, w);
  }
}

static long long vocab_ce_parts(const zb_vocab_ce_args* a) { return 2ll * ((a->vocab + 255) / 256); }

}  // namespace zb

extern "C" int64_t zb_vocab_ce_workspace_bytes(const zb_vocab_ce_args* a) {
  if (!a) return 0;
  const long long rows = (long long)a->batch * a->seq_len;
  return zb::vocab_ce_parts(a) * rows * 16 + rows * 8;
}

extern "C" int zb_vocab_ce(const zb_vocab_ce_args* a, zb_stream_t stream) {
  using namespace zb;
  ZB_REQUIRE(a && a->feat && a->table && a->labels && a->nll && a->workspace, "zb_vocab_ce: null pointer");
  ZB_REQUIRE(a->vocab >= 128 && a->batch >= 0 && a->seq_len > 0 && a->d > 0 && a->d % 8 == 0, "zb_vocab_ce: bad shape");
  ZB_REQUIRE(a->ldf % 8 == 0 && a->ldt % 8 == 0, "zb_vocab_ce: operand pitches must be multiples of 8 elements");
  ZB_REQUIRE(a->workspace_bytes >= zb_vocab_ce_workspace_bytes(a), "zb_vocab_ce: workspace too small (%lld < %lld)",
             (long long)a->workspace_bytes, (long long)zb_vocab_ce_workspace_bytes(a));
  ZB_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0, "zb_vocab_ce: workspace must be 16-byte aligned");
  const long long rows = (long long)a->batch * a->seq_len;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (rows == 0) return ZB_OK;
  const int parts = (int)vocab_ce_parts(a);
  float4* stats = reinterpret_cast<float4*>(a->workspace);
  float2* aux = reinterpret_cast<float2*>(stats + (long long)parts * rows);
  float p = 1.f, q = 0.f;
  if (a->smooth > 0.f && a->smooth < 1.f) {
    p = 1.f - a->smooth;
    q = a->smooth / (float)(a->vocab - 1);
  }
  zb_gemm_args g = {};
  g.a = a->feat; g.b = a->table; g.d = a->d_logits;
  g.m = rows; g.n = a->vocab; g.k = a->d;
  g.lda = a->ldf; g.ldb = a->ldt; g.ldd = a->ldd;
  g.a_layout = ZB_K_MAJOR; g.b_layout = ZB_K_MAJOR; g.d_dtype = ZB_BF16;
  g.alpha = 1.f; g.flags = 0;
  int rc = gemm2_launch_ce(&g, 1, nullptr, a->labels, stats, p, q, st);
  if (rc) return rc;
  const int wpb = 8;
  ZB_LAUNCH(vocab_ce_merge_kernel, (unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st, (const float4*)stats, parts, rows,
            a->labels, (int)a->batch, (int)a->seq_len, (int)a->vocab, a->smooth, a->loss_scale, a->nll, aux);
  rc = check_launch("zb_vocab_ce(merge)");
  if (rc) return rc;
  if (a->per_sample || a->loss) {
    rc = ce_reduce_launch(a->nll, a->labels, (int)a->batch, (int)a->seq_len, a->per_sample, a->loss, st);
    if (rc) return rc;
  }
  if (a->d_logits) rc = gemm2_launch_ce(&g, 2, reinterpret_cast<const float*>(aux), a->labels, nullptr, p, q, st);
  return rc;
}
