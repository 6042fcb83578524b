// vocab_topk.cu — K8 fused: the decode step's vocabulary projection reduced to beam-search candidates.
//
// The reference computes logits = feature @ E^T for every live beam ([batch * beam, V] fp32, models/transformer.py:
// 186-196), then search.py:147-176 takes a log-softmax over V, adds the beam's score and keeps the 2 * beam best of the
// beam * V continuations.  Score = (prev + x / T - lse) / penalty is monotone in the logit x inside a row, so the 2 * beam
// best of the sentence are among the 2 * beam (<= 8) largest logits of each row — and those are among the 8 largest of
// every 128-column piece of the row.  The tcgen05 GEMM (gemm2_tcgen05.cu, ce_mode 3) therefore never stores the
// logits (32.8 MB written and read back per step at BASELINE configs[2]): each epilogue thread reduces its row's 128
// accumulator columns, straight out of TMEM, to {max, sum exp(x - max)} and a sorted top-8 (value, column) list:
// 80 bytes per (row, part) instead of 512.  zb_beam_step (beam.cu, beam_cand_kernel) folds the V / 128 parts of a
// row into its log-sum-exp, scores the 8 * parts candidates exactly as the logits kernels score all V, and carries on
// with the same bookkeeping.
//
// Exactness.  The beam step orders continuations by (score desc, flat index asc) like tf.nn.top_k; the epilogue orders a
// part's columns by (logit desc, column asc).  The two agree wherever scores are strictly monotone in the logit, and for
// equal logits.  They can differ only where two DIFFERENT logits round to the same fp32 score, the larger one in the
// higher column, at the 8th / 9th place of one part — and that changes the result only if the sentence's whole top-2k
// lies in that part of that row.  tests/test_beam_candidates_cpu.py checks the algorithm against the oracle's restatement
// of search.py (first step with the EOS ban, dead beams, temperature, a frequency-sorted vocabulary, exact ties);
// tests/test_kernels_gpu.py checks the kernels against the logits path step for step.
//
// Workspace layout (parts = 2 * ceil(V / 256), rows = batch * beam):
//   float4 stats[parts][rows]      {max, sum exp(x - max), 0, 0}; {-inf, 0, 0, 0} for a part past V
//   float  cval [parts][rows][8]   descending, ties -> lower column; -inf = empty slot
//   int32  cidx [parts][rows][8]   column in [0, V)
#include <math.h>

#include "zb_common.h"

namespace zb {

int gemm2_launch_topk(const zb_gemm_args* a, float4* stats, float* cval, int32_t* cidx, int skip_col, float temperature,
                      cudaStream_t st);  // gemm2_tcgen05.cu

}  // namespace zb

extern "C" int32_t zb_vocab_topk_parts(int32_t vocab) { return 2 * ((vocab + 255) / 256); }

extern "C" int64_t zb_vocab_topk_workspace_bytes(const zb_vocab_topk_args* a) {
  if (!a) return 0;
  return (int64_t)zb_vocab_topk_parts(a->vocab) * a->rows * (16 + 32 + 32);
}

extern "C" int zb_vocab_topk(const zb_vocab_topk_args* a, zb_stream_t stream) {
  using namespace zb;
  ZB_REQUIRE(a && a->feat && a->table && a->workspace, "zb_vocab_topk: null pointer");
  ZB_REQUIRE(a->vocab >= 128 && a->rows >= 0 && a->d > 0 && a->d % 8 == 0, "zb_vocab_topk: bad shape");
  ZB_REQUIRE(a->ldf % 8 == 0 && a->ldt % 8 == 0, "zb_vocab_topk: operand pitches must be multiples of 8 elements");
  ZB_REQUIRE(a->temperature > 0.f, "zb_vocab_topk: temperature must be positive");
  ZB_REQUIRE(a->skip_col >= -1 && a->skip_col < a->vocab, "zb_vocab_topk: skip_col outside the vocabulary");
  ZB_REQUIRE(a->workspace_bytes >= zb_vocab_topk_workspace_bytes(a), "zb_vocab_topk: workspace too small (%lld < %lld)",
             (long long)a->workspace_bytes, (long long)zb_vocab_topk_workspace_bytes(a));
  ZB_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0, "zb_vocab_topk: workspace must be 16-byte aligned");
  if (a->rows == 0) return ZB_OK;
  const long long slots = (long long)zb_vocab_topk_parts(a->vocab) * a->rows;
  float4* stats = reinterpret_cast<float4*>(a->workspace);
  float* cval = reinterpret_cast<float*>(stats + slots);
  int32_t* cidx = reinterpret_cast<int32_t*>(cval + slots * 8);
  zb_gemm_args g = {};
  g.a = a->feat; g.b = a->table; g.d = nullptr;
  g.m = a->rows; g.n = a->vocab; g.k = a->d;
  g.lda = a->ldf; g.ldb = a->ldt; g.ldd = 0;
  g.a_layout = ZB_K_MAJOR; g.b_layout = ZB_K_MAJOR; g.d_dtype = ZB_F32;
  g.alpha = 1.f; g.flags = 0;
  return gemm2_launch_topk(&g, stats, cval, cidx, a->skip_col, a->temperature, reinterpret_cast<cudaStream_t>(stream));
}
