/* zero_b200.h — C ABI of libzero_b200.so: the sm_100a kernels behind Zero's Transformer hot path.
 *
 * The reference (bzhangGo/zero @ d97e2c2) has no native code and no FFI: its hot path is TF1.x graph ops
 * built by func.py / models/transformer*.py / search.py.  Each entry point below therefore cites the
 * Python call site(s) it replaces; INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates; this library never allocates,
 *     never synchronises, never throws);
 *   - every call enqueues work on `stream` and returns 0 (ZB_OK) or a negative zb_status;
 *     zb_last_error_string() gives a thread-local description;
 *   - activations / compute weights are bf16 row-major, accumulation / statistics / gradients-of-parameters
 *     are fp32 (mirrors the reference's mixed-precision contract, utils/dtype.py:55-69, with bf16 for fp16);
 *   - token ids are int32 with pad = 0, unk = 1, eos = 2 (vocab.py:20-22).
 */
#ifndef ZERO_B200_H_
#define ZERO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* zb_stream_t; /* == cudaStream_t */

#define ZB_ABI_VERSION 4

typedef enum {
  ZB_OK = 0,
  ZB_EINVAL = -1,       /* bad shape / alignment / null pointer */
  ZB_EUNSUPPORTED = -2, /* dtype / size / device not supported */
  ZB_ECUDA = -3         /* CUDA runtime or driver error, see zb_last_error_string() */
} zb_status;

typedef enum { ZB_BF16 = 0, ZB_F32 = 1 } zb_dtype;

int zb_abi_version(void);
/* SMs left free by the persistent tensor-core kernels (GEMM, attention) for a collective that runs next to them
 * (NCCL all-reduce kernels during the backward pass); 0 (default) on a single GPU.  Process-wide. */
int zb_set_sm_reserve(int32_t n);
const char* zb_last_error_string(void);
/* Number of kernels this library has launched in this process (bench.py's `gpu_launches`). */
int64_t zb_launch_count(void);
/* Launches per dispatch path, for callers (and parity tests) that need to know WHICH kernel served a call:
 * the entry points choose between several implementations of the same op by shape / alignment / switches. */
typedef enum {
  ZB_PATH_GEMM_TCGEN05 = 0,   /* single-CTA tcgen05 GEMM (gemm_tcgen05.cu) */
  ZB_PATH_GEMM_PAIR = 1,      /* CTA-pair tcgen05 GEMM (gemm2_tcgen05.cu), grouped launches count once */
  ZB_PATH_GEMM_SKINNY = 2,    /* <= 384-row mma.sync GEMM (gemm_skinny.cu), opt-in */
  ZB_PATH_ATTN_MMA = 3,       /* attention_mma.cu */
  ZB_PATH_ATTN_GENERIC = 4,   /* attention_generic.cu tiled kernels */
  ZB_PATH_ATTN_DECODE = 5,    /* attention_generic.cu lq = 1 kernel */
  ZB_PATH_BEAM_SENTENCE = 6,  /* beam.cu one CTA per sentence */
  ZB_PATH_BEAM_ROWS = 7,      /* beam.cu one CTA per (sentence, beam) row */
  ZB_PATH_BEAM_PARTS = 8,     /* beam.cu a 4-CTA cluster per row with a threshold pass, opt-in */
  ZB_PATH_GEMM_BM64 = 9,      /* single-CTA tcgen05 GEMM with 64-row tiles, opt-in (also counted as GEMM_TCGEN05) */
  ZB_PATH_ATTN_TC = 10,       /* attention_tc.cu: tcgen05 / TMEM / TMA attention forward and backward (dh = 64) */
  ZB_PATH_BEAM_CAND = 11,     /* beam.cu candidate kernel after zb_vocab_topk (no logits) */
  ZB_PATH_COUNT_ = 12
} zb_path;
int64_t zb_path_launch_count(int32_t which);
/* sizeof() of the argument records, by index: 0 gemm, 1 attention, 2 add_ln, 3 embed, 4 ce, 5 adam, 6 beam,
 * 7 colsum, 8 shard_adam, 9 vocab_ce, 10 vocab_topk; -1 for an unknown index.  Lets a binding check its mirrored struct layouts at load time. */
int64_t zb_abi_struct_size(int32_t which);

/* ------------------------------------------------------------------------------------------------ K1
 * zb_gemm: D[m,n] (+)= alpha * sum_k A(m,k) * B(n,k)  (+ bias[n]) (relu) (* (mask[m,n] > 0))
 * bf16 operands, fp32 accumulation in TMEM (tcgen05.mma), TMA-fed 128B-swizzled smem ring.
 * Replaces tf.matmul in func.linear (func.py:49, bias_add :59), the tied-softmax projection
 * (models/transformer.py:194) and — as dgrad / wgrad — the tf.gradients of both (main.py:28).
 *   a_layout ZB_K_MAJOR : A stored [m][k] (k contiguous), lda = row pitch in elements
 *            ZB_MN_MAJOR: A stored [k][m] (m contiguous)
 *   b_layout ZB_K_MAJOR : B stored [n][k];  ZB_MN_MAJOR: B stored [k][n]
 *   y = x W      : A = x  K-major,  B = W  MN-major        (W is [in,out], func.py:48)
 *   dx = dy W^T  : A = dy K-major,  B = W  K-major
 *   dW = x^T dy  : A = x  MN-major, B = dy MN-major, ZB_EPI_ACCUM into the fp32 gradient arena
 */
typedef enum { ZB_K_MAJOR = 0, ZB_MN_MAJOR = 1 } zb_layout;
enum {
  ZB_EPI_BIAS = 1,      /* + bias[n] (fp32)                                   func.py:59           */
  ZB_EPI_RELU = 2,      /* max(.,0)                                            func.py:332          */
  ZB_EPI_ACCUM = 4,     /* D is fp32 and is atomically accumulated into (D += ...); needed by split-K */
  ZB_EPI_RELU_MASK = 8  /* multiply by (mask[m,n] > 0): backward of relu through the saved activation */
};
typedef struct {
  const void* a;
  const void* b;
  void* d;
  int64_t m, n, k;
  int64_t lda, ldb, ldd; /* pitches in elements */
  int32_t a_layout, b_layout;
  int32_t d_dtype; /* ZB_BF16 or ZB_F32 */
  int32_t flags;
  const float* bias;   /* [n] fp32 or NULL */
  const void* mask;    /* bf16 [m, ldmask] or NULL */
  int64_t ldmask;
  float alpha;
  int32_t split_k;     /* 0 = choose; >1 requires ZB_EPI_ACCUM */
} zb_gemm_args;
int zb_gemm(const zb_gemm_args* a, zb_stream_t stream);
/* zb_gemm_grouped: `count` independent zb_gemm problems (non-overlapping outputs), same results as `count` zb_gemm
 * calls.  The weight gradients of one layer (dW += x^T dy for every func.linear of the layer, i.e. what
 * optimizer.compute_gradients emits per variable, main.py:33-37) are all MN-major / accumulate-into-fp32 and share
 * ONE persistent launch; anything else falls back to per-problem launches. */
int zb_gemm_grouped(const zb_gemm_args* problems, int32_t count, zb_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K2/K3
 * zb_attention_{fwd,bwd}: fused scaled-dot attention softmax(q k^T * scale + mask) v per (batch, head).
 * Replaces func.dot_attention's split_heads/matmul/mask-add/softmax/matmul/combine_heads
 * (func.py:218-256) and their gradients.  q/k/v/o are [batch, len, heads*dh] bf16 views with explicit
 * row pitches (so the fused [B,L,3d] qkv buffer of func.py:196-197 is consumed in place).
 * Masking follows func.attention_bias (func.py:372-388): additive -inf_value where key j is padding
 * (key_len[b] <= j, "masking") and/or j > i + causal_offset ("causal").
 * `lse` [batch, heads, lq] fp32 = log-sum-exp of the masked logits, saved for the backward.
 * Optional relative-position terms (modules/rpr.py:10-75): rpr_k / rpr_v tables [2*max_rel+1, dh] bf16;
 * distance = clip(i + q_offset - j, -max_rel, max_rel) + max_rel.
 */
typedef struct {
  const void* q; const void* k; const void* v; void* o;
  int64_t ldq, ldk, ldv, ldo;         /* row pitch (elements) between consecutive positions */
  int64_t bsq, bsk, bsv, bso;         /* batch stride (elements) */
  int32_t batch, heads, lq, lk, dh;
  const int32_t* key_len;             /* [batch] number of valid keys, or NULL = all valid */
  int32_t causal;                     /* 1: key j visible to query i iff j <= i + q_offset */
  int32_t q_offset;                   /* absolute position of query row 0 (cached decode: time) */
  float scale;                        /* dh^-0.5, func.py:222 */
  float inf_value;                    /* dtype.inf(), default 1e8 (utils/dtype.py:14) */
  float* lse;                         /* [batch, heads, lq] fp32 (fwd: out; bwd: in) */
  const void* rpr_k; const void* rpr_v; int32_t max_rel; /* NULL/0 when unused */
  int32_t relu_attn;                  /* 1: ReLA — relu(logits * keep) instead of softmax (modules/rela.py:65-70) */
  /* backward only */
  const void* d_o; void* dq; void* dk; void* dv;
  int64_t lddo, lddq, lddk, lddv, bsdo, bsdq, bsdk, bsdv;
  float* d_rpr_k; float* d_rpr_v;     /* fp32 [2*max_rel+1, dh], accumulated */
  float* delta;                       /* fp32 [batch, heads, lq] workspace: rowsum(dO * O) */
  int32_t kv_group;                   /* >= 1: query batch b reads keys/values of batch b / kv_group (beams of one
                                         sentence share the projected memory instead of tiling it, search.py:36-39) */
  /* attention dropout (func.py:245, modules/rela.py:74): the softmax / relu weights are multiplied by
     mask / (1 - rate) before the value product; mask = zb_dropout's function of (*seed, site, flat index into
     [batch, heads, lq, lk]).  rate = 0 or seed = NULL: off. */
  float dropout_rate;
  uint32_t dropout_site;
  const uint64_t* dropout_seed;       /* device pointer */
  /* caller-owned scratch of zb_attention_bwd (device, 16-byte aligned, contents undefined on return); may be NULL /
     0, which only narrows the choice of kernels: the tcgen05 backward reduces dq of sequences longer than one
     128-key block in fp32 there.  Size: zb_attention_bwd_workspace_bytes. */
  void* workspace;
  int64_t workspace_bytes;
} zb_attention_args;
int zb_attention_fwd(const zb_attention_args* a, zb_stream_t stream);
int zb_attention_bwd(const zb_attention_args* a, zb_stream_t stream);
/* bytes of `workspace` that let zb_attention_bwd use every kernel for this problem (0: none needed) */
int64_t zb_attention_bwd_workspace_bytes(const zb_attention_args* a);

/* ------------------------------------------------------------------------------------------------ K4
 * zb_add_ln_{fwd,bwd}: out = scale * (s - mean(s)) * rsqrt(var(s) + eps) + offset with s = x (+ y).
 * Replaces func.residual_fn + func.layer_norm (func.py:321-324, 289-303); biased variance, eps = 1e-8.
 * fwd saves mean / rstd (fp32 [rows]).  bwd takes the upstream gradient as d_out (+ d_out2: the gradient
 * arriving through the next residual skip), returns ds (bf16, the gradient wrt both x and y) and
 * accumulates dscale / doffset (fp32 [cols]).
 */
typedef struct {
  const void* x; const void* y; /* y may be NULL */
  void* out; float* mean; float* rstd;
  const float* scale; const float* offset;
  int64_t rows, cols;
  float eps;
  /* backward */
  const void* d_out; const void* d_out2; /* d_out2 (optional) is added to d_out on read */
  void* ds; float* dscale; float* doffset;
  float* dbias; /* optional fp32 [cols], accumulated: column sums of ds = the bias gradient of the linear layer
                   that produced y (tf.nn.bias_add grad, func.py:59) */
  /* forward only, optional: the branch output as an fp32 split-K accumulator [rows, cols] (+ its bias [cols], added
     here because a split-K GEMM cannot): s = x + y32 + ybias (y must be NULL).  The kernel writes zeros back into y32,
     so the next projection of the decode step accumulates into a clean buffer without a memset. */
  float* y32; const float* ybias;
} zb_add_ln_args;
int zb_add_ln_fwd(const zb_add_ln_args* a, zb_stream_t stream);
int zb_add_ln_bwd(const zb_add_ln_args* a, zb_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K5
 * zb_embed_{fwd,bwd}: out[b,l,:] = table[ids[b,l-shift]] * mult + bias + timing(l or time)
 * Replaces tf.gather * sqrt(d) + bias_add + add_timing_signal (models/transformer.py:29-31,104-117,
 * func.py:341-369).  shift = 1 reproduces the decoder's right shift with a zero first row
 * (models/transformer.py:108-111; the zero row still receives the timing signal).
 * zero_if_all_pad = 1 reproduces the decode-time rule (models/transformer.py:113-115): if every id is pad
 * the gathered+biased input is replaced by zeros before the timing signal.
 * bwd scatter-adds d_out * mult into d_table (fp32) and sums d_out into d_bias (fp32).
 */
typedef struct {
  const int32_t* ids; /* [batch, len] */
  const void* table;  /* bf16 [vocab, dim] */
  const float* bias;  /* fp32 [dim] */
  void* out;          /* bf16 [batch, len, dim] */
  int32_t batch, len, dim, vocab;
  int32_t shift; int32_t zero_if_all_pad;
  int32_t time;       /* >= 0: every row uses position `time` (cached decode); -1: position = l */
  float mult;
  const void* d_out; float* d_table; float* d_bias;
  const void* d_out2; /* optional second addend of the upstream gradient */
} zb_embed_args;
int zb_embed_fwd(const zb_embed_args* a, zb_stream_t stream);
int zb_embed_bwd(const zb_embed_args* a, zb_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K6
 * zb_softmax_ce: label-smoothed cross-entropy over fp32 logits [rows, vocab]
 * (util.label_smooth utils/util.py:88-103 + softmax_cross_entropy_with_logits_v2 + normaliser,
 *  models/transformer.py:198-205).  Writes per-token nll (already minus the normalising constant) and,
 * if d_logits != NULL, d_logits = (softmax - soft_label) * row_weight[row] in bf16 (in place allowed
 * only when d_logits != logits).
 */
typedef struct {
  const float* logits; int64_t ld;
  const int32_t* labels;      /* [batch, seq_len] target ids; rows = batch * seq_len */
  int32_t batch, seq_len;
  float* nll;                 /* [rows] per-token smoothed CE minus the normaliser */
  void* d_logits; int64_t ldd;/* bf16 [rows, ldd] or NULL.  Scaled by dLoss/dnll = (label != 0) / (len_b * batch) * loss_scale */
  int32_t vocab;
  float smooth;
  float loss_scale;
  float* per_sample;          /* [batch] sum_t nll * mask / sum_t mask       (models/transformer.py:209) */
  float* loss;                /* [1]     mean_b per_sample                   (models/transformer.py:210) */
} zb_ce_args;
int zb_softmax_ce(const zb_ce_args* a, zb_stream_t stream);

/* zb_vocab_ce: K6 fused — logits = feat @ table^T are reduced to the loss inside the GEMM epilogue and never
 * written ([rows, vocab] fp32 = 524 MB at BASELINE configs[1], models/transformer.py:186-211); with d_logits != NULL the
 * GEMM is run a second time and its epilogue emits d_logits = (softmax - soft_label) * row_weight in bf16.
 * Same outputs as zb_gemm + zb_softmax_ce.  vocab >= 128; d, ldf, ldt, ldd multiples of 8. */
typedef struct {
  const void* feat; int64_t ldf;     /* bf16 [rows, d], rows = batch * seq_len */
  const void* table; int64_t ldt;    /* bf16 [vocab, d]: the (tied) softmax embedding */
  const int32_t* labels;             /* [batch, seq_len] target ids */
  int32_t batch, seq_len, d, vocab;
  float smooth, loss_scale;
  float* nll;                        /* [rows] */
  float* per_sample; float* loss;    /* [batch], [1]; optional */
  void* d_logits; int64_t ldd;       /* bf16 [rows, ldd] or NULL (forward only) */
  void* workspace; int64_t workspace_bytes; /* caller-owned scratch, zb_vocab_ce_workspace_bytes, 16-byte aligned */
} zb_vocab_ce_args;
int zb_vocab_ce(const zb_vocab_ce_args* a, zb_stream_t stream);
int64_t zb_vocab_ce_workspace_bytes(const zb_vocab_ce_args* a);

/* zb_vocab_topk: K8 fused — the decode step's logits = feat @ table^T / temperature (models/transformer.py:186-196,
 * search.py:147) are reduced inside the GEMM epilogue to what the beam step needs and never written: per row and
 * 128-column part {max, sum exp(x - max)} and the part's 8 largest logits with their columns (ties -> lower column,
 * like tf.nn.top_k).  The workspace is handed to zb_beam_step as zb_beam_args.cand (layout: csrc/vocab_topk.cu).
 * skip_col: a column that is never offered as a candidate — EOS at time 0 (search.py:152-155 pushes it below every
 * other word of the row; with vocab - 1 >= 2 * beam it cannot be chosen) — or -1.  vocab >= 128; d, ldf, ldt
 * multiples of 8. */
typedef struct {
  const void* feat; int64_t ldf;     /* bf16 [rows, d], rows = batch * beam */
  const void* table; int64_t ldt;    /* bf16 [vocab, d] */
  int32_t rows, d, vocab;
  int32_t skip_col;
  float temperature;                 /* beam_search_temperature (search.py:147) */
  void* workspace; int64_t workspace_bytes; /* caller-owned, zb_vocab_topk_workspace_bytes, 16-byte aligned */
} zb_vocab_topk_args;
int zb_vocab_topk(const zb_vocab_topk_args* a, zb_stream_t stream);
int64_t zb_vocab_topk_workspace_bytes(const zb_vocab_topk_args* a);
int32_t zb_vocab_topk_parts(int32_t vocab);

/* ------------------------------------------------------------------------------------------------ misc
 * zb_colsum: out[n] += sum_m x[m,n] (bias gradients; tf.nn.bias_add grad).  x bf16, out fp32. */
int zb_colsum(const void* x, int64_t m, int64_t n, int64_t ld, float* out, zb_stream_t stream);
/* zb_colsum_grouped: up to 8 column sums (the bias gradients of one layer) in one launch. */
typedef struct {
  const void* x;       /* bf16 [m, ld] */
  int64_t m, n, ld;
  float* out;          /* [n] fp32, accumulated into */
} zb_colsum_args;
int zb_colsum_grouped(const zb_colsum_args* problems, int32_t count, zb_stream_t stream);
/* zb_cast_f32_bf16: bf16 compute copies of the fp32 master weights (utils/dtype.py:55-69). */
int zb_cast_f32_bf16(const float* src, void* dst, int64_t n, zb_stream_t stream);
int zb_cast_bf16_f32(const void* src, float* dst, int64_t n, zb_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K9
 * zb_adam_tf: TensorFlow-semantics Adam on a flat fp32 arena (tf.train.AdamOptimizer, main.py:178-181):
 *   g = grad * grad_scale;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *   p -= lr * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps);  also refreshes the bf16 compute copy.
 * grad_scale carries 1/world_size, 1/loss_scale and the clip_by_global_norm factor (utils/cycle.py:94-101).
 * lr_t is already bias-corrected by the host; clip_scale (device fp32[1], optional) multiplies grad_scale so
 * the clip factor can be computed on the device without a host sync.
 */
typedef struct {
  float* param; float* m; float* v; const float* grad; void* param_bf16;
  int64_t n;
  float beta1, beta2, eps;
  float lr_t, grad_scale;
  const float* clip_scale;
  float* norms; /* optional fp32[2], accumulated: {sum (grad*grad_scale)^2, sum param^2} (utils/cycle.py:94-95) */
} zb_adam_args;
int zb_adam_tf(const zb_adam_args* a, zb_stream_t stream);
/* zb_sumsq: out[0] += sum x^2 (tf.global_norm, utils/cycle.py:94). */
int zb_sumsq(const float* x, int64_t n, float* out, zb_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K10
 * zb_shard_adam: gradient aggregation + optimizer step of ONE rank's shard of the flat arenas in ONE kernel, over
 * NVLink / NVSwitch peer memory (opt-in; replaces utils/parallel.py:134-208 average_gradients + main.py:42-43 +
 * the TF Adam of main.py:178-181 when the batch is sharded over the GPUs of one box):
 *     g[i]  = sum over ranks of grad_r[i]           i in [lo, lo + n)   (reduce-scatter step)
 *     Adam on param / m / v [lo, lo + n)            (zb_adam_tf semantics, this rank owns the shard)
 *     mirror_r[i] = bf16(param[i]) on EVERY rank    (all-gather step, bf16: half the bytes of the fp32 gradients)
 * The sum is taken either inside the NVSwitch (grad_mc != NULL: the multicast address of the symmetric gradient
 * arena, multimem.ld_reduce — the reduced element arrives over this GPU's link once) or in registers from the ranks' unicast
 * addresses (grad_peer[r], rank order).  The bf16 copy goes out through multimem.st (mirror_mc) or one store per rank.
 * Every rank runs the call on its own shard; the shards tile the arena.  The caller brackets the call with a
 * cross-rank barrier on both sides (all gradients final before / all copies landed after); the kernel itself only
 * ends with a system-scope fence.  Optimizer state outside the shard is not touched (and goes stale on this rank),
 * except where the forward pass reads the fp32 master directly (biases, LayerNorm scale / offset — every 1-D
 * variable): wide_mask (optional, one byte per 64-element slot of the arena, non-zero = broadcast) marks the slots
 * whose refreshed fp32 values are ALSO stored into every rank's master arena (param_mc / param_peer[r]).
 * flags: ZB_SHARD_UPDATE       Adam + bf16 broadcast (without it nothing but the outputs below is written)
 *        ZB_SHARD_STORE_GRAD   grad_out[i] = g[i] (local fp32 arena; the two-pass clip_by_global_norm flow)
 *        ZB_SHARD_NORM_G / _P  norms[0] += sum (g * grad_scale)^2 / norms[1] += sum param^2 (pre-update)
 * norm_parts_peer (optional): fp32 [world][2] tables, one per rank; the last CTA to finish copies this rank's
 * norms[0..1] into row `rank` of every table, so after the closing barrier each rank can form tf.global_norm
 * (utils/cycle.py:94-95) without another collective.  done_counter: device uint32, zero on entry, zero on exit.
 * lo and n are multiples of 8 elements. */
#define ZB_SHARD_MAX_WORLD 16
enum { ZB_SHARD_UPDATE = 1, ZB_SHARD_STORE_GRAD = 2, ZB_SHARD_NORM_G = 4, ZB_SHARD_NORM_P = 8 };
typedef struct {
  int64_t lo, n;
  int32_t world, rank;   /* ranks the bf16 / fp32 copies go to; this rank's index (row of the norm tables) */
  int32_t grad_sources;  /* gradient copies summed from grad_peer[0 .. grad_sources): `world` for the reduce step,
                            1 when grad_peer[0] already holds the summed gradients (second pass of the clip flow) */
  int32_t flags;
  const float* grad_mc;
  const float* grad_peer[ZB_SHARD_MAX_WORLD];
  float* param; float* m; float* v; /* arena bases (element 0), local */
  const uint8_t* wide_mask;
  float* param_mc;
  float* param_peer[ZB_SHARD_MAX_WORLD];
  void* mirror_mc;
  void* mirror_peer[ZB_SHARD_MAX_WORLD];
  float* grad_out;
  float beta1, beta2, eps;
  float lr_t, grad_scale;
  const float* clip_scale;
  float* norms;
  float* norm_parts_peer[ZB_SHARD_MAX_WORLD];
  uint32_t* done_counter;
} zb_shard_adam_args;
int zb_shard_adam(const zb_shard_adam_args* a, zb_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K8
 * zb_beam_step: one expansion step of search.beam_search (search.py:115-238) for a whole batch.
 * See DESIGN.md for the state layout.  All scores fp32; indices int32; ties resolved to the lower
 * flat index like tf.nn.top_k.
 */
typedef struct {
  const float* logits;   /* [batch*beam, vocab] fp32; NULL when `cand` is given */
  int32_t batch, beam, vocab;
  int32_t time;          /* 0-based step */
  int32_t eos_id, pad_id;
  float temperature;     /* beam_search_temperature (search.py:147) */
  float inf_value;       /* dtype.inf(): forbids EOS at t = 0 (search.py:152-155) */
  float length_penalty;  /* ((5 + time + 1) / 6) ^ decode_alpha, computed by the host in fp32 (search.py:169) */
  const int32_t* max_len;/* [batch] int(src_len + decode_length) (search.py:33,195) */
  const float* max_penalty; /* [batch] ((5 + max_len) / 6) ^ alpha (search.py:95-96) */
  int32_t seq_cap;       /* allocated length of every sequence row; needs time + 2 <= seq_cap */
  /* alive state, in/out */
  int32_t* alive_seq;    /* [batch, beam, seq_cap]; positions [0, time] valid on entry ([0] = pad/BOS) */
  float* alive_logp;     /* [batch, beam] */
  float* alive_score;    /* [batch, beam] */
  /* finished state, in/out */
  int32_t* fin_seq;      /* [batch, beam, seq_cap] */
  float* fin_score;      /* [batch, beam] */
  int32_t* fin_flag;     /* [batch, beam] 0/1 */
  /* out: flat row (b * beam + previous beam) each new alive beam continues, for state reordering */
  int32_t* parent;       /* [batch * beam] */
  /* scratch */
  int32_t* tmp_seq;      /* [batch, 3 * beam, seq_cap] */
  /* loop control: zb_beam_cond writes active[0] = search.py:85-113's _not_finished(time); zb_beam_step is a
   * no-op when active[0] == 0, so the host may poll the flag every few steps without changing the result */
  int32_t* active;       /* [1] device */
  /* optional scratch of the row-parallel kernel (one CTA per (sentence, beam) row): batch * (4 * beam^2 + 1) 4-byte
   * words, zeroed ONCE by the caller (the kernel leaves it zeroed); NULL = one CTA per sentence */
  float* row_ws;
  /* instead of `logits`: the workspace zb_vocab_topk wrote for these batch * beam rows with this step's temperature and
   * (time == 0) skip_col = eos_id; needs row_ws, beam <= 4 and vocab - 1 >= 2 * beam.  `temperature` is not applied
   * again.  Same result as the logits path (the candidates are a superset of every row's 2 * beam best). */
  const void* cand;
} zb_beam_args;
int zb_beam_cond(const zb_beam_args* a, zb_stream_t stream);
int zb_beam_step(const zb_beam_args* a, zb_stream_t stream);

/* zb_gather_rows: dst[r, :row_bytes] = src[index[r], :row_bytes], rows `pitch_bytes` apart in both buffers
 * (beam state reordering, search.py:205-209; only the filled prefix of each cache row is moved). */
int zb_gather_rows(const void* src, const int32_t* index, void* dst, int64_t rows, int64_t row_bytes,
                   int64_t pitch_bytes, zb_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K7
 * zb_prefix_mean_{fwd,bwd}: y[b,t,:] = (sum_{s<=t} x[b,s,:]) / (t+1)   — Average Attention
 * (models/transformer_aan.py:92-117 with the "aan" bias of func.py:390-398 for an all-ones mask),
 * O(T d) scan instead of the reference's [T,T] matmul.  bwd: dx[b,s,:] = sum_{t>=s} dy[b,t,:]/(t+1).
 * `lens` [batch] (or NULL = all valid) carries the target mask; mode 0 = the masked "aan" bias (pad rows -> 0),
 * mode 1 = cumsum / count (aan_mask=False, transformer_aan.py:103-108).
 */
int zb_prefix_mean_fwd(const void* x, void* y, const int32_t* lens, int32_t batch, int32_t len, int32_t dim,
                       int32_t mode, zb_stream_t stream);
int zb_prefix_mean_bwd(const void* dy, void* dx, const int32_t* lens, int32_t batch, int32_t len, int32_t dim,
                       int32_t mode, zb_stream_t stream);
/* zb_aan_step: cached decode of the average layer, y = (x + sum) / (time + 1); sum += x (fp32 running sum)
 * (models/transformer_aan.py:110-112, func.py:262-272). */
int zb_aan_step(const void* x, float* sum, void* y, int64_t n, int32_t time, zb_stream_t stream);
/* zb_aan_cat_step: zb_aan_step plus the two copies around it in one launch: sum += x, y = sum / (time + 1),
 * cat[r] = [x[r] | y[r]] (row pitch ldcat >= 2 * dim; the tf.concat of transformer_aan.py:185). */
int zb_aan_cat_step(const void* x, float* sum, void* cat, int64_t ldcat, void* y, int64_t rows, int32_t dim,
                    int32_t time, zb_stream_t stream);
/* zb_aan_gate_ln: out = LayerNorm(x + (sigmoid(i) * x + sigmoid(f) * y)), z = [i | f]: zb_aan_gate_fwd followed by
 * zb_add_ln_fwd in one pass, bit-identical to the pair (transformer_aan.py:185-192). */
int zb_aan_gate_ln(const void* x, const void* y, const void* z, void* out, const float* scale, const float* offset,
                   int64_t rows, int32_t dim, float eps, zb_stream_t stream);
/* zb_aan_gate_{fwd,bwd}: out = sigmoid(i) * x + sigmoid(f) * y with z = [i | f] (transformer_aan.py:185-189). */
int zb_aan_gate_fwd(const void* x, const void* y, const void* z, void* out, int64_t rows, int32_t dim,
                    zb_stream_t stream);
int zb_aan_gate_bwd(const void* x, const void* y, const void* z, const void* dout, void* dx, void* dy, void* dz,
                    int64_t rows, int32_t dim, zb_stream_t stream);
/* zb_gated_rms_{fwd,bwd}: ReLA's post-attention norm scale * x * rsqrt(mean(x^2) + eps) * sigmoid(gate * x)
 * (modules/rela.py:95-109).  rstd [rows] fp32 saved by fwd; bwd accumulates dscale / dgate (fp32 [cols]). */
int zb_gated_rms_fwd(const void* x, void* out, float* rstd, const float* scale, const float* gate, int64_t rows,
                     int64_t cols, float eps, zb_stream_t stream);
int zb_gated_rms_bwd(const void* x, const void* dy, const float* rstd, const float* scale, const float* gate,
                     void* dx, float* dscale, float* dgate, int64_t rows, int64_t cols, zb_stream_t stream);
/* zb_add2d: out[r, :cols] = a[r, :cols] (+ b[r, :cols]) over strided 2-D bf16 views, b optional (plain copy).
 * Merged attention o_cross + aan_o (func.py:274-275); tf.concat([x, y], -1) of the AAN gate input
 * (models/transformer_aan.py:185) as two strided copies. */
int zb_add2d(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo, int64_t rows,
             int64_t cols, zb_stream_t stream);
/* zb_dropout: out[i] = x[i] (+ x2[i]) kept with probability 1 - rate and scaled by 1 / (1 - rate), else 0
 * (tf.nn.dropout via util.valid_apply_dropout, utils/util.py:75-79).  bf16 [n]; out may alias x; x2 optional second
 * addend (gradient fan-in).  The mask is a pure function of (*seed, site, i): the backward pass applies the SAME call
 * to the upstream gradient.  Sites: embedding dropout (models/transformer.py:33,119), relu dropout
 * (func.py:334), residual dropout (func.py:323).  seed: device pointer to one uint64. */
int zb_dropout(const void* x, const void* x2, void* out, int64_t n, float rate, const uint64_t* seed, uint32_t site,
               zb_stream_t stream);
/* zb_gumbel_add: x[i] += -log(-log(u_i + eps) + eps), u_i uniform in [0, 1) — util.gumbel_noise (utils/util.py:189-195)
 * added to the step logits when params.enable_noise_beam_search is set (search.py:143-145: Gumbel top-k sampling
 * without replacement).  fp32 [n] in place; u is a pure function of (*seed, site, i) (the counter-based generator of
 * zb_dropout), so the caller varies `site` per decode step and bumps *seed per search. */
int zb_gumbel_add(float* x, int64_t n, float eps, const uint64_t* seed, uint32_t site, zb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ZERO_B200_H_ */
