"""Thin Python wrappers: torch.Tensor (buffer carrier) -> raw pointers + current stream -> C ABI.

Each function cites the reference call site it stands in for.  No arithmetic happens in Python/torch here.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as L

bf16 = torch.bfloat16


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _rowmajor2d(t):
    assert t.dim() == 2 and t.stride(1) == 1, "expected a row-major 2-D view"
    return t.stride(0)


def gemm(a, b, out, a_layout=L.ZB_K_MAJOR, b_layout=L.ZB_MN_MAJOR, bias=None, relu=False, accum=False,
         relu_mask=None, alpha=1.0, split_k=0, m=None, n=None, k=None):
    """D[m,n] (+)= alpha * sum_k A(m,k) B(n,k) (+bias) (relu) (*mask>0).  a/b/out: 2-D row-major views.
    Layout semantics as in include/zero_b200.h: K-major operand is stored [rows, k]; MN-major is stored [k, rows].
    func.linear (func.py:49,59): gemm(x, W, y, b_layout=MN_MAJOR, bias=b)."""
    args = gemm_args(a, b, out, a_layout, b_layout, bias, relu, accum, relu_mask, alpha, split_k, m, n, k)
    L.check(L.load().zb_gemm(C.byref(args), _stream()), "zb_gemm")
    return out


def gemm_args(a, b, out, a_layout=L.ZB_K_MAJOR, b_layout=L.ZB_MN_MAJOR, bias=None, relu=False, accum=False,
              relu_mask=None, alpha=1.0, split_k=0, m=None, n=None, k=None):
    """The zb_gemm_args record of one problem (see gemm)."""
    M = m if m is not None else (a.shape[0] if a_layout == L.ZB_K_MAJOR else a.shape[1])
    Kd = k if k is not None else (a.shape[1] if a_layout == L.ZB_K_MAJOR else a.shape[0])
    N = n if n is not None else (b.shape[0] if b_layout == L.ZB_K_MAJOR else b.shape[1])
    flags = 0
    if bias is not None:
        flags |= L.ZB_EPI_BIAS
    if relu:
        flags |= L.ZB_EPI_RELU
    if accum:
        flags |= L.ZB_EPI_ACCUM
    if relu_mask is not None:
        flags |= L.ZB_EPI_RELU_MASK
    return L.GemmArgs(
        _p(a), _p(b), _p(out), M, N, Kd, _rowmajor2d(a), _rowmajor2d(b), _rowmajor2d(out),
        a_layout, b_layout, L.ZB_F32 if out.dtype == torch.float32 else L.ZB_BF16, flags,
        _p(bias), _p(relu_mask), _rowmajor2d(relu_mask) if relu_mask is not None else 0, float(alpha), int(split_k))


def gemm_grouped(problems):
    """zb_gemm_grouped over a list of gemm_args records (one persistent launch for a layer's weight gradients)."""
    if not problems:
        return
    arr = (L.GemmArgs * len(problems))(*problems)
    L.check(L.load().zb_gemm_grouped(arr, len(problems), _stream()), "zb_gemm_grouped")


def wgrad_args(x, dy, dw):
    """Problem record of linear_wgrad (dW += x^T dy)."""
    return gemm_args(x, dy, dw, L.ZB_MN_MAJOR, L.ZB_MN_MAJOR, accum=True)


def colsum_grouped(pairs):
    """db_i += sum_rows x_i for every (x_i, db_i): the bias gradients of one layer in one launch."""
    if not pairs:
        return
    arr = (L.ColsumArgs * len(pairs))(*[L.ColsumArgs(_p(x), x.shape[0], x.shape[1], x.stride(0), _p(o))
                                       for x, o in pairs])
    L.check(L.load().zb_colsum_grouped(arr, len(pairs), _stream()), "zb_colsum_grouped")


def linear_fwd(x, w, bias, out, relu=False):
    """func.linear: out = x @ W + b (func.py:49,59), optional fused relu (func.py:332)."""
    return gemm(x, w, out, L.ZB_K_MAJOR, L.ZB_MN_MAJOR, bias=bias, relu=relu)


def linear_dgrad(dy, w, dx, relu_mask=None, accum=False, alpha=1.0):
    """dx = alpha * dy @ W^T (gradient of func.linear wrt its input); optional relu mask of the producer."""
    return gemm(dy, w, dx, L.ZB_K_MAJOR, L.ZB_K_MAJOR, relu_mask=relu_mask, accum=accum, alpha=alpha)


def linear_wgrad(x, dy, dw):
    """dW += x^T @ dy into the fp32 gradient arena (split-K, atomic accumulate)."""
    return gemm(x, dy, dw, L.ZB_MN_MAJOR, L.ZB_MN_MAJOR, accum=True)


def colsum(x, out):
    """db += sum_rows x (gradient of tf.nn.bias_add, func.py:59)."""
    lib = L.load()
    L.check(lib.zb_colsum(_p(x), x.shape[0], x.shape[1], x.stride(0), _p(out), _stream()), "zb_colsum")


def attention_args(q, k, v, o, heads, key_len=None, causal=False, q_offset=0, inf_value=1e8, lse=None,
                   rpr_k=None, rpr_v=None, max_rel=0, relu_attn=False, kv_group=1, dropout=None):
    """q/k/v/o: [B, L, heads*dh] views (last dim contiguous).
    dropout: None or (rate, site, seed_tensor) — attention dropout of func.py:245."""
    B, Lq, D = q.shape
    assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1 and o.stride(2) == 1
    Lk = k.shape[1]
    dh = D // heads
    a = L.AttentionArgs()
    a.q, a.k, a.v, a.o = _p(q), _p(k), _p(v), _p(o)
    a.ldq, a.ldk, a.ldv, a.ldo = q.stride(1), k.stride(1), v.stride(1), o.stride(1)
    a.bsq, a.bsk, a.bsv, a.bso = q.stride(0), k.stride(0), v.stride(0), o.stride(0)
    a.batch, a.heads, a.lq, a.lk, a.dh = B, heads, Lq, Lk, dh
    a.key_len = _p(key_len)
    a.causal, a.q_offset = int(causal), int(q_offset)
    a.scale, a.inf_value = float(dh) ** -0.5, float(inf_value)
    a.lse = _p(lse)
    a.rpr_k, a.rpr_v, a.max_rel = _p(rpr_k), _p(rpr_v), int(max_rel)
    a.relu_attn = int(relu_attn)
    a.kv_group = int(kv_group)
    if dropout is not None and dropout[0] > 0.0:
        a.dropout_rate, a.dropout_site, a.dropout_seed = float(dropout[0]), int(dropout[1]), _p(dropout[2])
    return a


def dropout(x, out, rate, seed, site, x2=None):
    """out = tf.nn.dropout(x (+ x2), keep_prob = 1 - rate) with the counter-based mask of (seed, site)
    (utils/util.py:75-79).  bf16, contiguous; out may be x.  seed: device uint64/int64 tensor [1]."""
    assert x.is_contiguous() and out.is_contiguous() and (x2 is None or x2.is_contiguous())
    L.check(L.load().zb_dropout(_p(x), _p(x2), _p(out), x.numel(), float(rate), _p(seed), int(site), _stream()),
            "zb_dropout")
    return out


def gumbel_add(logits, seed, site, eps=1e-8):
    """logits += util.gumbel_noise(shape) (utils/util.py:189-195; search.py:143-145).  fp32, contiguous, in place."""
    assert logits.dtype == torch.float32 and logits.is_contiguous()
    L.check(L.load().zb_gumbel_add(_p(logits), logits.numel(), float(eps), _p(seed), int(site), _stream()),
            "zb_gumbel_add")
    return logits


def attention_fwd(a):
    """func.dot_attention core (func.py:218-256)."""
    L.check(L.load().zb_attention_fwd(C.byref(a), _stream()), "zb_attention_fwd")


def attention_bwd_workspace_bytes(a):
    """Bytes of scratch that let zb_attention_bwd pick any kernel for the problem `a` (0: none needed)."""
    return int(L.load().zb_attention_bwd_workspace_bytes(C.byref(a)))


def attention_bwd(a, d_o, dq, dk, dv, delta, d_rpr_k=None, d_rpr_v=None, workspace=None):
    """`workspace`: caller-owned scratch tensor (any dtype, >= attention_bwd_workspace_bytes(a) bytes), or a callable
    bytes -> tensor that is asked only when the problem needs one."""
    need = attention_bwd_workspace_bytes(a)
    if need and callable(workspace):
        workspace = workspace(need)
    if need and workspace is not None:
        assert workspace.numel() * workspace.element_size() >= need
        a.workspace, a.workspace_bytes = _p(workspace), workspace.numel() * workspace.element_size()
    a.d_o, a.dq, a.dk, a.dv = _p(d_o), _p(dq), _p(dk), _p(dv)
    a.lddo, a.lddq, a.lddk, a.lddv = d_o.stride(1), dq.stride(1), dk.stride(1), dv.stride(1)
    a.bsdo, a.bsdq, a.bsdk, a.bsdv = d_o.stride(0), dq.stride(0), dk.stride(0), dv.stride(0)
    a.delta = _p(delta)
    a.d_rpr_k, a.d_rpr_v = _p(d_rpr_k), _p(d_rpr_v)
    L.check(L.load().zb_attention_bwd(C.byref(a), _stream()), "zb_attention_bwd")


def add_ln_fwd(x, y, out, scale, offset, mean=None, rstd=None, eps=1e-8, y32=None, ybias=None):
    """func.residual_fn + func.layer_norm (func.py:321-324, 289-303).  y32 / ybias: the branch output as the fp32
    accumulator of a split-K projection plus that projection's bias (cleared by the kernel after it is read)."""
    a = L.AddLnArgs()
    a.x, a.y, a.out, a.mean, a.rstd = _p(x), _p(y), _p(out), _p(mean), _p(rstd)
    a.y32, a.ybias = _p(y32), _p(ybias)
    a.scale, a.offset = _p(scale), _p(offset)
    a.rows, a.cols, a.eps = x.numel() // x.shape[-1], x.shape[-1], float(eps)
    L.check(L.load().zb_add_ln_fwd(C.byref(a), _stream()), "zb_add_ln_fwd")


def add_ln_bwd(x, y, d_out, d_out2, mean, rstd, scale, ds, dscale, doffset, dbias=None):
    a = L.AddLnArgs()
    a.x, a.y, a.mean, a.rstd, a.scale = _p(x), _p(y), _p(mean), _p(rstd), _p(scale)
    a.rows, a.cols = x.numel() // x.shape[-1], x.shape[-1]
    a.d_out, a.d_out2, a.ds, a.dscale, a.doffset = _p(d_out), _p(d_out2), _p(ds), _p(dscale), _p(doffset)
    a.dbias = _p(dbias)
    L.check(L.load().zb_add_ln_bwd(C.byref(a), _stream()), "zb_add_ln_bwd")


def embed_fwd(ids, table, bias, out, mult, shift=0, zero_if_all_pad=False, time=-1):
    """tf.gather * sqrt(d) + bias + timing signal (models/transformer.py:29-31,104-117)."""
    a = L.EmbedArgs()
    a.ids, a.table, a.bias, a.out = _p(ids), _p(table), _p(bias), _p(out)
    a.batch, a.len, a.dim, a.vocab = ids.shape[0], ids.shape[1], table.shape[1], table.shape[0]
    a.shift, a.zero_if_all_pad, a.time, a.mult = int(shift), int(zero_if_all_pad), int(time), float(mult)
    L.check(L.load().zb_embed_fwd(C.byref(a), _stream()), "zb_embed_fwd")


def embed_bwd(ids, d_out, d_table, d_bias, mult, shift=0, d_out2=None):
    a = L.EmbedArgs()
    a.ids, a.d_out, a.d_table, a.d_bias, a.d_out2 = _p(ids), _p(d_out), _p(d_table), _p(d_bias), _p(d_out2)
    a.batch, a.len, a.dim, a.vocab = ids.shape[0], ids.shape[1], d_table.shape[1], d_table.shape[0]
    a.shift, a.mult = int(shift), float(mult)
    L.check(L.load().zb_embed_bwd(C.byref(a), _stream()), "zb_embed_bwd")


def softmax_ce(logits, labels, nll, smooth, d_logits=None, per_sample=None, loss=None, loss_scale=1.0):
    """label-smoothed CE + masked per-sample mean + batch mean (models/transformer.py:198-211)."""
    a = L.CeArgs()
    a.logits, a.ld, a.labels = _p(logits), logits.stride(0), _p(labels)
    a.batch, a.seq_len, a.nll = labels.shape[0], labels.shape[1], _p(nll)
    a.d_logits, a.ldd = _p(d_logits), (d_logits.stride(0) if d_logits is not None else 0)
    a.vocab, a.smooth, a.loss_scale = logits.shape[1], float(smooth), float(loss_scale)
    a.per_sample, a.loss = _p(per_sample), _p(loss)
    L.check(L.load().zb_softmax_ce(C.byref(a), _stream()), "zb_softmax_ce")


def vocab_ce(feat, table, labels, nll, smooth, workspace, d_logits=None, per_sample=None, loss=None, loss_scale=1.0):
    """K6: tied-softmax projection + label-smoothed CE without materialising the logits (models/transformer.py:186-211).
    feat bf16 [rows, d]; table bf16 [V, d]; labels int32 [B, T]; workspace: callable bytes -> tensor, or a tensor."""
    a = L.VocabCeArgs()
    a.feat, a.ldf, a.table, a.ldt = _p(feat), feat.stride(0), _p(table), table.stride(0)
    a.labels, a.batch, a.seq_len = _p(labels), labels.shape[0], labels.shape[1]
    a.d, a.vocab = feat.shape[1], table.shape[0]
    a.smooth, a.loss_scale = float(smooth), float(loss_scale)
    a.nll, a.per_sample, a.loss = _p(nll), _p(per_sample), _p(loss)
    a.d_logits, a.ldd = _p(d_logits), (d_logits.stride(0) if d_logits is not None else 0)
    need = int(L.load().zb_vocab_ce_workspace_bytes(C.byref(a)))
    ws = workspace(need) if callable(workspace) else workspace
    a.workspace, a.workspace_bytes = _p(ws), ws.numel() * ws.element_size()
    L.check(L.load().zb_vocab_ce(C.byref(a), _stream()), "zb_vocab_ce")


class BeamCandidates(object):
    """What zb_vocab_topk leaves for zb_beam_step in place of the [rows, V] logits (csrc/vocab_topk.cu): `buffer` holds
    per (row, 128-column part) the soft-max statistics and the 8 largest logits with their columns."""

    def __init__(self, buffer, rows, vocab, parts, skip_col, temperature):
        self.buffer, self.rows, self.vocab, self.parts = buffer, rows, vocab, parts
        self.skip_col, self.temperature = skip_col, temperature

    def unpack(self):
        """(stats [parts, rows, 4], values [parts, rows, 8], columns [parts, rows, 8]) views, for tests."""
        n = self.parts * self.rows
        flat = self.buffer.view(torch.float32).reshape(-1)
        stats = flat[:4 * n].view(self.parts, self.rows, 4)
        vals = flat[4 * n:12 * n].view(self.parts, self.rows, 8)
        cols = flat[12 * n:20 * n].view(torch.int32).view(self.parts, self.rows, 8)
        return stats, vals, cols


def vocab_topk(feat, table, workspace, skip_col=-1, temperature=1.0):
    """K8 fused: the decode step's vocabulary projection reduced to beam-search candidates inside the GEMM epilogue
    (models/transformer.py:186-196 + search.py:147-176); the logits are never written.  feat bf16 [rows, d]; table bf16
    [V, d]; workspace: callable bytes -> tensor, or a tensor.  Returns BeamCandidates for BeamState.step."""
    a = L.VocabTopkArgs()
    a.feat, a.ldf, a.table, a.ldt = _p(feat), feat.stride(0), _p(table), table.stride(0)
    a.rows, a.d, a.vocab = feat.shape[0], feat.shape[1], table.shape[0]
    a.skip_col, a.temperature = int(skip_col), float(temperature)
    need = int(L.load().zb_vocab_topk_workspace_bytes(C.byref(a)))
    ws = workspace(need) if callable(workspace) else workspace
    a.workspace, a.workspace_bytes = _p(ws), ws.numel() * ws.element_size()
    L.check(L.load().zb_vocab_topk(C.byref(a), _stream()), "zb_vocab_topk")
    return BeamCandidates(ws, int(a.rows), int(a.vocab), int(L.load().zb_vocab_topk_parts(a.vocab)), int(skip_col),
                          float(temperature))


def vocab_topk_supported(d, vocab, beam, feat=None, table=None):
    """Shapes the candidate path takes (the caller otherwise materialises the logits: zb_gemm + zb_beam_step)."""
    ok = vocab >= 128 and d % 8 == 0 and 2 * beam <= 8 and vocab - 1 >= 2 * beam
    for t in (feat, table):
        if t is not None:
            ok = ok and t.stride(0) % 8 == 0 and t.stride(1) == 1 and t.data_ptr() % 16 == 0
    return bool(ok)


def vocab_ce_supported(rows, d, vocab, feat=None, table=None, d_logits=None):
    """Shapes the fused kernel takes (the caller otherwise materialises the logits: zb_gemm + zb_softmax_ce)."""
    ok = vocab >= 128 and d % 8 == 0
    for t in (feat, table, d_logits):
        if t is not None:
            ok = ok and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0
    return ok


def cast_f32_bf16(src, dst):
    L.check(L.load().zb_cast_f32_bf16(_p(src), _p(dst), src.numel(), _stream()), "zb_cast_f32_bf16")


def cast_bf16_f32(src, dst):
    L.check(L.load().zb_cast_bf16_f32(_p(src), _p(dst), src.numel(), _stream()), "zb_cast_bf16_f32")


def sumsq(x, out):
    L.check(L.load().zb_sumsq(_p(x), x.numel(), _p(out), _stream()), "zb_sumsq")


def adam_tf(param, m, v, grad, param_bf16, beta1, beta2, eps, lr_t, grad_scale, clip_scale=None, norms=None):
    """tf.train.AdamOptimizer update on the flat arena (main.py:178-181)."""
    a = L.AdamArgs(_p(param), _p(m), _p(v), _p(grad), _p(param_bf16), param.numel(), float(beta1), float(beta2),
                   float(eps), float(lr_t), float(grad_scale), _p(clip_scale), _p(norms))
    L.check(L.load().zb_adam_tf(C.byref(a), _stream()), "zb_adam_tf")


def shard_adam(lo, n, world, rank, grad_ptrs, mirror_ptrs, param, m, v, beta1, beta2, eps, lr_t, grad_scale,
               flags=L.ZB_SHARD_UPDATE | L.ZB_SHARD_NORM_G | L.ZB_SHARD_NORM_P, grad_mc=0, mirror_mc=0, grad_out=None,
               clip_scale=None, norms=None, norm_parts_ptrs=None, done_counter=None, wide_mask=None, param_ptrs=None,
               param_mc=0, grad_sources=None):
    """Gradient aggregation (utils/parallel.py:134-208, main.py:42-43) + TF Adam (main.py:178-181) + refresh of every
    rank's bf16 compute copy for the shard [lo, lo + n) of the flat arenas, in one kernel over peer memory.
    `grad_ptrs` / `mirror_ptrs` / `norm_parts_ptrs`: per-rank device addresses (ints) of the symmetric arenas as
    mapped in THIS process; `grad_mc` / `mirror_mc`: their multicast addresses (0: use the per-rank addresses).
    param / m / v: the local fp32 arenas (whole tensors; the kernel indexes them from element 0).  `wide_mask`
    (uint8 per 64-element slot) marks the 1-D variables whose fp32 values also go to every rank's master arena
    (`param_ptrs` / `param_mc`): the forward pass reads biases and LayerNorm parameters from the master.
    `grad_sources` (default: world) entries of `grad_ptrs` are summed; `grad_out` is a tensor or a raw address."""
    a = L.ShardAdamArgs()
    a.lo, a.n, a.world, a.rank = int(lo), int(n), int(world), int(rank)
    a.grad_sources = int(world if grad_sources is None else grad_sources)
    a.grad_mc = grad_mc or None
    a.mirror_mc = mirror_mc or None
    for r in range(world):
        a.grad_peer[r] = grad_ptrs[r] if grad_ptrs and r < len(grad_ptrs) else None
        a.mirror_peer[r] = mirror_ptrs[r] if mirror_ptrs else None
        a.norm_parts_peer[r] = norm_parts_ptrs[r] if norm_parts_ptrs else None
        a.param_peer[r] = param_ptrs[r] if param_ptrs else None
    a.wide_mask, a.param_mc = _p(wide_mask), (param_mc or None)
    a.param, a.m, a.v = _p(param), _p(m), _p(v)
    a.grad_out = grad_out if isinstance(grad_out, int) else _p(grad_out)
    a.beta1, a.beta2, a.eps, a.lr_t, a.grad_scale = float(beta1), float(beta2), float(eps), float(lr_t), float(grad_scale)
    a.flags = int(flags)
    a.clip_scale, a.norms, a.done_counter = _p(clip_scale), _p(norms), _p(done_counter)
    L.check(L.load().zb_shard_adam(C.byref(a), _stream()), "zb_shard_adam")


def gather_rows(src, index, dst, row_elems=None):
    """dst[r, :row_elems] = src[index[r], :row_elems] over the leading dim (beam state reordering,
    search.py:205-209).  Rows are dst[0].numel() elements apart; only the first row_elems are moved."""
    rows = dst.shape[0]
    pitch = dst[0].numel() * dst.element_size()
    row_bytes = pitch if row_elems is None else row_elems * dst.element_size()
    L.check(L.load().zb_gather_rows(_p(src), _p(index), _p(dst), rows, row_bytes, pitch, _stream()), "zb_gather_rows")


def prefix_mean_fwd(x, y, lens=None, mode=0):
    """Average Attention prefix mean (models/transformer_aan.py:99-108; func.py:389-398)."""
    L.check(L.load().zb_prefix_mean_fwd(_p(x), _p(y), _p(lens), x.shape[0], x.shape[1], x.shape[2], int(mode),
                                        _stream()), "zb_prefix_mean_fwd")


def prefix_mean_bwd(dy, dx, lens=None, mode=0):
    L.check(L.load().zb_prefix_mean_bwd(_p(dy), _p(dx), _p(lens), dy.shape[0], dy.shape[1], dy.shape[2], int(mode),
                                        _stream()), "zb_prefix_mean_bwd")


def aan_step(x, running_sum, y, time):
    """cached decode of the average layer (models/transformer_aan.py:110-112)."""
    L.check(L.load().zb_aan_step(_p(x), _p(running_sum), _p(y), x.numel(), int(time), _stream()), "zb_aan_step")


def aan_cat_step(x, running_sum, cat, y, time):
    """aan_step + the concat of transformer_aan.py:185 in one launch: cat = [x | y], y = (sum += x) / (time + 1)."""
    assert x.is_contiguous() and y.is_contiguous() and running_sum.is_contiguous() and cat.stride(1) == 1
    L.check(L.load().zb_aan_cat_step(_p(x), _p(running_sum), _p(cat), cat.stride(0), _p(y), x.shape[0], x.shape[1],
                                     int(time), _stream()), "zb_aan_cat_step")


def aan_gate_ln(x, y, z, out, scale, offset, eps):
    """out = LN(x + sigmoid(i) x + sigmoid(f) y), z = [i | f] (transformer_aan.py:185-192) in one launch."""
    assert x.is_contiguous() and y.is_contiguous() and z.is_contiguous() and out.is_contiguous()
    L.check(L.load().zb_aan_gate_ln(_p(x), _p(y), _p(z), _p(out), _p(scale), _p(offset), x.shape[0], x.shape[1],
                                    float(eps), _stream()), "zb_aan_gate_ln")


def aan_gate_fwd(x, y, z, out):
    """y' = sigmoid(i) x + sigmoid(f) y (models/transformer_aan.py:185-189)."""
    L.check(L.load().zb_aan_gate_fwd(_p(x), _p(y), _p(z), _p(out), x.numel() // x.shape[-1], x.shape[-1], _stream()),
            "zb_aan_gate_fwd")


def aan_gate_bwd(x, y, z, dout, dx, dy, dz):
    L.check(L.load().zb_aan_gate_bwd(_p(x), _p(y), _p(z), _p(dout), _p(dx), _p(dy), _p(dz),
                                     x.numel() // x.shape[-1], x.shape[-1], _stream()), "zb_aan_gate_bwd")


def gated_rms_fwd(x, out, rstd, scale, gate, eps):
    """ReLA gated RMS norm (modules/rela.py:95-109)."""
    L.check(L.load().zb_gated_rms_fwd(_p(x), _p(out), _p(rstd), _p(scale), _p(gate), x.numel() // x.shape[-1],
                                      x.shape[-1], float(eps), _stream()), "zb_gated_rms_fwd")


def gated_rms_bwd(x, dy, rstd, scale, gate, dx, dscale, dgate):
    L.check(L.load().zb_gated_rms_bwd(_p(x), _p(dy), _p(rstd), _p(scale), _p(gate), _p(dx), _p(dscale), _p(dgate),
                                      x.numel() // x.shape[-1], x.shape[-1], _stream()), "zb_gated_rms_bwd")


def add2d(a, b, out):
    """out = a (+ b) over row-major 2-D bf16 views (strided allowed); b=None is a strided copy."""
    L.check(L.load().zb_add2d(_p(a), a.stride(0), _p(b), b.stride(0) if b is not None else 0, _p(out), out.stride(0),
                              a.shape[0], a.shape[1], _stream()), "zb_add2d")


def beam_args(**kw):
    a = L.BeamArgs()
    for k, v in kw.items():
        setattr(a, k, _p(v) if isinstance(v, torch.Tensor) else v)
    return a


def beam_cond(a):
    L.check(L.load().zb_beam_cond(C.byref(a), _stream()), "zb_beam_cond")


def beam_step(a):
    L.check(L.load().zb_beam_step(C.byref(a), _stream()), "zb_beam_step")
