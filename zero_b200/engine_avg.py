"""Average-attention family on the engine: transformer_aan (models/transformer_aan.py:120-260) and the merged
attention of transformer_fuse (models/transformer_fuse.py:120-165, func.py:258-275).

Teacher-forced forward (train_fn forward / score_fn) and cached decode are built; the backward pass of these two
variants is the next row to build ("what comes next").
"""
from __future__ import annotations

import torch

from . import lib as L
from . import ops
from .engine import DecodeState, Engine, _lens

bf16, f32 = torch.bfloat16, torch.float32


def _aan_layer_fwd(eng, key, x, B, T, tgt_len, tag):
    """average_attention sublayer in training mode (transformer_aan.py:165-192): returns LN(x + gate(...))."""
    c, ps, ws = eng.cfg, eng.ps, eng.ws
    N = B * T
    xf = ws.get(tag + ".xf", (N, c.d))
    ops.prefix_mean_fwd(x.view(B, T, c.d), xf.view(B, T, c.d), tgt_len, mode=0 if c.aan_mask else 1)
    cat = ws.get(tag + ".cat", (N, 2 * c.d))
    ops.add2d(x, None, cat[:, :c.d])
    if c.use_ffn:
        h = ws.get(tag + ".h", (N, c.f))
        ops.linear_fwd(xf, ps.w(key + ".aan.ffn.w1.W"), ps.p(key + ".aan.ffn.w1.b"), h, relu=True)
        ops.linear_fwd(h, ps.w(key + ".aan.ffn.w2.W"), ps.p(key + ".aan.ffn.w2.b"), cat[:, c.d:])
    else:
        ops.add2d(xf, None, cat[:, c.d:])
    y0 = ws.get(tag + ".y0", (N, c.d))
    ops.add2d(cat[:, c.d:], None, y0)
    z = ws.get(tag + ".z", (N, 2 * c.d))
    ops.linear_fwd(cat, ps.w(key + ".aan.z.W"), ps.p(key + ".aan.z.b"), z)
    yg = ws.get(tag + ".yg", (N, c.d))
    ops.aan_gate_fwd(x, y0, z, yg)
    x1 = ws.get(tag + ".x1", (N, c.d))
    ops.add_ln_fwd(x, yg, x1, ps.p(key + ".aan.ln.scale"), ps.p(key + ".aan.ln.offset"), eps=c.eps)
    return x1


def _decode_train_avg(self, target, enc, src_len, S, smooth, want_grad, save=None, tag="D"):
    c, ps, ws = self.cfg, self.ps, self.ws
    if want_grad:
        raise L.ZeroB200Error("backward pass of %s is not built yet (forward / score / decode are)" % c.model)
    B, T = target.shape
    N = B * T
    tgt_len = _lens(target)
    x = ws.get(tag + ".x0", (N, c.d))
    ops.embed_fwd(target, ps.w(self._tgt_table()), ps.p("emb_bias"), x, mult=c.d ** 0.5, shift=1)
    for l in range(c.ndec):
        key, t = "dec%d" % l, "%s.A%d" % (tag, l)
        sv = {"cross": {}, "ffn": {}, "lnc": {}, "ln2": {}}
        if c.aan:
            x1 = _aan_layer_fwd(self, key, x, B, T, tgt_len, t)
            yc = self._cross_attn_fwd(key + ".cross", x1, enc, B, T, S, src_len, sv["cross"], t + ".cross")
            xc = self._ln_fwd(key + ".cross.ln", x1, yc, N, sv["lnc"], t + ".lnc")
        else:
            # merged attention: o = cross_attention(x) + prefix_mean(v_map(x)), then o_map (func.py:258-278)
            kc = key + ".cross"
            q = ws.get(t + ".q", (N, c.d))
            ops.linear_fwd(x, ps.w(kc + ".q.W"), ps.p(kc + ".q.b"), q)
            kv = ws.get(t + ".kv", (B * S, 2 * c.d))
            ops.linear_fwd(enc, ps.w(kc + ".kv.W"), ps.p(kc + ".kv.b"), kv)
            kv3 = kv.view(B, S, 2 * c.d)
            ctx = ws.get(t + ".ctx", (N, c.d))
            a = ops.attention_args(q.view(B, T, c.d), kv3[:, :, :c.d], kv3[:, :, c.d:], ctx.view(B, T, c.d), c.h,
                                   key_len=src_len, inf_value=c.inf)
            ops.attention_fwd(a)
            vq = ws.get(t + ".vq", (N, c.d))
            ops.linear_fwd(x, ps.w(kc + ".kv.W")[:, c.d:], ps.p(kc + ".kv.b")[c.d:], vq)
            av = ws.get(t + ".av", (N, c.d))
            ops.prefix_mean_fwd(vq.view(B, T, c.d), av.view(B, T, c.d), tgt_len, mode=0)
            ops.add2d(ctx, av, ctx)
            yc = ws.get(t + ".y", (N, c.d))
            ops.linear_fwd(ctx, ps.w(kc + ".o.W"), ps.p(kc + ".o.b"), yc)
            xc = self._ln_fwd(kc + ".ln", x, yc, N, sv["lnc"], t + ".lnc")
        y2 = self._ffn_fwd(key + ".ffn", xc, N, sv["ffn"], t + ".ffn")
        x = self._ln_fwd(key + ".ffn.ln", xc, y2, N, sv["ln2"], t + ".ln2")
    logits = ws.get(tag + ".logits", (N, c.vt), f32)
    ops.gemm(x, ps.w(self._softmax_table()), logits, L.ZB_K_MAJOR, L.ZB_K_MAJOR)
    nll = ws.get(tag + ".nll", (N,), f32)
    per_sample = ws.get(tag + ".per_sample", (B,), f32)
    loss = ws.get(tag + ".loss", (1,), f32)
    ops.softmax_ce(logits, target, nll, smooth, per_sample=per_sample, loss=loss, loss_scale=c.loss_scale)
    return loss, per_sample, logits


def _decoding_fn_avg(self, target, state, time):
    """Cached decode step of transformer_aan / transformer_fuse: the growing K/V cache of self-attention is
    replaced by one fp32 running sum per layer (transformer_aan.py:110-112; func.py:262-272)."""
    c, ps, ws = self.cfg, self.ps, self.ws
    t = int(time)
    R = target.shape[0]
    K = state.K
    x = ws.get("dec.x", (R, c.d))
    ops.embed_fwd(target, ps.w(self._tgt_table()), ps.p("emb_bias"), x.view(R, 1, c.d), mult=c.d ** 0.5,
                  zero_if_all_pad=True, time=t)
    y = ws.get("dec.y", (R, c.d))
    ctx = ws.get("dec.ctx", (R, c.d))
    for l in range(c.ndec):
        key = "dec%d" % l
        kc = key + ".cross"
        mem = state.mem[l]
        if c.aan:
            xf = ws.get("dec.xf", (R, c.d))
            ops.aan_step(x, state.sums[l], xf, t)
            cat = ws.get("dec.cat", (R, 2 * c.d))
            ops.add2d(x, None, cat[:, :c.d])
            if c.use_ffn:
                h = ws.get("dec.h", (R, c.f))
                ops.linear_fwd(xf, ps.w(key + ".aan.ffn.w1.W"), ps.p(key + ".aan.ffn.w1.b"), h, relu=True)
                ops.linear_fwd(h, ps.w(key + ".aan.ffn.w2.W"), ps.p(key + ".aan.ffn.w2.b"), cat[:, c.d:])
            else:
                ops.add2d(xf, None, cat[:, c.d:])
            y0 = ws.get("dec.y0", (R, c.d))
            ops.add2d(cat[:, c.d:], None, y0)
            z = ws.get("dec.z", (R, 2 * c.d))
            ops.linear_fwd(cat, ps.w(key + ".aan.z.W"), ps.p(key + ".aan.z.b"), z)
            ops.aan_gate_fwd(x, y0, z, y)
            x1 = ws.get("dec.x1", (R, c.d))
            ops.add_ln_fwd(x, y, x1, ps.p(key + ".aan.ln.scale"), ps.p(key + ".aan.ln.offset"), eps=c.eps)
        else:
            x1 = x
        q = ws.get("dec.q", (R, c.d))
        ops.linear_fwd(x1, ps.w(kc + ".q.W"), ps.p(kc + ".q.b"), q)
        a = ops.attention_args(q.view(R, 1, c.d), mem[:, :, :c.d], mem[:, :, c.d:], ctx.view(R, 1, c.d), c.h,
                               key_len=state.src_len, q_offset=t, inf_value=c.inf, kv_group=K)
        ops.attention_fwd(a)
        if c.fuse:
            vq = ws.get("dec.vq", (R, c.d))
            ops.linear_fwd(x1, ps.w(kc + ".kv.W")[:, c.d:], ps.p(kc + ".kv.b")[c.d:], vq)
            av = ws.get("dec.av", (R, c.d))
            ops.aan_step(vq, state.sums[l], av, t)
            ops.add2d(ctx, av, ctx)
        ops.linear_fwd(ctx, ps.w(kc + ".o.W"), ps.p(kc + ".o.b"), y)
        xc = ws.get("dec.xc", (R, c.d))
        ops.add_ln_fwd(x1, y, xc, ps.p(kc + ".ln.scale"), ps.p(kc + ".ln.offset"), eps=c.eps)
        h = ws.get("dec.h", (R, c.f))
        ops.linear_fwd(xc, ps.w(key + ".ffn.w1.W"), ps.p(key + ".ffn.w1.b"), h, relu=True)
        ops.linear_fwd(h, ps.w(key + ".ffn.w2.W"), ps.p(key + ".ffn.w2.b"), y)
        ops.add_ln_fwd(xc, y, x, ps.p(key + ".ffn.ln.scale"), ps.p(key + ".ffn.ln.offset"), eps=c.eps)
    logits = ws.get("dec.logits", (R, c.vt), f32)
    ops.gemm(x, ps.w(self._softmax_table()), logits, L.ZB_K_MAJOR, L.ZB_K_MAJOR)
    return logits, state


# ---- DecodeState: running sums instead of K/V caches
_orig_begin = DecodeState.begin_search
_orig_reorder = DecodeState.reorder


def _begin_search(self, beam, cap=None):
    _orig_begin(self, beam, cap)
    c, eng = self.engine.cfg, self.engine
    if c.aan or c.fuse:
        R = self.B * self.K
        self.sums = [eng.ws.get("dec.sumA%d" % l, (R, c.d), f32) for l in range(c.ndec)]
        self.sums_alt = [eng.ws.get("dec.sumB%d" % l, (R, c.d), f32) for l in range(c.ndec)]
        for s in self.sums:
            s.zero_()


_orig_swap = DecodeState.swap_buffers


def _swap_buffers(self):
    c = self.engine.cfg
    if c.aan or c.fuse:
        self.sums, self.sums_alt = self.sums_alt, self.sums
        return
    _orig_swap(self)


def _reorder(self, parent, t):
    c = self.engine.cfg
    if c.aan or c.fuse:
        for l in range(c.ndec):
            ops.gather_rows(self.sums[l], parent, self.sums_alt[l])
        self.swap_buffers()
        return
    _orig_reorder(self, parent, t)


DecodeState.begin_search = _begin_search
DecodeState.swap_buffers = _swap_buffers
DecodeState.reorder = _reorder
Engine._decode_train_avg = _decode_train_avg
Engine._decoding_fn_avg = _decoding_fn_avg
