"""Average-attention family on the engine: transformer_aan (models/transformer_aan.py:120-260) and the merged
attention of transformer_fuse (models/transformer_fuse.py:120-165, func.py:258-275): teacher-forced forward,
backward, and cached decode (one fp32 running sum per layer instead of a growing K/V cache).
"""
from __future__ import annotations

import os

import torch

from . import ops
from .engine import DecodeState, Engine, _lens, dropout_site

bf16, f32 = torch.bfloat16, torch.float32


# ------------------------------------------------------------------------------------------------ training forward
def _aan_sublayer_fwd(eng, key, x, B, T, tgt_len, sv, tag):
    """average_attention sublayer (transformer_aan.py:165-192): LN(x + sigmoid(i) x + sigmoid(f) y)."""
    c, ps, ws = eng.cfg, eng.ps, eng.ws
    N = B * T
    xf = ws.get(tag + ".xf", (N, c.d))
    ops.prefix_mean_fwd(x.view(B, T, c.d), xf.view(B, T, c.d), tgt_len, mode=0 if c.aan_mask else 1)
    cat = ws.get(tag + ".cat", (N, 2 * c.d))  # tf.concat([x, y], -1) (transformer_aan.py:185)
    ops.add2d(x, None, cat[:, :c.d])
    if c.use_ffn:
        h = ws.get(tag + ".h", (N, c.f))
        ops.linear_fwd(xf, ps.w(key + ".aan.ffn.w1.W"), ps.p(key + ".aan.ffn.w1.b"), h, relu=True)
        r = eng._rate("relu")
        if r > 0.0:   # transformer_aan.py:180
            ops.dropout(h, h, r, eng.drop_seed, dropout_site(key + ".aan.relu"))
        ops.linear_fwd(h, ps.w(key + ".aan.ffn.w2.W"), ps.p(key + ".aan.ffn.w2.b"), cat[:, c.d:])
        sv["h"] = h
        sv["relu_rate"] = r
    else:
        ops.add2d(xf, None, cat[:, c.d:])
    y0 = ws.get(tag + ".y0", (N, c.d))
    ops.add2d(cat[:, c.d:], None, y0)
    z = ws.get(tag + ".z", (N, 2 * c.d))
    ops.linear_fwd(cat, ps.w(key + ".aan.z.W"), ps.p(key + ".aan.z.b"), z)
    yg = ws.get(tag + ".yg", (N, c.d))
    ops.aan_gate_fwd(x, y0, z, yg)
    sv.update(xf=xf, cat=cat, y0=y0, z=z, yg=yg, ln={})
    return eng._ln_fwd(key + ".aan.ln", x, yg, N, sv["ln"], tag + ".ln")


def _fuse_sublayer_fwd(eng, key, x, enc, B, T, S, src_len, tgt_len, sv, tag):
    """merged attention (func.py:206-278 with fuse_mask): o_map(cross_attention(x) + prefix_mean(v_map(x)))."""
    c, ps, ws = eng.cfg, eng.ps, eng.ws
    N = B * T
    kc = key + ".cross"
    q = ws.get(tag + ".q", (N, c.d))
    ops.linear_fwd(x, ps.w(kc + ".q.W"), ps.p(kc + ".q.b"), q)
    kv = ws.get(tag + ".kv", (B * S, 2 * c.d))
    ops.linear_fwd(enc, ps.w(kc + ".kv.W"), ps.p(kc + ".kv.b"), kv)
    kv3 = kv.view(B, S, 2 * c.d)
    ctx = ws.get(tag + ".ctx", (N, c.d))
    lse = ws.get(tag + ".lse", (B, c.h, T), f32)
    a = ops.attention_args(q.view(B, T, c.d), kv3[:, :, :c.d], kv3[:, :, c.d:], ctx.view(B, T, c.d), c.h,
                           key_len=src_len, inf_value=c.inf, lse=lse,
                           dropout=(eng._rate("att"), dropout_site(kc + ".att"), eng.drop_seed))
    ops.attention_fwd(a)
    vq = ws.get(tag + ".vq", (N, c.d))
    ops.linear_fwd(x, ps.w(kc + ".kv.W")[:, c.d:], ps.p(kc + ".kv.b")[c.d:], vq)  # v_map applied to the query
    av = ws.get(tag + ".av", (N, c.d))
    ops.prefix_mean_fwd(vq.view(B, T, c.d), av.view(B, T, c.d), tgt_len, mode=0)
    osum = ws.get(tag + ".osum", (N, c.d))
    ops.add2d(ctx, av, osum)
    yc = ws.get(tag + ".y", (N, c.d))
    ops.linear_fwd(osum, ps.w(kc + ".o.W"), ps.p(kc + ".o.b"), yc)
    sv.update(q=q, kv=kv, ctx=ctx, lse=lse, attn=a, osum=osum, ln={})
    return eng._ln_fwd(kc + ".ln", x, yc, N, sv["ln"], tag + ".ln")


def _decode_train_avg(self, target, enc, src_len, S, smooth, want_grad, save=None, tag="D", want_logits=True):
    c, ps, ws = self.cfg, self.ps, self.ws
    B, T = target.shape
    N = B * T
    tgt_len = ws.get(tag + ".tgt_len", (B,), torch.int32)
    tgt_len.copy_(_lens(target))
    x = ws.get(tag + ".x0", (N, c.d))
    ops.embed_fwd(target, ps.w(self._tgt_table()), ps.p("emb_bias"), x, mult=c.d ** 0.5, shift=1)
    r_emb = self._rate("emb")
    if r_emb > 0.0:   # transformer_aan.py:152 / transformer_fuse.py:118
        ops.dropout(x, x, r_emb, self.drop_seed, dropout_site("dec.emb"))
    layers = []
    for l in range(c.ndec):
        key, t = "dec%d" % l, "%s.A%d" % (tag, l)
        sv = {"sub": {}, "cross": {}, "ffn": {}, "lnc": {}, "ln2": {}, "x_in": x}
        if c.aan:
            x1 = _aan_sublayer_fwd(self, key, x, B, T, tgt_len, sv["sub"], t + ".aan")
            yc = self._cross_attn_fwd(key + ".cross", x1, enc, B, T, S, src_len, sv["cross"], t + ".cross")
            xc = self._ln_fwd(key + ".cross.ln", x1, yc, N, sv["lnc"], t + ".lnc")
            sv["x1"] = x1
        else:
            xc = _fuse_sublayer_fwd(self, key, x, enc, B, T, S, src_len, tgt_len, sv["sub"], t + ".fuse")
        y2 = self._ffn_fwd(key + ".ffn", xc, N, sv["ffn"], t + ".ffn")
        x = self._ln_fwd(key + ".ffn.ln", xc, y2, N, sv["ln2"], t + ".ln2")
        sv["xc"] = xc
        layers.append(sv)
    feat = x
    loss, per_sample, logits, dlogits = self._vocab_loss(feat, target, smooth, want_grad, want_logits, tag)
    if save is not None:
        save.update(layers=layers, target=target, B=B, T=T, S=S, feat=feat, dlogits=dlogits, enc=enc,
                    tgt_len=tgt_len, emb_rate=r_emb)
    return loss, per_sample, logits


# ------------------------------------------------------------------------------------------------ training backward
def _avg_layer_bwd(self, key, bw, sv, ds2, dxc, enc, d_enc_f32, B, T, S, save):
    """Backward of the attention part of one aan / fuse decoder layer.  (ds2, dxc) are the two addends of the
    gradient wrt xc (the input of the feed-forward sublayer).  Returns the two addends of d loss / d x_in."""
    c, ps, ws = self.cfg, self.ps, self.ws
    N = B * T
    tgt_len = save["tgt_len"]
    sub = sv["sub"]
    x = sv["x_in"]
    if c.aan:
        dsc, dyc = self._ln_bwd(key + ".cross.ln", ds2, dxc, N, sv["lnc"], bw + ".lnc", ps.g(key + ".cross.o.b"))
        dx1 = self._cross_attn_bwd(key + ".cross", sv["x1"], enc, dyc, d_enc_f32, B, T, S, sv["cross"], bw + ".cross")
        # LN(x + gate): ds is the gradient wrt x (skip), dyg the gradient wrt the gate output
        ds, dyg = self._ln_bwd(key + ".aan.ln", dsc, dx1, N, sub["ln"], bw + ".aln")
        dxg = ws.get(bw + ".dxg", (N, c.d))
        dy0 = ws.get(bw + ".dy0", (N, c.d))
        dz = ws.get(bw + ".dz", (N, 2 * c.d))
        ops.aan_gate_bwd(x, sub["y0"], sub["z"], dyg, dxg, dy0, dz)
        cat = sub["cat"]
        self._side(lambda: (ops.linear_wgrad(cat, dz, ps.g(key + ".aan.z.W")), ops.colsum(dz, ps.g(key + ".aan.z.b"))))
        dcat = ws.get(bw + ".dcat", (N, 2 * c.d))
        ops.linear_dgrad(dz, ps.w(key + ".aan.z.W"), dcat)
        ops.add2d(dy0, dcat[:, c.d:], dy0)          # y enters through the gate and through concat([x, y])
        if c.use_ffn:
            xf, h = sub["xf"], sub["h"]
            self._side(lambda: (ops.linear_wgrad(h, dy0, ps.g(key + ".aan.ffn.w2.W")),
                                ops.colsum(dy0, ps.g(key + ".aan.ffn.w2.b"))))
            dh = ws.get(bw + ".adh", (N, c.f))
            ops.linear_dgrad(dy0, ps.w(key + ".aan.ffn.w2.W"), dh, relu_mask=h,
                             alpha=1.0 / (1.0 - sub.get("relu_rate", 0.0)))
            self._side(lambda: (ops.linear_wgrad(xf, dh, ps.g(key + ".aan.ffn.w1.W")),
                                ops.colsum(dh, ps.g(key + ".aan.ffn.w1.b"))))
            dxf = ws.get(bw + ".dxf", (N, c.d))
            ops.linear_dgrad(dh, ps.w(key + ".aan.ffn.w1.W"), dxf)
        else:
            dxf = dy0
        dxpm = ws.get(bw + ".dxpm", (N, c.d))
        ops.prefix_mean_bwd(dxf.view(B, T, c.d), dxpm.view(B, T, c.d), tgt_len, mode=0 if c.aan_mask else 1)
        t1 = ws.get(bw + ".t1", (N, c.d))
        ops.add2d(ds, dxg, t1)
        t2 = ws.get(bw + ".t2", (N, c.d))
        ops.add2d(dcat[:, :c.d], dxpm, t2)
        return t1, t2
    # ---- merged attention
    kc = key + ".cross"
    dsc, dyc = self._ln_bwd(kc + ".ln", ds2, dxc, N, sub["ln"], bw + ".lnc", ps.g(kc + ".o.b"))
    osum = sub["osum"]
    self._side(lambda: ops.linear_wgrad(osum, dyc, ps.g(kc + ".o.W")))
    do = ws.get(bw + ".do", (N, c.d))
    ops.linear_dgrad(dyc, ps.w(kc + ".o.W"), do)
    dq = ws.get(bw + ".dq", (N, c.d))
    dkv = ws.get(bw + ".dkv", (B * S, 2 * c.d))
    dkv3 = dkv.view(B, S, 2 * c.d)
    delta = ws.get(bw + ".delta", (B, c.h, T), f32)
    ops.attention_bwd(sub["attn"], do.view(B, T, c.d), dq.view(B, T, c.d), dkv3[:, :, :c.d], dkv3[:, :, c.d:], delta,
                      workspace=self._attn_scratch)
    self._side(lambda: (ops.linear_wgrad(x, dq, ps.g(kc + ".q.W")), ops.colsum(dq, ps.g(kc + ".q.b")),
                        ops.linear_wgrad(enc, dkv, ps.g(kc + ".kv.W")), ops.colsum(dkv, ps.g(kc + ".kv.b"))))
    ops.linear_dgrad(dkv, ps.w(kc + ".kv.W"), d_enc_f32, accum=True)
    dxq = ws.get(bw + ".dxq", (N, c.d))
    ops.linear_dgrad(dq, ps.w(kc + ".q.W"), dxq)
    # the averaged branch: aan_o = prefix_mean(v_map(x)); v_map is the second half of the fused kv weights
    dvq = ws.get(bw + ".dvq", (N, c.d))
    ops.prefix_mean_bwd(do.view(B, T, c.d), dvq.view(B, T, c.d), tgt_len, mode=0)
    self._side(lambda: (ops.linear_wgrad(x, dvq, ps.g(kc + ".kv.W")[:, c.d:]),
                        ops.colsum(dvq, ps.g(kc + ".kv.b")[c.d:])))
    dxv = ws.get(bw + ".dxv", (N, c.d))
    ops.linear_dgrad(dvq, ps.w(kc + ".kv.W")[:, c.d:], dxv)
    ops.add2d(dxq, dxv, dxq)
    return dsc, dxq


# ------------------------------------------------------------------------------------------------ cached decode
def _decoding_fn_avg(self, target, state, time, candidates=None):
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
    fused_small = os.environ.get("ZB_DECODE_FUSED_SMALL", "1") != "0"
    for l in range(c.ndec):
        key = "dec%d" % l
        kc = key + ".cross"
        mem = state.mem[l]
        if c.aan and fused_small and not c.use_ffn:
            # opt-in (ZB_DECODE_FUSED_SMALL=1): 3 launches instead of 6 around the gate GEMM
            xf = ws.get("dec.xf", (R, c.d))
            cat = ws.get("dec.cat", (R, 2 * c.d))
            ops.aan_cat_step(x, state.sums[l], cat, xf, t)
            z = ws.get("dec.z", (R, 2 * c.d))
            ops.linear_fwd(cat, ps.w(key + ".aan.z.W"), ps.p(key + ".aan.z.b"), z)
            x1 = ws.get("dec.x1", (R, c.d))
            ops.aan_gate_ln(x, xf, z, x1, ps.p(key + ".aan.ln.scale"), ps.p(key + ".aan.ln.offset"), c.eps)
        elif c.aan:
            xf = ws.get("dec.xf", (R, c.d))
            ops.aan_step(x, state.sums[l], xf, t)
            cat = ws.get("dec.cat", (R, 2 * c.d))
            ops.add2d(x, None, cat[:, :c.d])
            if c.use_ffn:
                h = ws.get("dec.h", (R, c.f))
                ops.linear_fwd(xf, ps.w(key + ".aan.ffn.w1.W"), ps.p(key + ".aan.ffn.w1.b"), h, relu=True)
                ops.linear_fwd(h, ps.w(key + ".aan.ffn.w2.W"), ps.p(key + ".aan.ffn.w2.b"), cat[:, c.d:])
                y0 = ws.get("dec.y0", (R, c.d))
                ops.add2d(cat[:, c.d:], None, y0)      # contiguous copy of the FFN output for the gate kernel
            else:
                ops.add2d(xf, None, cat[:, c.d:])
                y0 = xf                                # without the FFN the averaged input IS y: no second copy
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
        xc = ws.get("dec.xc", (R, c.d))
        self._decode_proj_ln(ctx, kc + ".o", x1, xc, kc + ".ln", R)
        h = ws.get("dec.h", (R, c.f))
        ops.linear_fwd(xc, ps.w(key + ".ffn.w1.W"), ps.p(key + ".ffn.w1.b"), h, relu=True)
        self._decode_proj_ln(h, key + ".ffn.w2", xc, x, key + ".ffn.ln", R)
    return self._step_logits(x, R, candidates), state


# ---- DecodeState: running sums instead of K/V caches
_orig_begin = DecodeState.begin_search
_orig_reorder = DecodeState.reorder
_orig_swap = DecodeState.swap_buffers


def _begin_search(self, beam, cap=None):
    _orig_begin(self, beam, cap)
    c, eng = self.engine.cfg, self.engine
    if c.aan or c.fuse:
        R = self.B * self.K
        self.sums = [eng.ws.get("dec.sumA%d" % l, (R, c.d), f32) for l in range(c.ndec)]
        self.sums_alt = [eng.ws.get("dec.sumB%d" % l, (R, c.d), f32) for l in range(c.ndec)]
        for s in self.sums:
            s.zero_()


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
Engine._avg_layer_bwd = _avg_layer_bwd
Engine._decoding_fn_avg = _decoding_fn_avg
