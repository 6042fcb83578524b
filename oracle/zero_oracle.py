"""CPU oracle for Zero's Transformer hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Nothing under zero_b200/ imports it; the product path fails loudly when the CUDA library is missing.

What it is: an independent restatement (torch CPU, fp32 or fp64) of the arithmetic of
  func.py:14-104,164-400 · modules/rpr.py:10-75 · modules/rela.py:13-109 · models/transformer.py:15-285 ·
  models/transformer_aan.py:92-260 · models/transformer_fuse.py:120-165 · models/transformer_rpr.py ·
  models/transformer_rela.py · utils/util.py:88-103,198,274-287 · search.py:19-275
of bzhangGo/zero @ d97e2c2, written as pure functions over a {tf_variable_name: tensor} dict.

Pinning: the reference ships no tests or golden vectors and TF1.x cannot run here.  The oracle is pinned
instead against vectors produced by executing the reference's OWN unmodified Python over an eager TF1 shim
(oracle/tf1_shim, generator tests/golden/make_golden.py): tests/test_oracle_golden.py checks loss, every
gradient, scores, logits, per-step decode logits and beam-search sequences for transformer / aan / rpr /
rela / fuse.  Residual risk: the shim's restatement of the TF op semantics themselves.

`q` (a callable, default identity) is applied wherever the CUDA path stores a bf16 tensor, so the same code
doubles as a bf16-rounding emulation for tight kernel checks.
"""
from __future__ import annotations

import math

import numpy as np
import torch

F32_MIN = float(np.finfo(np.float32).min)


# ---- dropout hook (test infrastructure).  `DROP` is None (all dropout closed, like util.closing_dropout) or a
# callable drop(site_name, x) -> tf.nn.dropout(x) with an externally supplied mask; site names are the CUDA engine's
# ("enc0.self.att", "enc0.self.ln.res", "enc0.ffn.relu", "enc.emb", ...), so a test can hand the oracle exactly the
# masks the kernels generate (dropout_mask below restates the kernels' counter-based keep function).
DROP = None


def _drop(site, x):
    return x if DROP is None else DROP(site, x)


def dropout_keep(seed, site_name, numel, rate):
    """The CUDA path's keep mask (csrc/zb_ptx.cuh dropout_hash4 / dropout_mul) for flat indices [0, numel):
    one splitmix64 hash per 4 consecutive elements, 16 bits each, keep iff bits >= round(rate * 65536)."""
    import zlib
    site = np.uint64(zlib.crc32(site_name.encode("ascii")) & 0xFFFFFFFF)
    idx = np.arange(numel, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + site * np.uint64(0x9E3779B97F4A7C15) + (idx >> np.uint64(2)) * np.uint64(0xD1B54A32D192ED03)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
        bits = (z >> (np.uint64(16) * (idx & np.uint64(3)))) & np.uint64(0xFFFF)
    thr = min(max(int(np.float32(rate) * np.float32(65536.0) + np.float32(0.5)), 0), 65536)
    return bits >= np.uint64(thr)


def make_drop(seed, rates):
    """drop(site, x) for DROP: rates = {"emb", "att", "relu", "res"} -> rate; the kind is the site's suffix."""
    def drop(site, x):
        kind = site.rsplit(".", 1)[1]
        rate = float(rates.get(kind, 0.0))
        if rate <= 0.0:
            return x
        keep = torch.from_numpy(dropout_keep(seed, site, x.numel(), rate)).reshape(x.shape)
        return x * keep.to(x.dtype) / (1.0 - rate)
    return drop


def _ident(x):
    return x


def bf16_round(x):
    return x.to(torch.bfloat16).to(x.dtype) if x.dtype.is_floating_point else x


class Cfg(object):
    """The hyper-parameters the hot path reads (`Keys consumed by the hot path`)."""

    def __init__(self, hp, src_vocab, tgt_vocab):
        g = lambda k, d=None: getattr(hp, k) if (hasattr(hp, k) or (hasattr(hp, "__contains__") and k in hp)) else d  # noqa: E731
        self.model = str(g("model_name", "transformer")).lower()
        self.scope = g("scope_name") or "model"
        self.d = int(g("hidden_size"))
        self.e = int(g("embed_size", self.d))
        self.f = int(g("filter_size"))
        self.h = int(g("num_heads"))
        self.nenc = int(g("num_encoder_layer"))
        self.ndec = int(g("num_decoder_layer"))
        self.smooth = float(g("label_smooth", 0.1))
        self.share_st = bool(g("shared_source_target_embedding", False))
        self.share_ts = bool(g("shared_target_softmax_embedding", True))
        self.max_rel = int(g("max_relative_position", 16))
        self.aan_mask = bool(g("aan_mask", True))
        self.use_ffn = bool(g("use_ffn", False))
        self.eps = float(g("dtype_epsilon", 1e-8))
        self.inf = float(g("dtype_inf", 1e8))
        self.beam = int(g("beam_size", 4))
        self.alpha = float(g("decode_alpha", 0.6))
        self.decode_length = int(g("decode_length", 50))
        self.temperature = float(g("beam_search_temperature", 1.0))
        self.vs, self.vt = int(src_vocab), int(tgt_vocab)
        self.deep_init = bool(g("deep_transformer_init", False))
        self.init = g("initializer", "uniform_unit_scaling")
        self.init_gain = float(g("initializer_gain", 1.0))

    @property
    def rpr(self):
        return self.model == "transformer_rpr"

    @property
    def rela(self):
        return self.model == "transformer_rela"

    @property
    def aan(self):
        return self.model == "transformer_aan"

    @property
    def fuse(self):
        return self.model == "transformer_fuse"


# ------------------------------------------------------------------------------------------------ parameters
def param_shapes(c: Cfg):
    """Variable tree of SURVEY.md Appendix A (names exactly as TF builds them)."""
    s = c.scope
    out = {}
    if c.share_st:
        out[s + "/embedding"] = (c.vs, c.e)
    else:
        out[s + "/src_embedding"] = (c.vs, c.e)
        out[s + "/tgt_embedding"] = (c.vt, c.e)
        if not c.share_ts:
            out[s + "/softmax_embedding"] = (c.vt, c.e)
    out[s + "/bias"] = (c.e,)

    def lin(p, i, o):
        out[p + "/W_0_0"] = (i, o)
        out[p + "/b_0"] = (o,)

    def ln(p):
        out[p + "/layer_norm/scale"] = (c.d,)
        out[p + "/layer_norm/offset"] = (c.d,)

    def attn_extras(p, cross):
        if c.rpr:
            out[p + "/rpr_keys/embeddings"] = (2 * c.max_rel + 1, c.d // c.h)
            out[p + "/rpr_values/embeddings"] = (2 * c.max_rel + 1, c.d // c.h)
        if c.rela:
            out[p + "/post/scale"] = (c.d,)
            out[p + "/post/gate"] = (c.d,)

    def self_attn(p):
        lin(p + "/dot_attention/qkv_map", c.d, 3 * c.d)
        attn_extras(p + "/dot_attention", False)
        lin(p + "/dot_attention/o_map", c.d, c.d)
        ln(p)

    def cross_attn(p):
        for m in ("q_map", "k_map", "v_map"):
            lin(p + "/dot_attention/" + m, c.d, c.d)
        attn_extras(p + "/dot_attention", True)
        lin(p + "/dot_attention/o_map", c.d, c.d)
        ln(p)

    def ffn(p):
        lin(p + "/ffn_layer/enlarge", c.d, c.f)
        lin(p + "/ffn_layer/output", c.f, c.d)

    for l in range(c.nenc):
        p = "%s/encoder/layer_%d" % (s, l)
        self_attn(p + "/self_attention")
        ffn(p + "/feed_forward")
        ln(p + "/feed_forward")
    for l in range(c.ndec):
        p = "%s/decoder/layer_%d" % (s, l)
        if c.aan:
            a = p + "/average_attention"
            if c.use_ffn:
                ffn(a)
            lin(a + "/z_project", 2 * c.d, 2 * c.d)
            ln(a)
            cross_attn(p + "/cross_attention")
        elif c.fuse:
            cross_attn(p + "/fuse_attention")
        else:
            self_attn(p + "/self_attention")
            cross_attn(p + "/cross_attention")
        ffn(p + "/feed_forward")
        ln(p + "/feed_forward")
    return out


def init_params(c: Cfg, seed=1234, dtype=torch.float32):
    """Distribution-equivalent init (modules/initializer.py:11-32, models/transformer.py:18,38-45; App. C)."""
    g = torch.Generator().manual_seed(seed)
    out = {}

    def vs_uniform(shape, scale):
        fi, fo = (shape[0], shape[0]) if len(shape) == 1 else (shape[0], shape[1])
        lim = math.sqrt(3.0 * scale / ((fi + fo) / 2.0))
        return (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim

    for name, shape in param_shapes(c).items():
        leaf = name.rsplit("/", 1)[1]
        if leaf.endswith("embedding"):
            v = torch.randn(shape, generator=g, dtype=torch.float64) * c.d ** -0.5
        elif leaf == "b_0" or leaf == "offset":
            v = torch.zeros(shape, dtype=torch.float64)
        elif leaf == "scale":
            v = torch.ones(shape, dtype=torch.float64)
        else:
            scale = c.init_gain
            if c.deep_init and "/layer_" in name:
                layer = int(name.split("/layer_")[1].split("/")[0])
                scale = c.init_gain * (layer + 1) ** -0.5
            if c.init == "uniform" and not (c.deep_init and "/layer_" in name):
                v = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * c.init_gain
            elif c.init == "normal" and not (c.deep_init and "/layer_" in name):
                v = torch.randn(shape, generator=g, dtype=torch.float64) * c.init_gain
            else:
                v = vs_uniform(shape, scale)
        out[name] = v.to(dtype)
    return out


# ------------------------------------------------------------------------------------------------ primitives
def linear(P, prefix, x, q=_ident):
    """func.linear (func.py:14-65): x @ W_0_0 + b_0."""
    return q(x @ P[prefix + "/W_0_0"] + P[prefix + "/b_0"])


def layer_norm(P, prefix, x, eps, q=_ident):
    """func.layer_norm (func.py:289-303): biased variance, eps inside rsqrt."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return q(P[prefix + "/scale"] * (x - mu) * torch.rsqrt(var + eps) + P[prefix + "/offset"])


def timing_signal(length, channels, dtype, time=None):
    """func.add_timing_signal (func.py:341-369)."""
    pos = torch.arange(length, dtype=dtype) if time is None else torch.tensor([float(time)], dtype=dtype)
    nts = channels // 2
    inc = math.log(1.0e4 / 1.0) / (float(nts) - 1)
    inv = torch.exp(torch.arange(nts, dtype=dtype) * -inc)
    st = pos[:, None] * inv[None, :]
    sig = torch.cat([torch.sin(st), torch.cos(st)], 1)
    if channels % 2:
        sig = torch.nn.functional.pad(sig, (0, 1))
    return sig.reshape(1, -1, channels)


def heads_split(x, h):
    b, l, d = x.shape
    return x.reshape(b, l, h, d // h).permute(0, 2, 1, 3)


def heads_merge(x):
    b, h, l, dh = x.shape
    return x.permute(0, 2, 1, 3).reshape(b, l, h * dh)


def rel_index(lq, lk, k, q_offset=0):
    """modules/rpr.py:62-75: clip(i - j, -k, k) + k with i the (absolute) query position."""
    i = torch.arange(lq)[:, None] + q_offset
    j = torch.arange(lk)[None, :]
    return torch.clamp(i - j, -k, k) + k


def attention_core(c, P, prefix, qh, kh, vh, bias, q_offset=0, q=_ident, site=None):
    """func.dot_attention core (func.py:218-256) / rela variant (modules/rela.py:52-75).
    qh,kh,vh: [B,h,L,dh]; bias: additive, broadcastable to [B,h,Lq,Lk] (or None)."""
    dh = qh.shape[-1]
    qh = qh * dh ** -0.5
    logits = qh @ kh.transpose(-1, -2)
    lq, lk = logits.shape[-2], logits.shape[-1]
    if c.rpr:
        idx = rel_index(lq, lk, c.max_rel, q_offset)
        ek = P[prefix + "/rpr_keys/embeddings"][idx]  # [lq, lk, dh]
        logits = logits + torch.einsum("bhid,ijd->bhij", qh, ek)
    if c.rela:
        if bias is not None:
            logits = logits * (bias == 0).to(logits.dtype)
        w = torch.relu(logits)
    else:
        if bias is not None:
            logits = logits + bias
        w = torch.softmax(logits, -1)
    if site is not None:
        w = _drop(site + ".att", w)   # func.py:245 / modules/rela.py:74: the dropped weights feed both value terms
    o = w @ vh
    if c.rpr:
        ev = P[prefix + "/rpr_values/embeddings"][idx]
        o = o + torch.einsum("bhij,ijd->bhid", w, ev)
    o = q(heads_merge(o))
    if c.rela:
        x = o
        ms = (x ** 2).mean(-1, keepdim=True)
        o = q(P[prefix + "/post/scale"] * x * torch.rsqrt(ms + c.eps) * torch.sigmoid(P[prefix + "/post/gate"] * x))
    return o, w


def mask_bias(mask, inf):
    """func.attention_bias 'masking' (func.py:385-388)."""
    return ((1.0 - mask) * -inf)[:, None, None, :]


def causal_bias(n, inf, dtype):
    """func.attention_bias 'causal' (func.py:379-384)."""
    return (-inf * (1.0 - torch.tril(torch.ones(n, n, dtype=dtype))))[None, None]


def aan_matrix(mask, inf):
    """func.attention_bias 'aan' (func.py:389-398), softmax trick included."""
    b, t = mask.shape
    cum = torch.tril(torch.ones(t, t, dtype=mask.dtype))[None]
    m = mask[:, None, :] * mask[:, :, None] * cum
    w = torch.softmax(m + (1.0 - m) * -inf, -1)
    return w * m


def self_attention(c, P, p, x, bias, cache=None, q_offset=0, q=_ident, site=None):
    a = p + "/dot_attention"
    qkv = linear(P, a + "/qkv_map", x, q)
    qq, kk, vv = torch.split(qkv, c.d, -1)
    if cache is not None:
        kk = torch.cat([cache["k"], kk], 1)
        vv = torch.cat([cache["v"], vv], 1)
        cache = dict(cache, k=kk, v=vv)
    o, _ = attention_core(c, P, a, heads_split(qq, c.h), heads_split(kk, c.h), heads_split(vv, c.h), bias,
                          q_offset, q, site)
    return linear(P, a + "/o_map", o, q), cache


def cross_attention(c, P, p, x, memory, bias, cache=None, q_offset=0, fuse=None, q=_ident, site=None):
    """func.dot_attention with memory (func.py:206-216) and the merged-attention branch (func.py:258-275)."""
    a = p + "/dot_attention"
    qq = linear(P, a + "/q_map", x, q)
    if cache is not None and "mk" in cache:
        kk, vv = cache["mk"], cache["mv"]
    else:
        kk = linear(P, a + "/k_map", memory, q)
        vv = linear(P, a + "/v_map", memory, q)
    if cache is not None:
        cache = dict(cache, mk=kk, mv=vv)
    o, _ = attention_core(c, P, a, heads_split(qq, c.h), heads_split(kk, c.h), heads_split(vv, c.h), bias,
                          q_offset, q, site)
    if fuse is not None:
        vq = linear(P, a + "/v_map", x, q)  # query projected with the cross-attention v_map (func.py:260)
        if cache is not None and "aan" in cache:
            aan_o = (vq + cache["aan"]) / float(fuse + 1)
            cache = dict(cache, aan=vq + cache["aan"])
        else:
            aan_o = fuse @ vq
            if cache is not None:
                cache = dict(cache, aan=vq)
        o = q(o + aan_o)
    return linear(P, a + "/o_map", o, q), cache


def ffn(c, P, p, x, q=_ident, site=None):
    """func.ffn_layer (func.py:327-338)."""
    hdn = q(torch.relu(x @ P[p + "/ffn_layer/enlarge/W_0_0"] + P[p + "/ffn_layer/enlarge/b_0"]))
    if site is not None:
        hdn = _drop(site + ".relu", hdn)
    return linear(P, p + "/ffn_layer/output", hdn, q)


def remove_invalid_seq(seq, mask):
    """utils/util.py:274-287: drop all-pad columns, always keep column 0."""
    keep = mask.sum(0)
    keep[0] = keep[0] + 1
    keep = keep > 0
    return seq[:, keep], mask[:, keep]


# ------------------------------------------------------------------------------------------------ model
def encoder(c, P, source, dtype=torch.float32, q=_ident):
    """models/transformer.py:15-84."""
    s = c.scope
    mask = (source != 0).to(dtype)
    source, mask = remove_invalid_seq(source, mask)
    emb = P[s + ("/embedding" if c.share_st else "/src_embedding")]
    x = emb[source] * c.d ** 0.5 + P[s + "/bias"]
    x = q(x + timing_signal(x.shape[1], x.shape[2], dtype))
    x = _drop("enc.emb", x)
    bias = mask_bias(mask, c.inf)
    for l in range(c.nenc):
        p = "%s/encoder/layer_%d" % (s, l)
        k = "enc%d" % l
        y, _ = self_attention(c, P, p + "/self_attention", x, bias, q=q, site=k + ".self")
        x = layer_norm(P, p + "/self_attention/layer_norm", x + _drop(k + ".self.ln.res", y), c.eps, q)
        y = ffn(c, P, p + "/feed_forward", x, q, site=k + ".ffn")
        x = layer_norm(P, p + "/feed_forward/layer_norm", x + _drop(k + ".ffn.ln.res", y), c.eps, q)
    return {"encodes": x, "mask": mask}


def init_decode_state(c, enc, dtype=torch.float32):
    b = enc["encodes"].shape[0]
    layers = {}
    for l in range(c.ndec):
        if c.aan or c.fuse:
            layers["layer_%d" % l] = {"aan": torch.zeros(b, 1, c.d, dtype=dtype)}
        else:
            layers["layer_%d" % l] = {"k": torch.zeros(b, 0, c.d, dtype=dtype), "v": torch.zeros(b, 0, c.d, dtype=dtype)}
    return {"encodes": enc["encodes"], "mask": enc["mask"], "decoder": {"state": layers}}


def decoder(c, P, target, state, time=None, smooth=None, dtype=torch.float32, q=_ident):
    """models/transformer.py:87-218 (+aan :120-260, fuse :120-165).  Training when `time is None`.
    Returns loss, logits [N,V] fp32-equivalent, new state, per-sample loss."""
    s = c.scope
    training = time is None
    mask = (target != 0).to(dtype)
    if training:
        target, mask = remove_invalid_seq(target, mask)
    emb = P[s + ("/embedding" if c.share_st else "/tgt_embedding")]
    x = emb[target] * c.d ** 0.5 + P[s + "/bias"]
    if training:
        x = torch.nn.functional.pad(x, (0, 0, 1, 0))[:, :-1]
        x = q(x + timing_signal(x.shape[1], x.shape[2], dtype))
        x = _drop("dec.emb", x)
    else:
        if bool((target == 0).all()):
            x = torch.zeros_like(x)
        mask = torch.ones_like(mask)
        x = q(x + timing_signal(1, x.shape[2], dtype, time=time))
    t = x.shape[1]
    q_off = 0 if training else int(time)
    cbias = causal_bias(t, c.inf, dtype)
    mbias = mask_bias(state["mask"], c.inf)
    new_layers = {}
    for l in range(c.ndec):
        p = "%s/decoder/layer_%d" % (s, l)
        cache = None if training else dict(state["decoder"]["state"]["layer_%d" % l])
        k = "dec%d" % l
        if c.aan:
            a = p + "/average_attention"
            if training:
                if c.aan_mask:
                    xf = aan_matrix(mask, c.inf) @ x
                else:
                    cnt = torch.cumsum(mask, 1)
                    cnt = torch.where(cnt <= 0, torch.ones_like(cnt), cnt)[:, :, None]
                    xf = torch.cumsum(x, 1) / cnt
            else:
                xf = (x + cache["aan"]) / float(time + 1)
                cache["aan"] = x + cache["aan"]
            xf = q(xf)
            y = ffn(c, P, a, xf, q, site=k + ".aan") if c.use_ffn else xf
            z = linear(P, a + "/z_project", torch.cat([x, y], -1), q)
            gi, gf = torch.split(z, c.d, -1)
            y = q(torch.sigmoid(gi) * x + torch.sigmoid(gf) * y)
            x = layer_norm(P, a + "/layer_norm", x + _drop(k + ".aan.ln.res", y), c.eps, q)
            y, cache = cross_attention(c, P, p + "/cross_attention", x, state["encodes"], mbias, cache, q_off, q=q,
                                       site=k + ".cross")
            x = layer_norm(P, p + "/cross_attention/layer_norm", x + _drop(k + ".cross.ln.res", y), c.eps, q)
        elif c.fuse:
            fuse = aan_matrix(mask, c.inf) if training else time
            y, cache = cross_attention(c, P, p + "/fuse_attention", x, state["encodes"], mbias, cache, q_off,
                                       fuse=fuse, q=q, site=k + ".cross")
            x = layer_norm(P, p + "/fuse_attention/layer_norm", x + _drop(k + ".cross.ln.res", y), c.eps, q)
        else:
            # decoder self-attention: causal bias only, no key-padding mask (models/transformer.py:136)
            y, cache = self_attention(c, P, p + "/self_attention", x, cbias, cache, q_off, q, site=k + ".self")
            x = layer_norm(P, p + "/self_attention/layer_norm", x + _drop(k + ".self.ln.res", y), c.eps, q)
            y, cache = cross_attention(c, P, p + "/cross_attention", x, state["encodes"], mbias, cache, q_off, q=q,
                                       site=k + ".cross")
            x = layer_norm(P, p + "/cross_attention/layer_norm", x + _drop(k + ".cross.ln.res", y), c.eps, q)
        y = ffn(c, P, p + "/feed_forward", x, q, site=k + ".ffn")
        x = layer_norm(P, p + "/feed_forward/layer_norm", x + _drop(k + ".ffn.ln.res", y), c.eps, q)
        if not training:
            new_layers["layer_%d" % l] = cache
    feat = x.reshape(-1, c.e)
    if c.share_st:
        sm = P[s + "/embedding"]
    else:
        sm = P[s + ("/tgt_embedding" if c.share_ts else "/softmax_embedding")]
    logits = (feat @ sm.t()).float() if dtype == torch.float32 else feat @ sm.t()
    smooth = c.smooth if smooth is None else smooth
    ce = smoothed_ce(logits, target.reshape(-1), smooth).reshape(target.shape)
    m = mask.to(ce.dtype)
    per_sample = (ce * m).sum(-1) / m.sum(-1)
    loss = per_sample.mean() if target.shape[0] > 0 else per_sample.sum() * 0
    new_state = state if training else dict(state, decoder={"state": new_layers})
    return loss, logits, new_state, per_sample


def smoothed_ce(logits, labels, factor):
    """util.label_smooth + softmax_cross_entropy_with_logits_v2 - normaliser (utils/util.py:88-103,
    models/transformer.py:198-205)."""
    v = logits.shape[-1]
    lsm = torch.log_softmax(logits, -1)
    gold = lsm.gather(-1, labels.reshape(-1, 1).long()).squeeze(-1)
    if 0.0 < factor < 1.0:
        n = float(v - 1)
        p, qv = 1.0 - factor, factor / n
        norm = -(p * math.log(p) + n * qv * math.log(qv + 1e-20))
        return -(p * gold + qv * (lsm.sum(-1) - gold)) - norm
    return -gold


def train_loss(c, P, source, target, dtype=torch.float32, q=_ident):
    """train_fn (models/transformer.py:221-232)."""
    enc = encoder(c, P, source, dtype, q)
    loss, logits, _, per_sample = decoder(c, P, target, enc, None, None, dtype, q)
    return loss, logits, per_sample, enc


def score(c, P, source, target, dtype=torch.float32, q=_ident):
    """score_fn (models/transformer.py:235-249): dropout and label smoothing off."""
    enc = encoder(c, P, source, dtype, q)
    return decoder(c, P, target, enc, None, 0.0, dtype, q)[3]


def make_infer_fns(c, P, dtype=torch.float32, q=_ident):
    """infer_fn in 'cache' search mode (models/transformer.py:252-285)."""

    def encoding_fn(source):
        return init_decode_state(c, encoder(c, P, source, dtype, q), dtype)

    def decoding_fn(target, state, time):
        _, logits, new_state, _ = decoder(c, P, target, state, int(time), None, dtype, q)
        return logits, new_state

    return encoding_fn, decoding_fn


# ------------------------------------------------------------------------------------------------ beam search
def top_k(x, k):
    """tf.nn.top_k: descending, ties -> lower index."""
    v, i = torch.sort(x, dim=-1, descending=True, stable=True)
    return v[..., :k], i[..., :k]


def _map_state(fn, s):
    if isinstance(s, dict):
        return {k: _map_state(fn, v) for k, v in s.items()}
    return fn(s)


def beam_search(c, source, encoding_fn, decoding_fn, eos=2, pad=0, logits_hook=None):
    """search.beam_search (search.py:19-275), 'cache' mode, restated per SURVEY.md Appendix B.  fp32 scores."""
    K, alpha = c.beam, c.alpha
    B = source.shape[0]
    state = encoding_fn(source)
    src_len = (source != 0).float().sum(-1)
    max_len = src_len + c.decode_length
    max_len_i = max_len.to(torch.int64)
    tile = lambda x: x[:, None].expand(B, K, *x.shape[1:]).reshape(B * K, *x.shape[1:])  # noqa: E731
    state = _map_state(tile, state)
    # cache_init (search.py:56-77): dummy step at time 0; keys that already existed keep their old value,
    # so only the newly created ones (mk / mv) survive.
    _, dummy = decoding_fn(torch.full((B * K, 1), pad, dtype=torch.int64), state, 0)
    for l, cache in dummy["decoder"]["state"].items():
        for key in ("mk", "mv"):
            if key in cache:
                state["decoder"]["state"][l][key] = cache[key]

    seq = torch.full((B, K, 1), pad, dtype=torch.int64)
    logp = torch.tensor([[0.0] + [F32_MIN] * (K - 1)], dtype=torch.float32).repeat(B, 1)
    score = torch.zeros_like(logp)
    fin_seq = torch.zeros_like(seq)
    fin_score = torch.full((B, K), F32_MIN, dtype=torch.float32)
    fin_flag = torch.zeros((B, K), dtype=torch.bool)
    MIN = torch.tensor(F32_MIN, dtype=torch.float32)
    bidx = torch.arange(B)[:, None]
    t = 0
    while True:
        # _not_finished (search.py:85-113)
        max_pen = torch.pow((5.0 + max_len.float()) / 6.0, alpha)
        best_alive = logp[:, 0] / max_pen
        worst_fin = (fin_score * fin_flag.float()).min(1).values
        worst_fin = worst_fin + (1.0 - fin_flag.any(1).float()) * MIN
        bound_met = bool((worst_fin > best_alive).all())
        length_ok = bool((t < max_len_i).any())
        if bound_met or not length_ok:
            break
        logits, new_state = decoding_fn(seq.reshape(B * K, -1)[:, -1:], state, t)
        logits = logits.float()
        if logits_hook is not None:
            logits_hook(t, logits)
        logits = logits / c.temperature
        lp = logits - torch.logsumexp(logits, -1, keepdim=True)
        V = lp.shape[-1]
        if t < 1:
            eos_mask = (torch.arange(V) == eos).float()
            lp = lp + eos_mask[None, :] * -c.inf
        lp = lp.reshape(B, K, V)
        cand = logp[:, :, None] + lp
        pen = torch.pow(torch.tensor((5.0 + float(t + 1)) / 6.0, dtype=torch.float32), alpha)
        cs = cand / pen
        top_s, top_i = top_k(cs.reshape(B, K * V), 2 * K)
        bi = top_i // V
        wi = top_i % V
        cseq = torch.cat([seq[bidx, bi], wi[:, :, None]], 2)
        done = (wi == eos) | (t >= max_len_i)[:, None]
        a_s, a_i = top_k(top_s + done.float() * MIN, K)
        seq = cseq[bidx, a_i]
        parent = bi[bidx, a_i]
        flat_parent = (parent + torch.arange(B)[:, None] * K).reshape(-1)
        state = _map_state(lambda x: x[flat_parent], new_state)
        logp = a_s * pen
        score = a_s
        f_all = torch.cat([fin_score, top_s + (1.0 - done.float()) * MIN], 1)
        flag_all = torch.cat([fin_flag, done], 1)
        fin_score, f_i = top_k(f_all, K)
        fin_flag = flag_all[bidx, f_i]
        fin_seq = torch.cat([torch.cat([fin_seq, torch.full((B, K, 1), pad, dtype=torch.int64)], 2), cseq], 1)[bidx, f_i]
        t += 1
    any_fin = fin_flag.any(1)
    out_seq = torch.where(any_fin[:, None, None], fin_seq, seq)[:, :, 1:]
    out_score = torch.where(any_fin[:, None], fin_score, score)
    return {"seq": out_seq, "score": out_score, "steps": t}


# ------------------------------------------------------------------------------------------------ optimizer
def adam_tf_step(p, m, v, g, step, lr, b1, b2, eps):
    """tf.train.AdamOptimizer update: epsilon outside the bias correction."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    return p - lr_t * m / (torch.sqrt(v) + eps), m, v
