"""Execution engine: Zero's Transformer family (models/transformer*.py) as an explicit forward / backward
schedule of C-ABI kernel calls over pre-planned device buffers.

The reference builds a TF1 graph and lets `optimizer.compute_gradients` derive the backward pass
(main.py:22-45).  Here both directions are written out: every activation the backward needs is kept in a
per-shape workspace (static addresses -> the whole step is CUDA-graph capturable), every parameter lives in
one flat fp32 master arena with a bf16 compute mirror (utils/dtype.py:55-69), every parameter gradient is
accumulated in one flat fp32 arena (-> a single NCCL all-reduce, utils/parallel.py:134-208).

Parameter names are the reference's TF variable names; `load_state_dict` /
`state_dict` speak that vocabulary.  Cross-attention k_map / v_map weights are stored fused ([d, 2d]) so memory
is projected by one GEMM; their TF names are strided views into the fused tensor.
"""
from __future__ import annotations

import contextlib
import math
import os
import zlib
from collections import OrderedDict

import torch

from . import lib as L
from . import ops

bf16 = torch.bfloat16
f32 = torch.float32


def _hp(hp, key, default=None):
    try:
        return getattr(hp, key)
    except AttributeError:
        return default


class ModelConfig(object):
    """Hot-path hyper-parameters."""

    def __init__(self, hp, src_vocab=None, tgt_vocab=None):
        self.model = str(_hp(hp, "model_name", "transformer")).lower()
        self.scope = _hp(hp, "scope_name") or "model"
        self.d = int(hp.hidden_size)
        self.e = int(_hp(hp, "embed_size", self.d))
        self.f = int(hp.filter_size)
        self.h = int(hp.num_heads)
        self.nenc = int(hp.num_encoder_layer)
        self.ndec = int(hp.num_decoder_layer)
        self.smooth = float(_hp(hp, "label_smooth", 0.1))
        self.share_st = bool(_hp(hp, "shared_source_target_embedding", False))
        self.share_ts = bool(_hp(hp, "shared_target_softmax_embedding", True))
        self.max_rel = int(_hp(hp, "max_relative_position", 16))
        self.aan_mask = bool(_hp(hp, "aan_mask", True))
        self.use_ffn = bool(_hp(hp, "use_ffn", False))
        self.eps = float(_hp(hp, "dtype_epsilon", 1e-8))
        self.inf = float(_hp(hp, "dtype_inf", 1e8))
        self.loss_scale = float(_hp(hp, "loss_scale", 1.0))
        self.deep_init = bool(_hp(hp, "deep_transformer_init", False))
        self.init = _hp(hp, "initializer", "uniform_unit_scaling")
        self.init_gain = float(_hp(hp, "initializer_gain", 1.0))
        sv = src_vocab if src_vocab is not None else _hp(hp, "src_vocab")
        tv = tgt_vocab if tgt_vocab is not None else _hp(hp, "tgt_vocab")
        self.vs = int(sv.size() if hasattr(sv, "size") else sv)
        self.vt = int(tv.size() if hasattr(tv, "size") else tv)
        if self.e != self.d:
            raise L.ZeroB200Error("embed_size must equal hidden_size on the Transformer path")
        if self.d % self.h or (self.d // self.h) not in (16, 32, 64):
            raise L.ZeroB200Error("head size %d unsupported (16/32/64)" % (self.d // max(self.h, 1)))
        if self.d % 8 or self.f % 8:
            raise L.ZeroB200Error("hidden/filter sizes must be multiples of 8 (16-byte TMA pitch)")
        # any vocabulary size (real vocabularies are 3 specials + N words): [rows, V] logits / d_logits rows are laid
        # out with a pitch rounded up to 8 elements, the kernels see n = V columns of it
        self.vt_pitch = (self.vt + 7) // 8 * 8
        known = ("transformer", "transformer_aan", "transformer_rpr", "transformer_rela", "transformer_fuse")
        if self.model not in known:
            raise L.ZeroB200Error("model %r is outside the hot path (supported: %s)" % (self.model, ", ".join(known)))
        # the k_map | v_map weights of ALL decoder layers live side by side in one [d, ndec * 2d] matrix, so the memory
        # projections of a training step are ONE GEMM forward (n = ndec * 2d) and ONE dgrad / wgrad pair backward
        # (k = ndec * 2d) instead of ndec small ones each: +1.3 % tokens/s at configs[1] (profiles/r02a_summary.txt);
        # ZB_BATCH_MEM_PROJ=0 restores one projection per layer
        self.batch_mem = os.environ.get("ZB_BATCH_MEM_PROJ", "1") != "0" and self.model not in (
            "transformer_aan", "transformer_fuse")

    rpr = property(lambda s: s.model == "transformer_rpr")
    rela = property(lambda s: s.model == "transformer_rela")
    aan = property(lambda s: s.model == "transformer_aan")
    fuse = property(lambda s: s.model == "transformer_fuse")


# ------------------------------------------------------------------------------------------------ parameters
class ParamStore(object):
    """Flat fp32 master / bf16 mirror / fp32 gradient / Adam m,v arenas + TF-named views."""

    ALIGN = 64  # elements; keeps every bf16 view 128-byte aligned

    def __init__(self, cfg: ModelConfig, device):
        self.cfg = cfg
        self.device = device
        self.slots = OrderedDict()     # engine tensor name -> (offset, shape)
        self.alias = {}                # engine tensor name -> (parent slot, slicer): a strided window of another slot
        self._wc, self._pc, self._gc = {}, {}, {}   # name -> (arena, view) of the compute copy / master / gradient
        self.tf_views = OrderedDict()  # TF variable name -> (engine name, slicer)
        self._plan()
        n = self.total
        self.master = torch.zeros(n, dtype=f32, device=device)
        self.mirror = torch.zeros(n, dtype=bf16, device=device)
        self.grad = torch.zeros(n, dtype=f32, device=device)
        self.adam_m = None
        self.adam_v = None

    # -- layout -------------------------------------------------------------------------------------
    def _add(self, name, shape, tf_name=None):
        size = 1
        for s in shape:
            size *= s
        off = getattr(self, "total", 0)
        self.slots[name] = (off, tuple(shape))
        self.total = (off + size + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if tf_name is not None:
            self.tf_views[tf_name] = (name, None)

    def _plan(self):
        c = self.cfg
        s = c.scope
        self.total = 0
        if c.share_st:
            self._add("src_emb", (c.vs, c.e), s + "/embedding")
            self.tf_alias_tgt = "src_emb"
        else:
            self._add("src_emb", (c.vs, c.e), s + "/src_embedding")
        self._add("emb_bias", (c.e,), s + "/bias")
        dh = c.d // c.h
        nb = 2 * c.max_rel + 1

        def lin(key, tfp, i, o):
            self._add(key + ".W", (i, o), tfp + "/W_0_0")
            self._add(key + ".b", (o,), tfp + "/b_0")

        def ln(key, tfp):
            self._add(key + ".scale", (c.d,), tfp + "/layer_norm/scale")
            self._add(key + ".offset", (c.d,), tfp + "/layer_norm/offset")

        def extras(key, tfp):
            if c.rpr:
                self._add(key + ".rpr_k", (nb, dh), tfp + "/rpr_keys/embeddings")
                self._add(key + ".rpr_v", (nb, dh), tfp + "/rpr_values/embeddings")
            if c.rela:
                self._add(key + ".post_scale", (c.d,), tfp + "/post/scale")
                self._add(key + ".post_gate", (c.d,), tfp + "/post/gate")

        def self_attn(key, tfp):
            lin(key + ".qkv", tfp + "/dot_attention/qkv_map", c.d, 3 * c.d)
            extras(key, tfp + "/dot_attention")
            lin(key + ".o", tfp + "/dot_attention/o_map", c.d, c.d)
            ln(key + ".ln", tfp)

        def cross_attn(key, tfp, l=0):
            lin(key + ".q", tfp + "/dot_attention/q_map", c.d, c.d)
            # fused k_map | v_map storage; TF names are column slices
            if c.batch_mem:
                self.alias[key + ".kv.W"] = ("dec.kvall.W", (slice(None), slice(l * 2 * c.d, (l + 1) * 2 * c.d)))
                self.alias[key + ".kv.b"] = ("dec.kvall.b", (slice(l * 2 * c.d, (l + 1) * 2 * c.d),))
            else:
                self._add(key + ".kv.W", (c.d, 2 * c.d))
                self._add(key + ".kv.b", (2 * c.d,))
            self.tf_views[tfp + "/dot_attention/k_map/W_0_0"] = (key + ".kv.W", (slice(None), slice(0, c.d)))
            self.tf_views[tfp + "/dot_attention/k_map/b_0"] = (key + ".kv.b", (slice(0, c.d),))
            self.tf_views[tfp + "/dot_attention/v_map/W_0_0"] = (key + ".kv.W", (slice(None), slice(c.d, 2 * c.d)))
            self.tf_views[tfp + "/dot_attention/v_map/b_0"] = (key + ".kv.b", (slice(c.d, 2 * c.d),))
            extras(key, tfp + "/dot_attention")
            lin(key + ".o", tfp + "/dot_attention/o_map", c.d, c.d)
            ln(key + ".ln", tfp)

        def ffn(key, tfp):
            lin(key + ".w1", tfp + "/ffn_layer/enlarge", c.d, c.f)
            lin(key + ".w2", tfp + "/ffn_layer/output", c.f, c.d)

        self.enc_layer_offset = []     # arena offset of every encoder layer's first parameter (all-reduce buckets)
        for l in range(c.nenc):
            key, tfp = "enc%d" % l, "%s/encoder/layer_%d" % (s, l)
            self.enc_layer_offset.append(self.total)
            self_attn(key + ".self", tfp + "/self_attention")
            ffn(key + ".ffn", tfp + "/feed_forward")
            ln(key + ".ffn.ln", tfp + "/feed_forward")
        # Everything from here on receives its last gradient contribution during the DECODER backward, which
        # finishes before the encoder backward starts: [dec_offset, total) is the first all-reduce bucket and
        # overlaps with the encoder backward (train.py).
        self.dec_offset = self.total
        if not c.share_st:
            self._add("tgt_emb", (c.vt, c.e), s + "/tgt_embedding")
            if not c.share_ts:
                self._add("softmax_emb", (c.vt, c.e), s + "/softmax_embedding")
        if c.batch_mem:
            self._add("dec.kvall.W", (c.d, c.ndec * 2 * c.d))
            self._add("dec.kvall.b", (c.ndec * 2 * c.d,))
        for l in range(c.ndec):
            key, tfp = "dec%d" % l, "%s/decoder/layer_%d" % (s, l)
            if c.aan:
                a = tfp + "/average_attention"
                if c.use_ffn:
                    ffn(key + ".aan.ffn", a)
                lin(key + ".aan.z", a + "/z_project", 2 * c.d, 2 * c.d)
                ln(key + ".aan.ln", a)
                cross_attn(key + ".cross", tfp + "/cross_attention")
            elif c.fuse:
                cross_attn(key + ".cross", tfp + "/fuse_attention")
            else:
                self_attn(key + ".self", tfp + "/self_attention")
                cross_attn(key + ".cross", tfp + "/cross_attention", l)
            ffn(key + ".ffn", tfp + "/feed_forward")
            ln(key + ".ffn.ln", tfp + "/feed_forward")

    # -- views ---------------------------------------------------------------------------------------
    def _view(self, arena, name):
        if name in self.alias:
            parent, sl = self.alias[name]
            return self._view(arena, parent)[sl]
        off, shape = self.slots[name]
        n = 1
        for s in shape:
            n *= s
        return arena[off:off + n].view(shape)

    # The step asks for ~500 parameter views; building one (slice + view) costs a few microseconds of Python, which adds
    # up to milliseconds per eager step, so w / p / g remember theirs.  An entry is tied to the arena OBJECT it was
    # taken from: rebinding an arena (the sharded optimizer step moves them into symmetric memory) refreshes it.
    def _cached(self, cache, arena, name):
        hit = cache.get(name)
        if hit is None or hit[0] is not arena:
            hit = cache[name] = (arena, self._view(arena, name))
        return hit[1]

    def w(self, name):      # bf16 compute copy
        return self._cached(self._wc, self.mirror, name)

    def p(self, name):      # fp32 master
        return self._cached(self._pc, self.master, name)

    def g(self, name):      # fp32 gradient
        return self._cached(self._gc, self.grad, name)

    def tf_names(self):
        return list(self.tf_views.keys())

    def tf_view(self, arena, tf_name):
        eng, sl = self.tf_views[tf_name]
        v = self._view(arena, eng)
        return v if sl is None else v[sl]

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self.tf_views if k not in sd]
        extra = [k for k in sd if k not in self.tf_views]
        if strict and (missing or extra):
            raise L.ZeroB200Error("state dict mismatch: missing %s, unexpected %s" % (missing[:4], extra[:4]))
        for k, (eng, sl) in self.tf_views.items():
            if k in sd:
                self.tf_view(self.master, k).copy_(torch.as_tensor(sd[k]).to(self.device, f32))
        self.refresh_mirror()

    def state_dict(self):
        return OrderedDict((k, self.tf_view(self.master, k).detach().clone().cpu()) for k in self.tf_views)

    def grad_dict(self):
        return OrderedDict((k, self.tf_view(self.grad, k).detach().clone().cpu()) for k in self.tf_views)

    def refresh_mirror(self):
        ops.cast_f32_bf16(self.master, self.mirror)

    def zero_grad(self):
        self.grad.zero_()

    def init_random(self, seed=1234):
        """Distribution-equivalent init (modules/initializer.py:11-32; models/transformer.py:18,38-45)."""
        c = self.cfg
        g = torch.Generator(device="cpu").manual_seed(seed)
        for k in self.tf_views:
            v = self.tf_view(self.master, k)
            shape = tuple(v.shape)
            leaf = k.rsplit("/", 1)[1]
            if leaf.endswith("embedding"):
                t = torch.randn(shape, generator=g) * c.d ** -0.5
            elif leaf in ("b_0", "offset"):
                t = torch.zeros(shape)
            elif leaf == "scale":
                t = torch.ones(shape)
            else:
                scale = c.init_gain
                layered = "/layer_" in k
                if c.deep_init and layered:
                    scale = c.init_gain * (int(k.split("/layer_")[1].split("/")[0]) + 1) ** -0.5
                scoped = not (c.deep_init and layered)     # a layer's own variance-scaling initializer wins
                fi, fo = (shape[0], shape[0]) if len(shape) == 1 else (shape[0], shape[1])
                if c.init == "uniform" and scoped:
                    t = (torch.rand(shape, generator=g) * 2 - 1) * c.init_gain
                elif c.init == "normal" and scoped:
                    t = torch.randn(shape, generator=g) * c.init_gain
                elif c.init == "normal_unit_scaling" and scoped:
                    # tf.variance_scaling_initializer(distribution="normal"): a normal truncated at two standard
                    # deviations, its stddev corrected by 0.8796... so that the truncated variance is scale / fan_avg
                    std = math.sqrt(scale / ((fi + fo) / 2.0)) / 0.87962566103423978
                    t = torch.randn(shape, generator=g)
                    for _ in range(64):
                        bad = t.abs() > 2.0
                        if not bool(bad.any()):
                            break
                        t = torch.where(bad, torch.randn(shape, generator=g), t)
                    t = t.clamp_(-2.0, 2.0) * std
                else:
                    if scoped and c.init != "uniform_unit_scaling":
                        scale = 1.0                        # unrecognised name: glorot_uniform (initializer.py:29-32)
                    lim = math.sqrt(3.0 * scale / ((fi + fo) / 2.0))
                    t = (torch.rand(shape, generator=g) * 2 - 1) * lim
            v.copy_(t.to(self.device))
        self.refresh_mirror()


# ------------------------------------------------------------------------------------------------ workspace
class Workspace(object):
    """Named device buffers.  One flat allocation per (name, dtype), sized for the largest request seen so far; `get`
    returns its leading `prod(shape)` elements viewed as `shape`.  With one batch shape (the benchmark, a captured CUDA
    graph) that is one allocation per name at a static address, exactly as large as needed; with the reference's
    token-budget batches — a new (B, S, T) almost every step — the footprint stays that of the largest batch instead of
    growing with every distinct shape.  Contract: a buffer's contents are valid until the next `get` of the same name
    (whatever its shape); callers that keep results across steps clone them.
    `generation` counts the times an existing buffer had to be replaced by a larger one: a captured CUDA graph holds raw
    pointers into the buffers of its generation, so whoever caches graphs drops them when the generation has moved
    (growth is geometric, so this happens a handful of times at the start of a run and never for a fixed shape)."""

    GROW = 1.25

    def __init__(self, device):
        self.device = device
        self.pool = {}          # (name, dtype) -> flat tensor
        self.views = {}         # ((name, dtype), shape) -> view of the flat tensor
        self.generation = 0
        self.shared = os.environ.get("ZB_WS_POOL", "1") != "0"

    def get(self, name, shape, dtype=bf16, zero=False):
        n = 1
        for s in shape:
            n *= int(s)
        # ZB_WS_POOL=0 (debugging aid): one allocation per distinct shape, nothing shared between shapes
        key = (name, dtype) if self.shared else (name, dtype, tuple(int(s) for s in shape))
        flat = self.pool.get(key)
        if flat is None or flat.numel() < n:
            cap = n
            if flat is not None:
                cap = max(n, int(flat.numel() * self.GROW))
                self.generation += 1
                self.views = {k: v for k, v in self.views.items() if k[0] != key}   # they alias the old allocation
            flat = torch.zeros(max(cap, 1), dtype=dtype, device=self.device) if zero else \
                torch.empty(max(cap, 1), dtype=dtype, device=self.device)
            self.pool[key] = flat
        # the view itself is remembered per shape (hundreds of requests per step, a few microseconds each to rebuild)
        vkey = (key, shape if type(shape) is tuple else tuple(shape))
        view = self.views.get(vkey)
        if view is None:
            if len(self.views) > 65536:
                self.views.clear()
            view = self.views[vkey] = flat[:n].view(vkey[1])
        return view

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.pool.values())


def dropout_site(name):
    """32-bit id of a dropout call site (the `site` argument of zb_dropout / zb_attention): crc32 of its name,
    e.g. "enc0.self.att", "enc0.self.ln.res", "enc0.ffn.relu", "enc.emb"."""
    return zlib.crc32(name.encode("ascii")) & 0xFFFFFFFF


def _lens(ids):
    """Number of leading non-pad tokens per row; the reference's masks are `id != 0` (models/transformer.py:16)."""
    return (ids != 0).sum(1).to(torch.int32)


def compact_columns(ids):
    """util.remove_invalid_seq (utils/util.py:274-287): drop all-pad columns, always keep column 0."""
    keep = (ids != 0).any(0)
    keep[0] = True
    if bool(keep.all()):
        return ids
    return ids[:, keep].contiguous()


# ------------------------------------------------------------------------------------------------ engine
class Engine(object):
    _dtype_notice = False

    def _range(self, name):
        """NVTX range around a sublayer (ZB_NVTX=1), else a no-op context."""
        return torch.cuda.nvtx.range(name) if self._nvtx else contextlib.nullcontext()

    def __init__(self, hp, src_vocab=None, tgt_vocab=None, device="cuda"):
        if not torch.cuda.is_available():
            raise L.ZeroB200Error("zero_b200 needs a CUDA device (sm_100a); there is no CPU path")
        L.load()
        self.cfg = ModelConfig(hp, src_vocab, tgt_vocab)
        self.device = torch.device(device)
        self.ps = ParamStore(self.cfg, self.device)
        self.ws = Workspace(self.device)
        self.step_count = 0
        # dropout (utils/util.py:75-79), applied only inside the training step; score / infer close it
        # (util.closing_dropout, utils/util.py:106-114).  The seed lives on the device so that CUDA-graph replays of
        # the step see a new mask each time; the trainer advances it once per micro-batch.
        self.rates = {"emb": float(_hp(hp, "dropout", 0.0) or 0.0),
                      "att": float(_hp(hp, "attention_dropout", 0.0) or 0.0),
                      "relu": float(_hp(hp, "relu_dropout", 0.0) or 0.0),
                      "res": float(_hp(hp, "residual_dropout", 0.0) or 0.0)}
        for k, r in self.rates.items():
            if not 0.0 <= r < 1.0:
                raise L.ZeroB200Error("%s dropout rate %r outside [0, 1)" % (k, r))
        self._training = False
        # compute dtype (utils/dtype.py:12-44, run.py:397-399).  The reference computes in float32 by default and in
        # float16 with fp32 master variables on request; this path has ONE numerics contract — bf16 operands, fp32
        # accumulation, fp32 master weights / optimizer state / logits statistics — whatever `default_dtype` says
        # (north_star: outputs within 1e-2 of the reference's fp32 path).  Unknown names are rejected like set_floatx.
        dt = str(_hp(hp, "default_dtype", "float32") or "float32")
        if dt not in ("float32", "float16", "bfloat16"):
            raise L.ZeroB200Error("default_dtype %r: expected float32, float16 or bfloat16 (utils/dtype.py:42)" % dt)
        if dt != "bfloat16" and not Engine._dtype_notice:
            Engine._dtype_notice = True
            import sys
            print("zero_b200: default_dtype=%s requested; the sm_100a path computes in bfloat16 with fp32 accumulation "
                  "and fp32 master weights (there is no fp32 / fp16 kernel set)" % dt, file=sys.stderr, flush=True)
        # ZB_NVTX=1: one NVTX range per sublayer and direction (nsys / ncu --nvtx filtering); off: no overhead
        self._nvtx = os.environ.get("ZB_NVTX") == "1"
        seed0 = (int(_hp(hp, "random_seed", 1234)) * 0x9E3779B97F4A7C15 + 0x1234567) % (1 << 62)
        self.drop_seed = torch.tensor([seed0], dtype=torch.int64, device=self.device)

    # ================================================================================== dropout state
    def _rate(self, kind):
        return self.rates[kind] if self._training else 0.0

    def set_dropout_seed(self, value):
        self.drop_seed.fill_(int(value) % (1 << 62))

    def advance_dropout_seed(self):
        """New masks for the next micro-batch (a device-side increment: legal between CUDA-graph replays)."""
        if any(self.rates.values()):
            self.drop_seed.add_(1)

    def _queue_colsum(self, dy, db):
        if not hasattr(self, "_wg"):
            self._wg, self._cs = [], []
        self._cs.append((dy, db))

    # ================================================================================== side stream
    # Weight-gradient work (wgrad GEMMs + bias column sums) only writes the gradient arena, so it runs on a
    # second stream next to the dgrad critical path (captured as parallel branches of the step's CUDA graph).
    # Backward temporaries exist in two sets (layer parity); layer l waits for the side work of layer l + 2.
    def enable_side_stream(self, on=True):
        self.side = torch.cuda.Stream(device=self.device) if on else None
        self._side_ev = {}

    def _side(self, fn):
        side = getattr(self, "side", None)
        if side is None:
            fn()
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(side):
            side.wait_event(ev)
            fn()

    def _side_layer_begin(self, tag, l):
        if getattr(self, "side", None) is None:
            return
        ev = self._side_ev.get((tag, l + 2))
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def _wgrad(self, x, dy, dw, db=None):
        """Queue dW += x^T dy (and db += column sums of dy) of one func.linear; the queue is flushed once per layer
        as ONE grouped GEMM launch + ONE grouped column-sum launch on the side stream (_side_layer_end)."""
        if not hasattr(self, "_wg"):
            self._wg, self._cs = [], []
        self._wg.append((x, dy, dw))
        if db is not None:
            self._cs.append((dy, db))

    def _flush_wgrads(self):
        wg, cs = getattr(self, "_wg", []), getattr(self, "_cs", [])
        if not wg and not cs:
            return
        self._wg, self._cs = [], []
        self._side(lambda: (ops.gemm_grouped([ops.wgrad_args(*t) for t in wg]), ops.colsum_grouped(cs)))

    def _side_layer_end(self, tag, l):
        self._flush_wgrads()
        side = getattr(self, "side", None)
        if side is None:
            return
        ev = torch.cuda.Event()
        ev.record(side)
        self._side_ev[(tag, l)] = ev

    def _side_join(self):
        self._flush_wgrads()
        side = getattr(self, "side", None)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
            self._side_ev = {}

    # ================================================================================== forward pieces
    def _self_attn_fwd(self, key, x, B, Lq, key_len, causal, sv, tag):
        """func.dot_attention with memory=None (func.py:194-205, 218-256, 277-278)."""
        c, ps, ws = self.cfg, self.ps, self.ws
        N = B * Lq
        qkv = ws.get(tag + ".qkv", (N, 3 * c.d))
        ops.linear_fwd(x, ps.w(key + ".qkv.W"), ps.p(key + ".qkv.b"), qkv)
        q3 = qkv.view(B, Lq, 3 * c.d)
        ctx = ws.get(tag + ".ctx", (N, c.d))
        lse = ws.get(tag + ".lse", (B, c.h, Lq), f32)
        a = ops.attention_args(q3[:, :, :c.d], q3[:, :, c.d:2 * c.d], q3[:, :, 2 * c.d:], ctx.view(B, Lq, c.d), c.h,
                               key_len=key_len, causal=causal, inf_value=c.inf, lse=lse,
                               rpr_k=ps.w(key + ".rpr_k") if c.rpr else None,
                               rpr_v=ps.w(key + ".rpr_v") if c.rpr else None, max_rel=c.max_rel,
                               relu_attn=c.rela,
                               dropout=(self._rate("att"), dropout_site(key + ".att"), self.drop_seed))
        ops.attention_fwd(a)
        feed = self._post_fwd(key, ctx, N, sv, tag)
        y = ws.get(tag + ".y", (N, c.d))
        ops.linear_fwd(feed, ps.w(key + ".o.W"), ps.p(key + ".o.b"), y)
        sv.update(qkv=qkv, ctx=ctx, feed=feed, lse=lse, attn=a, y=y)
        return y

    def _self_attn_bwd(self, key, x, dy, B, Lq, sv, tag):
        """Returns dx (gradient wrt the sublayer input through the attention branch)."""
        c, ps, ws = self.cfg, self.ps, self.ws
        N = B * Lq
        self._wgrad(sv["feed"], dy, ps.g(key + ".o.W"))  # o.b: summed by the LN backward
        dctx = ws.get(tag + ".dctx", (N, c.d))
        ops.linear_dgrad(dy, ps.w(key + ".o.W"), dctx)
        dctx = self._post_bwd(key, dctx, N, sv, tag)
        dqkv = ws.get(tag + ".dqkv", (N, 3 * c.d))
        d3 = dqkv.view(B, Lq, 3 * c.d)
        delta = ws.get(tag + ".delta", (B, c.h, Lq), f32)
        ops.attention_bwd(sv["attn"], dctx.view(B, Lq, c.d), d3[:, :, :c.d], d3[:, :, c.d:2 * c.d], d3[:, :, 2 * c.d:],
                          delta, ps.g(key + ".rpr_k") if c.rpr else None, ps.g(key + ".rpr_v") if c.rpr else None,
                          workspace=self._attn_scratch)
        self._wgrad(x, dqkv, ps.g(key + ".qkv.W"), ps.g(key + ".qkv.b"))
        dx = ws.get(tag + ".dx", (N, c.d))
        ops.linear_dgrad(dqkv, ps.w(key + ".qkv.W"), dx)
        return dx

    def _cross_attn_fwd(self, key, x, enc, B, Lq, S, src_len, sv, tag, kv=None):
        """func.dot_attention with memory (func.py:206-216, 218-256, 277-278).  `kv`: this layer's window of the
        batched memory projection (cfg.batch_mem), else the projection runs here."""
        c, ps, ws = self.cfg, self.ps, self.ws
        N = B * Lq
        q = ws.get(tag + ".q", (N, c.d))
        ops.linear_fwd(x, ps.w(key + ".q.W"), ps.p(key + ".q.b"), q)
        if kv is None:
            kv = ws.get(tag + ".kv", (B * S, 2 * c.d))
            ops.linear_fwd(enc, ps.w(key + ".kv.W"), ps.p(key + ".kv.b"), kv)
        kv3 = kv.view(B, S, 2 * c.d)
        ctx = ws.get(tag + ".ctx", (N, c.d))
        lse = ws.get(tag + ".lse", (B, c.h, Lq), f32)
        a = ops.attention_args(q.view(B, Lq, c.d), kv3[:, :, :c.d], kv3[:, :, c.d:], ctx.view(B, Lq, c.d), c.h,
                               key_len=src_len, causal=False, inf_value=c.inf, lse=lse,
                               rpr_k=ps.w(key + ".rpr_k") if c.rpr else None,
                               rpr_v=ps.w(key + ".rpr_v") if c.rpr else None, max_rel=c.max_rel,
                               relu_attn=c.rela,
                               dropout=(self._rate("att"), dropout_site(key + ".att"), self.drop_seed))
        ops.attention_fwd(a)
        feed = self._post_fwd(key, ctx, N, sv, tag)
        y = ws.get(tag + ".y", (N, c.d))
        ops.linear_fwd(feed, ps.w(key + ".o.W"), ps.p(key + ".o.b"), y)
        sv.update(q=q, kv=kv, ctx=ctx, feed=feed, lse=lse, attn=a, y=y)
        return y

    def _cross_attn_bwd(self, key, x, enc, dy, d_enc, B, Lq, S, sv, tag, dkv=None):
        """`dkv`: this layer's window of the batched d(memory projection) (cfg.batch_mem): its weight / input
        gradients are then taken once for all layers by decode_train_bwd."""
        c, ps, ws = self.cfg, self.ps, self.ws
        N = B * Lq
        batched = dkv is not None
        self._wgrad(sv["feed"], dy, ps.g(key + ".o.W"))  # o.b: summed by the LN backward
        dctx = ws.get(tag + ".dctx", (N, c.d))
        ops.linear_dgrad(dy, ps.w(key + ".o.W"), dctx)
        dctx = self._post_bwd(key, dctx, N, sv, tag)
        dq = ws.get(tag + ".dq", (N, c.d))
        if not batched:
            dkv = ws.get(tag + ".dkv", (B * S, 2 * c.d))
        dkv3 = dkv.view(B, S, 2 * c.d)
        delta = ws.get(tag + ".delta", (B, c.h, Lq), f32)
        ops.attention_bwd(sv["attn"], dctx.view(B, Lq, c.d), dq.view(B, Lq, c.d), dkv3[:, :, :c.d], dkv3[:, :, c.d:],
                          delta, ps.g(key + ".rpr_k") if c.rpr else None, ps.g(key + ".rpr_v") if c.rpr else None,
                          workspace=self._attn_scratch)
        self._wgrad(x, dq, ps.g(key + ".q.W"), ps.g(key + ".q.b"))
        if not batched:
            self._wgrad(enc, dkv, ps.g(key + ".kv.W"), ps.g(key + ".kv.b"))
            # d_enc (fp32, accumulated over decoder layers) += dkv @ Wkv^T
            ops.linear_dgrad(dkv, ps.w(key + ".kv.W"), d_enc, accum=True)
        dx = ws.get(tag + ".dx", (N, c.d))
        ops.linear_dgrad(dq, ps.w(key + ".q.W"), dx)
        return dx

    def _attn_scratch(self, nbytes):
        """Scratch of zb_attention_bwd (the fp32 dq reduction of sequences longer than one 128-key block): one buffer
        per engine, reused by every attention backward of the step (they run in stream order)."""
        return self.ws.get("attn.bwd_scratch", ((nbytes + 3) // 4,), f32)

    def _post_fwd(self, key, ctx, N, sv, tag):
        """ReLA's gated RMS norm between the merged heads and o_map (modules/rela.py:78-81); identity otherwise."""
        c, ps, ws = self.cfg, self.ps, self.ws
        if not c.rela:
            return ctx
        out = ws.get(tag + ".ctxn", (N, c.d))
        rstd = ws.get(tag + ".rrstd", (N,), f32)
        ops.gated_rms_fwd(ctx, out, rstd, ps.p(key + ".post_scale"), ps.p(key + ".post_gate"), c.eps)
        sv["rrstd"] = rstd
        return out

    def _post_bwd(self, key, dctx, N, sv, tag):
        c, ps, ws = self.cfg, self.ps, self.ws
        if not c.rela:
            return dctx
        dx = ws.get(tag + ".dctx2", (N, c.d))
        ops.gated_rms_bwd(sv["ctx"], dctx, sv["rrstd"], ps.p(key + ".post_scale"), ps.p(key + ".post_gate"), dx,
                          ps.g(key + ".post_scale"), ps.g(key + ".post_gate"))
        return dx

    def _ffn_fwd(self, key, x, N, sv, tag):
        """func.ffn_layer (func.py:327-338): relu fused into the first GEMM's epilogue."""
        c, ps, ws = self.cfg, self.ps, self.ws
        h = ws.get(tag + ".h", (N, c.f))
        ops.linear_fwd(x, ps.w(key + ".w1.W"), ps.p(key + ".w1.b"), h, relu=True)
        r = self._rate("relu")
        if r > 0.0:   # func.py:334; dropped and relu-inactive units are both exact zeros of h afterwards
            ops.dropout(h, h, r, self.drop_seed, dropout_site(key + ".relu"))
        y = ws.get(tag + ".y", (N, c.d))
        ops.linear_fwd(h, ps.w(key + ".w2.W"), ps.p(key + ".w2.b"), y)
        sv.update(h=h, y=y, relu_rate=r)
        return y

    def _ffn_bwd(self, key, x, dy, N, sv, tag):
        c, ps, ws = self.cfg, self.ps, self.ws
        self._wgrad(sv["h"], dy, ps.g(key + ".w2.W"))  # w2.b: summed by the LN backward
        dh = ws.get(tag + ".dh", (N, c.f))
        # h > 0 is the relu AND the dropout mask; the kept units carry the 1 / keep scale
        ops.linear_dgrad(dy, ps.w(key + ".w2.W"), dh, relu_mask=sv["h"], alpha=1.0 / (1.0 - sv.get("relu_rate", 0.0)))
        self._wgrad(x, dh, ps.g(key + ".w1.W"), ps.g(key + ".w1.b"))
        dx = ws.get(tag + ".dx", (N, c.d))
        ops.linear_dgrad(dh, ps.w(key + ".w1.W"), dx)
        return dx

    def _ln_fwd(self, key, x, y, N, sv, tag):
        c, ps, ws = self.cfg, self.ps, self.ws
        out = ws.get(tag + ".out", (N, c.d))
        mean = ws.get(tag + ".mean", (N,), f32)
        rstd = ws.get(tag + ".rstd", (N,), f32)
        r = self._rate("res")
        if r > 0.0:   # func.residual_fn: x + dropout(y) (func.py:321-324); y is replaced by its dropped version
            ops.dropout(y, y, r, self.drop_seed, dropout_site(key + ".res"))
        ops.add_ln_fwd(x, y, out, ps.p(key + ".scale"), ps.p(key + ".offset"), mean, rstd, c.eps)
        sv.update(x=x, y=y, mean=mean, rstd=rstd, res_rate=r, res_site=dropout_site(key + ".res"))
        return out

    def _ln_bwd(self, key, d_out, d_out2, N, sv, tag, dbias=None):
        """Returns (ds, dy): ds = gradient wrt the sum x + y, i.e. wrt the skip input x; dy = gradient wrt the branch
        output y (the same tensor unless residual dropout is on, then its masked / rescaled copy).
        dbias: gradient slot of the bias of the linear layer that produced y (the column sum of dy)."""
        c, ps, ws = self.cfg, self.ps, self.ws
        ds = ws.get(tag + ".ds", (N, c.d))
        r = sv.get("res_rate", 0.0)
        ops.add_ln_bwd(sv["x"], sv["y"], d_out, d_out2, sv["mean"], sv["rstd"], ps.p(key + ".scale"), ds,
                       ps.g(key + ".scale"), ps.g(key + ".offset"), dbias if r == 0.0 else None)
        if r == 0.0:
            return ds, ds
        dy = ws.get(tag + ".dyb", (N, c.d))
        ops.dropout(ds, dy, r, self.drop_seed, sv["res_site"])
        if dbias is not None:
            self._queue_colsum(dy, dbias)
        return ds, dy

    # ================================================================================== encoder
    def encode(self, source, save=None, tag="E"):
        """models/transformer.py:15-84.  source: int32 [B, S] (pad = 0).  Returns enc [B*S, d] bf16, src_len."""
        c, ps, ws = self.cfg, self.ps, self.ws
        B, S = source.shape
        N = B * S
        src_len = _lens(source)
        x = ws.get(tag + ".x0", (N, c.d))
        ops.embed_fwd(source, ps.w("src_emb"), ps.p("emb_bias"), x, mult=c.d ** 0.5)
        r_emb = self._rate("emb")
        if r_emb > 0.0:   # models/transformer.py:33
            ops.dropout(x, x, r_emb, self.drop_seed, dropout_site("enc.emb"))
        layers = []
        for l in range(c.nenc):
            key, t = "enc%d" % l, "%s.L%d" % (tag, l)
            sv = {"att": {}, "ln1": {}, "ffn": {}, "ln2": {}, "x_in": x}
            with self._range(key + ".self_attention.fwd"):
                y = self._self_attn_fwd(key + ".self", x, B, S, src_len, False, sv["att"], t + ".att")
                x1 = self._ln_fwd(key + ".self.ln", x, y, N, sv["ln1"], t + ".ln1")
            with self._range(key + ".feed_forward.fwd"):
                y2 = self._ffn_fwd(key + ".ffn", x1, N, sv["ffn"], t + ".ffn")
                x = self._ln_fwd(key + ".ffn.ln", x1, y2, N, sv["ln2"], t + ".ln2")
            sv["x1"] = x1
            layers.append(sv)
        if save is not None:
            save.update(layers=layers, source=source, src_len=src_len, B=B, S=S, emb_rate=r_emb)
        return x, src_len

    def encode_bwd(self, d_enc, save, tag="E", stop_layer=0, carry=None):
        """d_enc: bf16 [B*S, d] gradient wrt the encoder output.  `stop_layer` > 0 processes the layers down to it
        and returns the (d1, d2) pair to hand back as `carry` for the remaining ones: the trainer all-reduces a group
        of layers' gradients while the next group's backward runs."""
        c, ps = self.cfg, self.ps
        B, S = save["B"], save["S"]
        N = B * S
        d1, d2 = (d_enc, None) if carry is None else carry[:2]
        top = c.nenc if carry is None else carry[2]
        for l in reversed(range(stop_layer, top)):
            key, bw = "enc%d" % l, "%s.bw%d" % (tag, l & 1)   # backward temporaries: two sets, by layer parity
            sv = save["layers"][l]
            self._side_layer_begin(tag, l)
            with self._range(key + ".feed_forward.bwd"):
                ds2, dy2 = self._ln_bwd(key + ".ffn.ln", d1, d2, N, sv["ln2"], bw + ".ln2", ps.g(key + ".ffn.w2.b"))
                dx1 = self._ffn_bwd(key + ".ffn", sv["x1"], dy2, N, sv["ffn"], bw + ".ffn")
            with self._range(key + ".self_attention.bwd"):
                ds1, dy1 = self._ln_bwd(key + ".self.ln", ds2, dx1, N, sv["ln1"], bw + ".ln1", ps.g(key + ".self.o.b"))
                dx = self._self_attn_bwd(key + ".self", sv["x_in"], dy1, B, S, sv["att"], bw + ".att")
            self._side_layer_end(tag, l)
            d1, d2 = ds1, dx
        if stop_layer > 0:
            return (d1, d2, stop_layer)
        d1, d2 = self._embed_dropout_bwd(d1, d2, save.get("emb_rate", 0.0), "enc.emb", tag)
        ops.embed_bwd(save["source"], d1, ps.g("src_emb"), ps.g("emb_bias"), mult=c.d ** 0.5, d_out2=d2)
        return None

    def _embed_dropout_bwd(self, d1, d2, rate, site, tag):
        """Gradient through the embedding dropout: the two addends are summed, masked and rescaled in one pass."""
        if rate == 0.0:
            return d1, d2
        d = self.ws.get(tag + ".demb", tuple(d1.shape))
        ops.dropout(d1, d, rate, self.drop_seed, dropout_site(site), x2=d2)
        return d, None

    # ================================================================================== decoder (training)
    def _tgt_table(self):
        return "src_emb" if self.cfg.share_st else "tgt_emb"

    def _softmax_table(self):
        c = self.cfg
        if c.share_st:
            return "src_emb"
        return "tgt_emb" if c.share_ts else "softmax_emb"

    def _vocab_loss(self, feat, target, smooth, want_grad, want_logits, tag):
        """Tied-softmax projection + label-smoothed CE (models/transformer.py:186-211).  Returns
        (loss [1], per_sample [B], logits fp32 [N, V] or None, d_logits bf16 [N, V] or None).
        Unless the caller asks for the logits themselves they are never materialised: zb_vocab_ce reduces them to
        the loss inside the GEMM epilogue and, for training, emits only the bf16 d_logits (K6).  ZB_FUSED_CE=0 keeps
        the two-kernel path (fp32 logits + zb_softmax_ce)."""
        c, ps, ws = self.cfg, self.ps, self.ws
        B, T = target.shape
        N = B * T
        table = ps.w(self._softmax_table())
        nll = ws.get(tag + ".nll", (N,), f32)
        per_sample = ws.get(tag + ".per_sample", (B,), f32)
        loss = ws.get(tag + ".loss", (1,), f32)
        dlogits = self._vocab_rows(tag + ".dlogits", N) if want_grad else None
        fused = (not want_logits and os.environ.get("ZB_FUSED_CE", "1") != "0"
                 and ops.vocab_ce_supported(N, c.d, c.vt, feat, table, dlogits))
        if fused:
            ops.vocab_ce(feat, table, target, nll, smooth, lambda nbytes: ws.get("ce.scratch", ((nbytes + 3) // 4,), f32),
                         d_logits=dlogits, per_sample=per_sample, loss=loss, loss_scale=c.loss_scale)
            return loss, per_sample, None, dlogits
        logits = self._vocab_rows(tag + ".logits", N, f32)
        ops.gemm(feat, table, logits, L.ZB_K_MAJOR, L.ZB_K_MAJOR)  # feature @ E^T, cast fp32 (transformer.py:194-196)
        ops.softmax_ce(logits, target, nll, smooth, d_logits=dlogits, per_sample=per_sample, loss=loss,
                       loss_scale=c.loss_scale)
        return loss, per_sample, logits, dlogits

    def decode_train(self, target, enc, src_len, S, smooth, want_grad, save=None, tag="D", want_logits=True):
        """models/transformer.py:87-218 in training mode.  Returns (loss[1], per_sample[B], logits fp32 [N,V] — None
        unless want_logits)."""
        c, ps, ws = self.cfg, self.ps, self.ws
        if c.aan or c.fuse:
            return self._decode_train_avg(target, enc, src_len, S, smooth, want_grad, save, tag, want_logits)
        B, T = target.shape
        N = B * T
        x = ws.get(tag + ".x0", (N, c.d))
        ops.embed_fwd(target, ps.w(self._tgt_table()), ps.p("emb_bias"), x, mult=c.d ** 0.5, shift=1)
        r_emb = self._rate("emb")
        if r_emb > 0.0:   # models/transformer.py:119
            ops.dropout(x, x, r_emb, self.drop_seed, dropout_site("dec.emb"))
        layers = []
        kv_all = None
        if c.batch_mem:   # every layer's k_map | v_map of the encoder output in one GEMM (n = ndec * 2d)
            kv_all = ws.get(tag + ".kv_all", (B * S, c.ndec * 2 * c.d))
            ops.linear_fwd(enc, ps.w("dec.kvall.W"), ps.p("dec.kvall.b"), kv_all)
        for l in range(c.ndec):
            key, t = "dec%d" % l, "%s.L%d" % (tag, l)
            sv = {"att": {}, "ln1": {}, "cross": {}, "lnc": {}, "ffn": {}, "ln2": {}, "x_in": x}
            # decoder self-attention: causal bias only, no key-padding mask (models/transformer.py:136)
            with self._range(key + ".self_attention.fwd"):
                y = self._self_attn_fwd(key + ".self", x, B, T, None, True, sv["att"], t + ".att")
                x1 = self._ln_fwd(key + ".self.ln", x, y, N, sv["ln1"], t + ".ln1")
            with self._range(key + ".cross_attention.fwd"):
                yc = self._cross_attn_fwd(key + ".cross", x1, enc, B, T, S, src_len, sv["cross"], t + ".cross",
                                          kv=None if kv_all is None else kv_all[:, l * 2 * c.d:(l + 1) * 2 * c.d])
                xc = self._ln_fwd(key + ".cross.ln", x1, yc, N, sv["lnc"], t + ".lnc")
            with self._range(key + ".feed_forward.fwd"):
                y2 = self._ffn_fwd(key + ".ffn", xc, N, sv["ffn"], t + ".ffn")
                x = self._ln_fwd(key + ".ffn.ln", xc, y2, N, sv["ln2"], t + ".ln2")
            sv.update(x1=x1, xc=xc)
            layers.append(sv)
        feat = x
        with self._range("vocab_projection_ce"):
            loss, per_sample, logits, dlogits = self._vocab_loss(feat, target, smooth, want_grad, want_logits, tag)
        if save is not None:
            save.update(layers=layers, target=target, B=B, T=T, S=S, feat=feat, dlogits=dlogits, enc=enc,
                        emb_rate=r_emb)
        return loss, per_sample, logits

    def decode_train_bwd(self, save, d_enc_f32, tag="D"):
        """Backward of decode_train; accumulates the encoder-output gradient into d_enc_f32 (fp32 [B*S, d])."""
        c, ps, ws = self.cfg, self.ps, self.ws
        B, T, S = save["B"], save["T"], save["S"]
        N = B * T
        feat, dlogits, enc = save["feat"], save["dlogits"], save["enc"]
        table = self._softmax_table()
        # dE += dlogits^T feat ; dfeat = dlogits E
        self._side(lambda: ops.gemm(dlogits, feat, ps.g(table), L.ZB_MN_MAJOR, L.ZB_MN_MAJOR, accum=True))
        # dfeat = dlogits @ E is a K = V contraction with few output tiles: split-K into an fp32 scratch
        # (wide 128x256 tiles, every SM busy), then one cast
        dfeat32 = ws.get(tag + ".dfeat32", (N, c.d), f32)
        dfeat32.zero_()
        ops.gemm(dlogits, ps.w(table), dfeat32, L.ZB_K_MAJOR, L.ZB_MN_MAJOR, accum=True)
        dfeat = ws.get(tag + ".dfeat", (N, c.d))
        ops.cast_f32_bf16(dfeat32, dfeat)
        d1, d2 = dfeat, None
        batched = c.batch_mem and not (c.aan or c.fuse)
        dkv_all = ws.get(tag + ".dkv_all", (B * S, c.ndec * 2 * c.d)) if batched else None
        for l in reversed(range(c.ndec)):
            key, bw = "dec%d" % l, "%s.bw%d" % (tag, l & 1)
            sv = save["layers"][l]
            self._side_layer_begin(tag, l)
            ds2, dy2 = self._ln_bwd(key + ".ffn.ln", d1, d2, N, sv["ln2"], bw + ".ln2", ps.g(key + ".ffn.w2.b"))
            dxc = self._ffn_bwd(key + ".ffn", sv["xc"], dy2, N, sv["ffn"], bw + ".ffn")
            if c.aan or c.fuse:
                d1, d2 = self._avg_layer_bwd(key, bw, sv, ds2, dxc, enc, d_enc_f32, B, T, S, save)
            else:
                dsc, dyc = self._ln_bwd(key + ".cross.ln", ds2, dxc, N, sv["lnc"], bw + ".lnc",
                                        ps.g(key + ".cross.o.b"))
                dx1 = self._cross_attn_bwd(key + ".cross", sv["x1"], enc, dyc, d_enc_f32, B, T, S, sv["cross"],
                                           bw + ".cross",
                                           dkv=None if dkv_all is None else dkv_all[:, l * 2 * c.d:(l + 1) * 2 * c.d])
                ds1, dy1 = self._ln_bwd(key + ".self.ln", dsc, dx1, N, sv["ln1"], bw + ".ln1",
                                        ps.g(key + ".self.o.b"))
                dx = self._self_attn_bwd(key + ".self", sv["x_in"], dy1, B, T, sv["att"], bw + ".att")
                d1, d2 = ds1, dx
            self._side_layer_end(tag, l)
        if batched:
            # all layers at once: dWkv += enc^T dkv_all (+ bias sums), d_enc += dkv_all Wkv_all^T (k = ndec * 2d)
            self._wgrad(enc, dkv_all, ps.g("dec.kvall.W"), ps.g("dec.kvall.b"))
            self._flush_wgrads()
            ops.linear_dgrad(dkv_all, ps.w("dec.kvall.W"), d_enc_f32, accum=True)
        d1, d2 = self._embed_dropout_bwd(d1, d2, save.get("emb_rate", 0.0), "dec.emb", tag)
        ops.embed_bwd(save["target"], d1, ps.g(self._tgt_table()), ps.g("emb_bias"), mult=c.d ** 0.5, shift=1, d_out2=d2)

    # ================================================================================== public steps
    def _vocab_rows(self, tag, rows, dtype=bf16):
        """[rows, V] buffer over the target vocabulary; its row pitch is a multiple of 8 elements (TMA), so for a
        vocabulary size that is not it is a strided view of a slightly wider buffer."""
        c = self.cfg
        buf = self.ws.get(tag, (rows, c.vt_pitch), dtype)
        return buf if c.vt_pitch == c.vt else buf[:, :c.vt]

    def _dense_logits(self, logits):
        """decoding_fn hands out contiguous fp32 [rows, V] (models/transformer.py:267-283; zb_beam_step indexes it
        densely): a plain copy when the pitch had to be padded."""
        if logits.is_contiguous():
            return logits
        dense = self.ws.get("dec.logits_dense", tuple(logits.shape), f32)
        dense.copy_(logits)
        return dense

    @staticmethod
    def _prep_ids(ids, device, compact=True):
        ids = torch.as_tensor(ids)
        if compact and ids.device.type == "cpu":
            # host-resident ids (the data pipeline's pinned batches): drop the all-pad columns before the copy — the
            # same result without the device -> host round trip compact_columns needs for a device tensor
            ids, compact = compact_columns(ids), False
        if ids.device != device:
            ids = ids.to(device, non_blocking=True)
        ids = ids.to(torch.int32)
        return (compact_columns(ids) if compact else ids).contiguous()

    def forward_backward(self, source, target, zero_grad=True, compact=True):
        """train_fn + tf.gradients (models/transformer.py:221-232, main.py:22-45) for one tower.
        Gradients land in self.ps.grad (fp32).  Returns the device loss tensor [1]."""
        loss = self.forward_backward_decoder(source, target, zero_grad, compact)
        self.backward_encoder()
        return loss

    def forward_backward_decoder(self, source, target, zero_grad=True, compact=True):
        """Phase 1 of a training step: forward of the whole model + backward of the decoder.  When it returns
        (stream order), every gradient in ps.grad[ps.dec_offset:] is final — the trainer starts all-reducing that
        bucket while phase 2 runs."""
        c, ws = self.cfg, self.ws
        source = self._prep_ids(source, self.device, compact)
        target = self._prep_ids(target, self.device, compact)
        if zero_grad:
            self.ps.zero_grad()
        B, S = source.shape
        if B == 0 or target.shape[0] == 0:
            # zero-shape guard (models/transformer.py:213-216): an empty tower contributes loss 0 and no gradients
            self._pending = None
            return torch.zeros(1, dtype=f32, device=self.device)
        esave, dsave = {}, {}
        self._training = True     # dropout is part of train_fn only (score_fn / infer_fn close it)
        try:
            enc, src_len = self.encode(source, esave)
            loss, per_sample, _ = self.decode_train(target, enc, src_len, S, c.smooth, True, dsave, want_logits=False)
        finally:
            self._training = False
        d_enc32 = ws.get("d_enc32", (B * S, c.d), f32)
        d_enc32.zero_()
        self.decode_train_bwd(dsave, d_enc32)
        d_enc = ws.get("d_enc", (B * S, c.d))
        ops.cast_f32_bf16(d_enc32, d_enc)
        self._side_join()
        self._pending = (d_enc, esave)
        return loss

    def backward_encoder(self, stop_layer=0):
        """Phase 2: backward of the encoder (finalises ps.grad[:ps.dec_offset]).  With stop_layer > 0 only the layers
        from the current position down to `stop_layer` run (their gradients — arena [enc_layer_offset[stop_layer], end
        of the previous group) — are final when the call returns in stream order); call again to continue."""
        if self._pending is None:      # empty tower (see forward_backward_decoder)
            return
        d_enc, esave = self._pending[:2]
        carry = self._pending[2] if len(self._pending) > 2 else None
        carry = self.encode_bwd(d_enc, esave, stop_layer=stop_layer, carry=carry)
        self._side_join()
        self._pending = None if carry is None else (d_enc, esave, carry)

    def encoder_buckets(self, groups):
        """Arena ranges [(stop_layer, lo, hi)] of `groups` groups of encoder layers, last layers first, followed by
        (0, 0, first layer offset) for the source embedding + shared bias, whose gradients are the last to complete."""
        c, ps = self.cfg, self.ps
        groups = max(1, min(int(groups), c.nenc))
        per = (c.nenc + groups - 1) // groups
        out, hi, top = [], ps.dec_offset, c.nenc
        while top > 0:
            stop = max(0, top - per)
            out.append((stop, ps.enc_layer_offset[stop], hi))
            hi, top = ps.enc_layer_offset[stop], stop
        out.append((0, 0, ps.enc_layer_offset[0]))
        return out

    def train_loss(self, source, target):
        """train_fn forward only -> (loss, per_sample, logits)."""
        source = self._prep_ids(source, self.device)
        target = self._prep_ids(target, self.device)
        enc, src_len = self.encode(source)
        return self.decode_train(target, enc, src_len, source.shape[1], self.cfg.smooth, False)

    def score(self, source, target):
        """score_fn (models/transformer.py:235-249): label smoothing off, returns per-sentence NLL [B]."""
        source = self._prep_ids(source, self.device)
        target = self._prep_ids(target, self.device)
        if source.shape[0] == 0:
            return torch.zeros(0, dtype=f32, device=self.device)
        enc, src_len = self.encode(source)
        return self.decode_train(target, enc, src_len, source.shape[1], 0.0, False, want_logits=False)[1]


# ------------------------------------------------------------------------------------------------ cached decode
class DecodeState(object):
    """State of infer_fn's (encoding_fn, decoding_fn) pair in 'cache' search mode
    (models/transformer.py:252-285).  Per-sentence tensors (encoder output, projected
    memory mk/mv) are kept [B, ...] and shared by the beams of a sentence; only the per-beam self-attention
    caches are reordered (double-buffered gather of the filled prefix)."""

    def __init__(self, engine, enc, src_len, B, S):
        self.engine = engine
        self.device = engine.device
        self.vocab = engine.cfg.vt
        self.enc, self.src_len, self.B, self.S = enc, src_len, B, S
        self.K = 1
        self.mem = None
        self.cache = None
        self.cache_alt = None

    def begin_search(self, beam, cap=None):
        """search.py:36-39,56-77: tile per-beam state, run the dummy step that creates mk / mv."""
        eng, c = self.engine, self.engine.cfg
        self.K = int(beam)
        R = self.B * self.K
        if cap is None:
            cap = int(self.src_len.max().item()) + int(getattr(eng, "decode_length", 50)) + 2
        self.cap = cap
        self.mem = []
        for l in range(c.ndec):
            key = "dec%d.cross" % l
            kv = eng.ws.get("dec.mem%d" % l, (self.B * self.S, 2 * c.d))
            ops.linear_fwd(self.enc, eng.ps.w(key + ".kv.W"), eng.ps.p(key + ".kv.b"), kv)
            self.mem.append(kv.view(self.B, self.S, 2 * c.d))
        if not (c.aan or c.fuse):
            self.cache = [eng.ws.get("dec.cacheA%d" % l, (R, cap, 3 * c.d)) for l in range(c.ndec)]
            self.cache_alt = [eng.ws.get("dec.cacheB%d" % l, (R, cap, 3 * c.d)) for l in range(c.ndec)]

    def reorder(self, parent, t):
        """search.py:205-209: new alive beam r continues previous beam parent[r]; moves positions [0, t]."""
        c = self.engine.cfg
        if self.cache is None:
            return
        for l in range(c.ndec):
            ops.gather_rows(self.cache[l], parent, self.cache_alt[l], row_elems=(t + 1) * 3 * c.d)
        self.swap_buffers()

    def swap_buffers(self):
        """Host-side half of reorder(): the double-buffered caches trade places (also called after a CUDA-graph
        replay of a step, which runs the kernels but not this bookkeeping)."""
        if self.cache is not None:
            self.cache, self.cache_alt = self.cache_alt, self.cache


def _engine_encoding_fn(self, source):
    source = self._prep_ids(source, self.device)
    enc, src_len = self.encode(source, tag="I")
    # fixed address across batches of the same shape: decode steps are replayed from CUDA graphs
    src_len_static = self.ws.get("I.src_len", (source.shape[0],), torch.int32)
    src_len_static.copy_(src_len)
    return DecodeState(self, enc, src_len_static, source.shape[0], source.shape[1])


def _engine_step_logits(self, x, R, candidates):
    """The decode step's vocabulary projection (models/transformer.py:186-196): fp32 [R, V] logits, or — when the
    search asked for `candidates` = {skip_col, temperature} — the K8 fused reduction of them (ops.vocab_topk)."""
    c, ps, ws = self.cfg, self.ps, self.ws
    table = ps.w(self._softmax_table())
    if candidates is not None:
        return ops.vocab_topk(x, table, lambda nbytes: ws.get("dec.cand", ((nbytes + 3) // 4,), f32), **candidates)
    logits = self._vocab_rows("dec.logits", R, f32)
    ops.gemm(x, table, logits, L.ZB_K_MAJOR, L.ZB_K_MAJOR)
    return self._dense_logits(logits)


def _engine_decoding_fn(self, target, state, time, candidates=None):
    """decoding_fn(target [R,1], state, time) -> (logits fp32 [R, V], state)  (models/transformer.py:267-283).
    `candidates` (search.beam_search's own fused path only): return ops.BeamCandidates instead of the logits."""
    c, ps, ws = self.cfg, self.ps, self.ws
    if state.mem is None:
        state.begin_search(state.K)
    if c.aan or c.fuse:
        return self._decoding_fn_avg(target, state, time, candidates)
    t = int(time)
    R = target.shape[0]
    K = state.K
    x = ws.get("dec.x", (R, c.d))
    ops.embed_fwd(target, ps.w(self._tgt_table()), ps.p("emb_bias"), x.view(R, 1, c.d), mult=c.d ** 0.5,
                  zero_if_all_pad=True, time=t)
    for l in range(c.ndec):
        key = "dec%d" % l
        cache = state.cache[l]
        ops.linear_fwd(x, ps.w(key + ".self.qkv.W"), ps.p(key + ".self.qkv.b"), cache[:, t, :])
        ctx = ws.get("dec.ctx", (R, c.d))
        a = ops.attention_args(cache[:, t:t + 1, :c.d], cache[:, :t + 1, c.d:2 * c.d], cache[:, :t + 1, 2 * c.d:],
                               ctx.view(R, 1, c.d), c.h, q_offset=t, inf_value=c.inf,
                               rpr_k=ps.w(key + ".self.rpr_k") if c.rpr else None,
                               rpr_v=ps.w(key + ".self.rpr_v") if c.rpr else None, max_rel=c.max_rel,
                               relu_attn=c.rela)
        ops.attention_fwd(a)
        ctx = self._post_attn(key + ".self", ctx, R)
        x1 = ws.get("dec.x1", (R, c.d))
        self._decode_proj_ln(ctx, key + ".self.o", x, x1, key + ".self.ln", R)
        q = ws.get("dec.q", (R, c.d))
        ops.linear_fwd(x1, ps.w(key + ".cross.q.W"), ps.p(key + ".cross.q.b"), q)
        mem = state.mem[l]
        a = ops.attention_args(q.view(R, 1, c.d), mem[:, :, :c.d], mem[:, :, c.d:], ctx.view(R, 1, c.d), c.h,
                               key_len=state.src_len, q_offset=t, inf_value=c.inf, kv_group=K,
                               rpr_k=ps.w(key + ".cross.rpr_k") if c.rpr else None,
                               rpr_v=ps.w(key + ".cross.rpr_v") if c.rpr else None, max_rel=c.max_rel,
                               relu_attn=c.rela)
        ops.attention_fwd(a)
        ctx = self._post_attn(key + ".cross", ctx, R)
        xc = ws.get("dec.xc", (R, c.d))
        self._decode_proj_ln(ctx, key + ".cross.o", x1, xc, key + ".cross.ln", R)
        h = ws.get("dec.h", (R, c.f))
        ops.linear_fwd(xc, ps.w(key + ".ffn.w1.W"), ps.p(key + ".ffn.w1.b"), h, relu=True)
        self._decode_proj_ln(h, key + ".ffn.w2", xc, x, key + ".ffn.ln", R)
    return self._step_logits(x, R, candidates), state


def _engine_decoding_fn_dev(self, target, state, time):
    """decoding_fn of search_mode = "dev" (models/transformer.py:276-281, search.py:129-140): no caches — the whole
    decoder is re-run, teacher-forced, on the partial target [R, t + 1] (the tokens generated so far followed by a
    placeholder), and the logits of the last position are returned (models/transformer.py:183-184).  A development aid
    in the reference ("at the cost of slower decoding") and here the cache-independent check of the cached path."""
    c = self.cfg
    target = self._prep_ids(target, self.device, compact=False)
    R, T = target.shape
    K = max(1, R // max(state.B, 1))
    S = state.S
    enc = state.enc.view(state.B, S, c.d)
    if K > 1:       # every beam of a sentence reads the same memory: tile it (the reference tiles its whole state)
        enc_r = self.ws.get("dev.enc", (R * S, c.d))
        enc_r.view(state.B, K, S, c.d).copy_(enc.unsqueeze(1).expand(state.B, K, S, c.d))
        src_len = state.src_len.repeat_interleave(K)
    else:
        enc_r, src_len = state.enc, state.src_len
    # the placeholder in the last column must count as a real token for the models that mask by target id (AAN / fuse)
    _, _, logits = self.decode_train(target, enc_r, src_len, S, 0.0, False, tag="V", want_logits=True)
    last = self.ws.get("dev.logits", (R, c.vt), f32)
    last.copy_(logits.view(R, T, -1)[:, T - 1, :])
    return last, state


def _engine_decode_proj_ln(self, inp, lin, res, out, ln, rows):
    """out = LayerNorm(res + inp @ W + b) of a decode step (func.linear + residual_fn + layer_norm on [rows, .] with
    rows = batch * beam).  A 256-row projection makes only rows / 128 x n / 64 output tiles — 16 CTAs for n = 512, each
    walking the whole k dimension at the SM's L2 ingest rate (0.26 us per 64-k block: 8.6 us for the k = 2048 FFN
    output projection, profiles/r02t_decode_gemm_trace.log) — so the k dimension is split over ~8x as many CTAs, which
    add their fp32 partial tiles into one accumulator by TMA reduce-add; the residual + LayerNorm kernel that consumes
    it adds the bias, rounds like the unsplit epilogue does, and clears the accumulator for the next projection.
    MEASURED (configs[2], profiles/r02u_decode_splitk_ab.log): 0.548-0.554 ms/step against 0.542-0.543 for the unsplit
    projection + bf16 round trip — the reduce-adds and the fp32 traffic cost what the shorter k loops save — so the
    split is opt-in (ZB_DECODE_SPLITK=1); parity-tested either way."""
    c, ps, ws = self.cfg, self.ps, self.ws
    w, b = ps.w(lin + ".W"), ps.p(lin + ".b")
    k, n = w.shape
    kb = (k + 63) // 64
    tiles = ((rows + 127) // 128) * ((n + 63) // 64)
    splits = min(kb // 2, max(1, 148 // tiles))
    if rows > 512 or splits < 2 or os.environ.get("ZB_DECODE_SPLITK", "0") != "1":
        y = ws.get("dec.y", (rows, n))
        ops.linear_fwd(inp, w, b, y)
        ops.add_ln_fwd(res, y, out, ps.p(ln + ".scale"), ps.p(ln + ".offset"), eps=c.eps)
        return out
    y32 = ws.get("dec.y32", (rows, n), f32, zero=True)    # zero when allocated; kept clean by its consumer
    ops.gemm(inp, w, y32, L.ZB_K_MAJOR, L.ZB_MN_MAJOR, accum=True, split_k=splits)
    ops.add_ln_fwd(res, None, out, ps.p(ln + ".scale"), ps.p(ln + ".offset"), eps=c.eps, y32=y32, ybias=b)
    return out


def _engine_post_attn(self, key, ctx, rows):
    """ReLA's gated RMS norm on the merged heads (modules/rela.py:78-81, 95-109); identity otherwise."""
    c, ps = self.cfg, self.ps
    if not c.rela:
        return ctx
    out = self.ws.get("dec.ctxn", (rows, c.d))
    ops.gated_rms_fwd(ctx, out, None, ps.p(key + ".post_scale"), ps.p(key + ".post_gate"), c.eps)
    return out


Engine.encoding_fn = _engine_encoding_fn
Engine._step_logits = _engine_step_logits
Engine.decoding_fn = _engine_decoding_fn
Engine.decoding_fn_dev = _engine_decoding_fn_dev
Engine._post_attn = _engine_post_attn
Engine._decode_proj_ln = _engine_decode_proj_ln

from . import engine_avg  # noqa: E402,F401  (registers the average-attention family on Engine / DecodeState)
