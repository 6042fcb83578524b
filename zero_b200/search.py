"""search.beam_search (reference search.py:19-275) on the device.

`beam_search(features, encoding_fn, decoding_fn, params)` keeps the reference's signature and result
({'seq': int32 [B, beam, L], 'score': fp32 [B, beam]}).  The per-step expansion (log-softmax, EOS ban at
t = 0, GNMT length penalty, top-2k over beam*V, alive / finished bookkeeping) is one fused kernel
(zb_beam_step, csrc/beam.cu); the loop condition of search.py:85-113 is zb_beam_cond.  Model state is not
tiled per beam the way search.py:36-39 does: per-sentence tensors (encoder output, projected memory) stay
[B, ...] and only per-beam caches are reordered, by the `parent` rows the step kernel returns.

The reference wraps the step in a tf.while_loop (search.py:251); here every step index t owns a CUDA graph
(decoder step + beam step + cache reorder, ~100 kernels) captured the second time that (shape, t) is seen and
replayed afterwards, so the host only evaluates the loop condition - one step behind the device, see beam_search.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops

F32_MIN = float(np.finfo(np.float32).min)


class BeamState(object):
    """Device-side alive / finished buffers of search.py:46-54 plus the step / condition kernels."""

    def __init__(self, batch, beam, vocab, source, decode_length, alpha, temperature, inf_value, device,
                 eos_id=2, pad_id=0, cap=None):
        self.B, self.K, self.V = batch, beam, vocab
        self.alpha, self.temperature, self.inf_value = float(alpha), float(temperature), float(inf_value)
        self.eos_id, self.pad_id = eos_id, pad_id
        self.decode_length = int(decode_length)
        self.device = device
        source = torch.as_tensor(source)
        self.cap = int(cap) if cap is not None else int(source.shape[1]) + self.decode_length + 2
        i32, f32 = torch.int32, torch.float32
        self.max_len = torch.zeros(batch, dtype=i32, device=device)
        self.max_penalty = torch.ones(batch, dtype=f32, device=device)
        self.alive_seq = torch.zeros(batch, beam, self.cap, dtype=i32, device=device)
        self.alive_logp = torch.zeros(batch, beam, dtype=f32, device=device)
        self.alive_score = torch.zeros(batch, beam, dtype=f32, device=device)
        self.fin_seq = torch.zeros(batch, beam, self.cap, dtype=i32, device=device)
        self.fin_score = torch.zeros(batch, beam, dtype=f32, device=device)
        self.fin_flag = torch.zeros(batch, beam, dtype=i32, device=device)
        self.parent = torch.arange(batch * beam, dtype=i32, device=device)
        self.tmp_seq = torch.zeros(batch, 3 * beam, self.cap, dtype=i32, device=device)
        self.active = torch.ones(1, dtype=i32, device=device)
        # scratch of the row-parallel step kernel (one CTA per (sentence, beam) row); ZB_BEAM_ROWS=0 keeps the
        # one-CTA-per-sentence kernel
        self.row_ws = None if os.environ.get("ZB_BEAM_ROWS", "1") == "0" else \
            torch.zeros(batch * (4 * beam * beam + 1), dtype=f32, device=device)
        self.tok_buf = torch.zeros(batch * beam, 1, dtype=i32, device=device)
        self.noise_seed = torch.zeros(1, dtype=torch.int64, device=device)   # zb_gumbel_add (noise beam search)
        self._host_flag, self._flag_events = None, None
        init_logp = torch.full((batch, beam), F32_MIN, dtype=f32)
        init_logp[:, 0] = 0.0
        self._init_logp = init_logp.to(device)
        self.time = 0
        self.reset(source)

    def reset(self, source):
        """search.py:46-54 initial state for a new batch (buffers are reused, addresses stay fixed)."""
        src = torch.as_tensor(source)
        src_len = (src != 0).sum(1).cpu()
        assert int(src_len.max()) + self.decode_length + 2 <= self.cap, "BeamState capacity too small"
        self.max_len.copy_((src_len + self.decode_length).to(torch.int32))
        # ((5 + max_len) / 6) ^ alpha in fp32 on the host, like the reference's tf.pow on a float tensor
        ml = src_len.float() + float(self.decode_length)
        self.max_penalty.copy_(torch.pow((5.0 + ml) / 6.0, self.alpha))
        self.alive_seq.fill_(self.pad_id)
        self.alive_logp.copy_(self._init_logp)
        self.alive_score.zero_()
        self.fin_seq.zero_()
        self.fin_score.fill_(F32_MIN)
        self.fin_flag.zero_()
        if self.row_ws is not None:
            self.row_ws.zero_()
        self.time = 0

    def _args(self, logits, t):
        pen = float(torch.pow(torch.tensor((5.0 + float(t + 1)) / 6.0, dtype=torch.float32), self.alpha))
        cand = None
        if isinstance(logits, ops.BeamCandidates):
            # the vocabulary projection already applied this step's temperature and EOS ban (csrc/vocab_topk.cu)
            assert logits.rows == self.B * self.K and logits.vocab == self.V
            assert logits.temperature == self.temperature and logits.skip_col == (self.eos_id if t < 1 else -1)
            logits, cand = None, logits.buffer
        return ops.beam_args(
            cand=cand, logits=logits, batch=self.B, beam=self.K, vocab=self.V, time=int(t), eos_id=self.eos_id,
            pad_id=self.pad_id, temperature=self.temperature, inf_value=self.inf_value, length_penalty=pen,
            max_len=self.max_len, max_penalty=self.max_penalty, seq_cap=self.cap, alive_seq=self.alive_seq,
            alive_logp=self.alive_logp, alive_score=self.alive_score, fin_seq=self.fin_seq,
            fin_score=self.fin_score, fin_flag=self.fin_flag, parent=self.parent, tmp_seq=self.tmp_seq,
            active=self.active, row_ws=self.row_ws)

    def not_finished(self, t):
        """search.py:85-113, evaluated on the device; one 4-byte read-back."""
        if t + 2 > self.cap:
            return False
        ops.beam_cond(self._args(None, t))
        return bool(self.active.item())

    def cond_async(self, t):
        """Enqueue search.py:85-113 for step t and an asynchronous copy of the flag to pinned host memory.
        Returns the slot to pass to cond_wait, or None when the sequence buffers are full."""
        if t + 2 > self.cap:
            return None
        if self._host_flag is None:
            self._host_flag = torch.zeros(2, dtype=torch.int32).pin_memory()
            self._flag_events = [torch.cuda.Event(), torch.cuda.Event()]
        slot = t & 1
        ops.beam_cond(self._args(None, t))
        self._host_flag[slot:slot + 1].copy_(self.active, non_blocking=True)
        self._flag_events[slot].record()
        return slot

    def cond_wait(self, slot):
        self._flag_events[slot].synchronize()
        return bool(int(self._host_flag[slot]))

    def last_tokens(self, t):
        """[B*beam, 1] int32: the token fed to decoding_fn at step t (search.py:130); static buffer."""
        self.tok_buf.copy_(self.alive_seq[:, :, t].reshape(self.B * self.K, 1))
        return self.tok_buf

    def prefix_tokens(self, t):
        """search_mode = "dev" (search.py:131-140): the partial target [B*beam, t + 1] — the tokens generated so far
        (the leading start column dropped) followed by one placeholder token (id 1)."""
        seq = self.alive_seq[:, :, 1:t + 1].reshape(self.B * self.K, t)
        return torch.cat([seq, torch.ones(self.B * self.K, 1, dtype=seq.dtype, device=seq.device)], 1).contiguous()

    def candidate_request(self, t):
        """What decoding_fn(..., candidates=) needs to reduce step t's logits for this search (ops.vocab_topk)."""
        return {"skip_col": self.eos_id if t < 1 else -1, "temperature": self.temperature}

    def step(self, logits, t):
        """`logits`: fp32 [B*beam, V], or the ops.BeamCandidates the fused vocabulary projection left instead."""
        ops.beam_step(self._args(logits, t))
        self.time = t + 1
        return self.parent

    def result(self):
        """search.py:258-275: finished beams where any exist, alive beams otherwise; BOS column dropped."""
        t = self.time
        any_fin = (self.fin_flag != 0).any(1)
        seq = torch.where(any_fin[:, None, None], self.fin_seq[:, :, :t + 1], self.alive_seq[:, :, :t + 1])
        score = torch.where(any_fin[:, None], self.fin_score, self.alive_score)
        return {"seq": seq[:, :, 1:].clone(), "score": score.clone()}


# Captured decode steps kept per engine (one graph per batch shape and step index); past this many, further shapes run
# their steps eagerly.  A dev set decoded at every evaluation repeats its batch shapes, so its steps are all captured
# after the second evaluation; a one-off test set never reaches the second visit a capture needs.
MAX_DECODE_GRAPHS = 8192
MAX_BEAM_STATES = 256


def beam_search(features, encoding_fn, decoding_fn, params):
    """Drop-in for reference search.beam_search (search.py:19).  `encoding_fn(source) -> state`,
    `decoding_fn(target [B*beam,1], state, time) -> (logits fp32 [B*beam,V], state)` as returned by
    zero_b200's infer_fn; `state.reorder(parent)` replaces the gather_nd over the tiled state."""
    # search.py:143-145: Gumbel noise on the step logits turns the beam into top-k sampling without replacement.  The
    # draws come from the library's counter-based generator (seeded from random_seed, advanced per search), so runs
    # are reproducible here but — like any sampler — not draw-for-draw comparable with TF's tf.random_uniform.
    noise = bool(getattr(params, "enable_noise_beam_search", False))
    source = features["source"]
    state = encoding_fn(source)
    eng = state.engine
    dev = state.device
    src = torch.as_tensor(source).to(dev)
    B = src.shape[0]
    K = int(params.beam_size)
    cap = int(src.shape[1]) + int(params.decode_length) + 2
    # Everything a captured step or the bookkeeping buffers bake in: encoding_fn compacts all-pad columns, so the
    # memory width is state.S (NOT the padded width of `source`); temperature / inf / eos / pad are kernel arguments.
    key = (B, K, state.vocab, cap, int(state.S), float(params.decode_alpha), int(params.decode_length), noise,
           float(getattr(params, "beam_search_temperature", 1.0)), float(getattr(params, "dtype_inf", 1e8)),
           int(params.tgt_vocab.eos()), int(params.tgt_vocab.pad()))
    cache = eng.__dict__.setdefault("_beam_states", {})
    st = cache.get(key)
    if st is None:
        st = BeamState(B, K, state.vocab, src, params.decode_length, params.decode_alpha,
                       getattr(params, "beam_search_temperature", 1.0), getattr(params, "dtype_inf", 1e8), dev,
                       eos_id=params.tgt_vocab.eos(), pad_id=params.tgt_vocab.pad(), cap=cap)
        if len(cache) >= MAX_BEAM_STATES:
            # drop the oldest batch shape's bookkeeping buffers — and the captured steps that point into them
            old = next(iter(cache))
            cache.pop(old)
            for store in (eng.__dict__.get("_decode_graphs", {}), eng.__dict__.get("_decode_seen", {})):
                for gk in [gk for gk in store if gk[:len(old)] == old]:
                    del store[gk]
        cache[key] = st
        st.noise_seed.fill_(int(getattr(params, "random_seed", 1234)))
    else:
        st.reset(src)
    if noise:
        st.noise_seed.add_(1)          # a device-side value: replayed step graphs read the new seed
    state.begin_search(K, cap)
    dev_mode = str(getattr(params, "search_mode", "cache")) != "cache"
    # CUDA-graph replay is only valid for the engine's own decoding_fn (a wrapped one may have side effects)
    own = getattr(decoding_fn, "__self__", None) is eng and getattr(decoding_fn, "__func__", None) is type(eng).decoding_fn
    use_graph = own and bool(getattr(params, "decode_graph", True)) and not dev_mode
    # K8 fused (csrc/vocab_topk.cu): the engine's own cached step hands the beam step per-part top-8 candidates
    # instead of [B*beam, V] logits.  Needs the logits themselves for nothing else: not with Gumbel noise on them,
    # not through a wrapped decoding_fn (it expects logits), not in "dev" mode.  Default since the r02ae A/B (0.542 -> 0.503 ms/step); ZB_BEAM_FUSED=0 keeps the logits path.
    fused = (own and not noise and not dev_mode and st.row_ws is not None
             and os.environ.get("ZB_BEAM_FUSED", "1") != "0"
             and ops.vocab_topk_supported(eng.cfg.d, state.vocab, K))
    graphs = eng.__dict__.setdefault("_decode_graphs", {})
    seen = eng.__dict__.setdefault("_decode_seen", {})
    if eng.__dict__.get("_decode_graphs_gen", 0) != eng.ws.generation:
        # the encoder pass / begin_search above (or an earlier, larger batch) replaced a workspace buffer by a larger
        # one: the captured steps point into the old allocation
        graphs.clear()
        seen.clear()
        eng.__dict__["_decode_graphs_gen"] = eng.ws.generation

    def run_step(t):
        nonlocal state
        if fused:
            logits, state = decoding_fn(st.last_tokens(t), state, t, candidates=st.candidate_request(t))
        else:
            logits, state = decoding_fn(st.prefix_tokens(t) if dev_mode else st.last_tokens(t), state, t)
        if noise:
            ops.gumbel_add(logits, st.noise_seed, t, eps=float(getattr(params, "dtype_epsilon", 1e-8)))
        parent = st.step(logits, t)
        state.reorder(parent, t)

    # The loop condition of step t is one kernel + a 4-byte read-back.  Waiting for it before launching step t
    # leaves the GPU idle for a launch + sync round trip every step, so with the engine's own decoding_fn the host
    # runs one step ahead: step t (and the condition of t + 1) are enqueued before the flag of t is read.  A step
    # enqueued past the end is harmless: zb_beam_step is a no-op once active[0] == 0 and the per-beam state it
    # would reorder is discarded with the search.  ZB_DECODE_SPEC=0 restores the lock-step loop.
    speculate = own and os.environ.get("ZB_DECODE_SPEC", "1") != "0"

    def enqueue_step(t):
        gkey = key + (t, fused)
        g = graphs.get(gkey) if use_graph else None
        if g is not None:
            g.replay()
            state.swap_buffers()
        elif use_graph and seen.get(gkey) and len(graphs) < MAX_DECODE_GRAPHS:
            # second visit: every workspace buffer exists, capture this step (capture does not execute it)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run_step(t)      # records the kernels; the host-side buffer swap happens now
            graphs[gkey] = g
            g.replay()
        else:
            seen[gkey] = True
            run_step(t)

    t = 0
    slot = st.cond_async(0)
    while slot is not None:
        if speculate:
            enqueue_step(t)
            nxt = st.cond_async(t + 1)
            if not st.cond_wait(slot):
                break
        else:
            if not st.cond_wait(slot):
                break
            enqueue_step(t)
            nxt = st.cond_async(t + 1)
        t += 1
        slot = nxt
    st.time = t
    return st.result()
