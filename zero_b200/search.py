"""search.beam_search (reference search.py:19-275) on the device.

`beam_search(features, encoding_fn, decoding_fn, params)` keeps the reference's signature and result
({'seq': int32 [B, beam, L], 'score': fp32 [B, beam]}).  The per-step expansion (log-softmax, EOS ban at
t = 0, GNMT length penalty, top-2k over beam*V, alive / finished bookkeeping) is one fused kernel
(zb_beam_step, csrc/beam.cu); the loop condition of search.py:85-113 is zb_beam_cond.  Model state is not
tiled per beam the way search.py:36-39 does: per-sentence tensors (encoder output, projected memory) stay
[B, ...] and only per-beam caches are reordered, by the `parent` rows the step kernel returns.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

F32_MIN = float(np.finfo(np.float32).min)


class BeamState(object):
    """Device-side alive / finished buffers of search.py:46-54 plus the step / condition kernels."""

    def __init__(self, batch, beam, vocab, source, decode_length, alpha, temperature, inf_value, device,
                 eos_id=2, pad_id=0):
        self.B, self.K, self.V = batch, beam, vocab
        self.alpha, self.temperature, self.inf_value = float(alpha), float(temperature), float(inf_value)
        self.eos_id, self.pad_id = eos_id, pad_id
        src_len = (source != 0).sum(1)
        self.max_len = (src_len + int(decode_length)).to(torch.int32).contiguous()
        self.cap = int(self.max_len.max().item()) + 2
        # ((5 + max_len) / 6) ^ alpha in fp32 on the host, like the reference's tf.pow on a float tensor
        ml = (src_len.float().cpu() + float(decode_length))
        self.max_penalty = torch.pow((5.0 + ml) / 6.0, self.alpha).to(device)
        i32, f32 = torch.int32, torch.float32
        self.alive_seq = torch.full((batch, beam, self.cap), pad_id, dtype=i32, device=device)
        logp = torch.full((batch, beam), F32_MIN, dtype=f32)
        logp[:, 0] = 0.0
        self.alive_logp = logp.to(device)
        self.alive_score = torch.zeros(batch, beam, dtype=f32, device=device)
        self.fin_seq = torch.zeros(batch, beam, self.cap, dtype=i32, device=device)
        self.fin_score = torch.full((batch, beam), F32_MIN, dtype=f32, device=device)
        self.fin_flag = torch.zeros(batch, beam, dtype=i32, device=device)
        self.parent = torch.arange(batch * beam, dtype=i32, device=device)
        self.tmp_seq = torch.zeros(batch, 3 * beam, self.cap, dtype=i32, device=device)
        self.active = torch.ones(1, dtype=i32, device=device)
        self.time = 0

    def _args(self, logits, t):
        pen = float(torch.pow(torch.tensor((5.0 + float(t + 1)) / 6.0, dtype=torch.float32), self.alpha))
        return ops.beam_args(
            logits=logits, batch=self.B, beam=self.K, vocab=self.V, time=int(t), eos_id=self.eos_id,
            pad_id=self.pad_id, temperature=self.temperature, inf_value=self.inf_value, length_penalty=pen,
            max_len=self.max_len, max_penalty=self.max_penalty, seq_cap=self.cap, alive_seq=self.alive_seq,
            alive_logp=self.alive_logp, alive_score=self.alive_score, fin_seq=self.fin_seq,
            fin_score=self.fin_score, fin_flag=self.fin_flag, parent=self.parent, tmp_seq=self.tmp_seq,
            active=self.active)

    def not_finished(self, t):
        """search.py:85-113, evaluated on the device; one 4-byte read-back."""
        if t + 2 > self.cap:
            return False
        ops.beam_cond(self._args(None, t))
        return bool(self.active.item())

    def last_tokens(self, t):
        """[B*beam, 1] int32: the token fed to decoding_fn at step t (search.py:130)."""
        return self.alive_seq[:, :, t].reshape(self.B * self.K, 1).contiguous()

    def step(self, logits, t):
        ops.beam_step(self._args(logits, t))
        self.time = t + 1
        return self.parent

    def result(self):
        """search.py:258-275: finished beams where any exist, alive beams otherwise; BOS column dropped."""
        t = self.time
        any_fin = (self.fin_flag != 0).any(1)
        seq = torch.where(any_fin[:, None, None], self.fin_seq[:, :, :t + 1], self.alive_seq[:, :, :t + 1])
        score = torch.where(any_fin[:, None], self.fin_score, self.alive_score)
        return {"seq": seq[:, :, 1:], "score": score}


def beam_search(features, encoding_fn, decoding_fn, params):
    """Drop-in for reference search.beam_search (search.py:19).  `encoding_fn(source) -> state`,
    `decoding_fn(target [B*beam,1], state, time) -> (logits fp32 [B*beam,V], state)` as returned by
    zero_b200's infer_fn; `state.reorder(parent)` replaces the gather_nd over the tiled state."""
    source = features["source"]
    state = encoding_fn(source)
    dev = state.device
    src = torch.as_tensor(source).to(dev)
    B = src.shape[0]
    K = int(params.beam_size)
    state.begin_search(K)
    st = BeamState(B, K, state.vocab, src, params.decode_length, params.decode_alpha,
                   getattr(params, "beam_search_temperature", 1.0), getattr(params, "dtype_inf", 1e8), dev,
                   eos_id=params.tgt_vocab.eos(), pad_id=params.tgt_vocab.pad())
    t = 0
    while st.not_finished(t):
        logits, state = decoding_fn(st.last_tokens(t), state, t)
        parent = st.step(logits, t)
        state.reorder(parent, t)
        t += 1
    return st.result()
