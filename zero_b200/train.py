"""Training step around the engine: what main.tower_train_graph + parallel.average_gradients +
cycle.create_train_op do in the reference (main.py:22-45, utils/parallel.py:134-208, utils/cycle.py:47-135).

One process per GPU.  Every rank runs forward/backward on its own batch (main.py:268-273 feeds one batch per
tower), the flat fp32 gradient arena is summed with ONE NCCL all-reduce over NVLink (the reference's per-variable
concat + reduce_mean, utils/parallel.py:184-196), and the 1/world average, 1/loss_scale and the
clip_by_global_norm factor are folded into the fused TF-semantics Adam kernel.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from . import ops
from .engine import Engine

f32 = torch.float32


def noam_lr(step, init_lr, warmup_steps, hidden_size, min_lr=0.0, max_lr=1.0):
    """lrs/noamlr.py:28-36 (+ the clamp of lrs/lr.py)."""
    step = float(step)
    decay = float(hidden_size) ** -0.5 * min((step + 1) * float(warmup_steps) ** -1.5, (step + 1) ** -0.5)
    return max(min(init_lr * decay, max_lr), min_lr)


class Trainer(object):
    def __init__(self, engine: Engine, hp, world_size=1, use_graph=False, side_stream=True):
        self.eng = engine
        if side_stream:
            engine.enable_side_stream(True)
        self.hp = hp
        self.world = int(world_size)
        n = engine.ps.total
        dev = engine.device
        engine.ps.adam_m = torch.zeros(n, dtype=f32, device=dev)
        engine.ps.adam_v = torch.zeros(n, dtype=f32, device=dev)
        self.clip_scale = torch.ones(1, dtype=f32, device=dev)    # clip_by_global_norm factor, device side
        self.norms = torch.zeros(2, dtype=f32, device=dev)        # {sum g^2, sum p^2}
        self.global_step = 0
        self.beta1, self.beta2, self.eps = float(hp.beta1), float(hp.beta2), float(hp.epsilon)
        clip = getattr(hp, "clip_grad_norm", 0.0)
        self.clip = float(clip) if isinstance(clip or None, float) else None   # utils/cycle.py:98
        self.loss_scale = float(getattr(hp, "loss_scale", 1.0))
        self.use_graph = use_graph
        self._graph = None
        self._static = None

    # ------------------------------------------------------------------------------------------ lr
    def lr(self):
        hp = self.hp
        if getattr(hp, "lrate_strategy", "noam") == "noam":
            return noam_lr(self.global_step, hp.lrate, hp.warmup_steps, hp.hidden_size,
                           getattr(hp, "min_lrate", 0.0), getattr(hp, "max_lrate", 1.0))
        return float(hp.lrate)

    # ------------------------------------------------------------------------------------------ step
    def _phases(self, source, target):
        """Runs phase 1 (forward + decoder backward), calls `between()` hooks via the caller, then phase 2."""
        eng = self.eng
        if not self.use_graph:
            loss = eng.forward_backward_decoder(source, target)
            return loss, eng.backward_encoder
        key = (tuple(source.shape), tuple(target.shape))
        if self._graph is None or self._static[0] != key:
            s_src = torch.empty(source.shape, dtype=torch.int32, device=eng.device)
            s_tgt = torch.empty(target.shape, dtype=torch.int32, device=eng.device)
            s_src.copy_(source)
            s_tgt.copy_(target)
            # warm-up on a side stream (allocates every workspace buffer), then capture the two phases
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                eng.forward_backward(s_src, s_tgt, compact=False)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                loss = eng.forward_backward_decoder(s_src, s_tgt, compact=False)
            with torch.cuda.graph(g2, pool=g1.pool()):
                eng.backward_encoder()
            self._graph, self._static = (g1, g2), (key, s_src, s_tgt, loss)
        _, s_src, s_tgt, loss = self._static
        s_src.copy_(source, non_blocking=True)
        s_tgt.copy_(target, non_blocking=True)
        self._graph[0].replay()
        return loss, self._graph[1].replay

    def step(self, source, target):
        """One optimizer step on this rank's batch.  Returns the device loss tensor (no host sync).
        Data parallel: the decoder-side bucket of the flat gradient arena is all-reduced (NCCL, async) while the
        encoder backward is still running; the encoder-side bucket follows; both complete before Adam."""
        eng, ps = self.eng, self.eng.ps
        loss, phase2 = self._phases(source, target)
        works = []
        if self.world > 1:
            works.append(dist.all_reduce(ps.grad[ps.dec_offset:], op=dist.ReduceOp.SUM, async_op=True))
        phase2()
        if self.world > 1:
            works.append(dist.all_reduce(ps.grad[:ps.dec_offset], op=dist.ReduceOp.SUM, async_op=True))
            for w in works:
                w.wait()
        # tf.global_norm of gradients and parameters (utils/cycle.py:94-95): a separate pass only when the clip
        # factor needs the gradient norm before the update; otherwise fused into the Adam kernel
        self.norms.zero_()
        if self.clip is not None:
            ops.sumsq(ps.grad, self.norms[0:1])
        self.global_step += 1
        t = self.global_step
        lr_t = self.lr() * math.sqrt(1.0 - self.beta2 ** t) / (1.0 - self.beta1 ** t)
        gscale = 1.0 / (self.world * self.loss_scale)
        clip_scale = None
        if self.clip is not None:
            # clip_by_global_norm: g * clip / max(norm, clip)   (device-side scalar ops, no host sync)
            gn = torch.sqrt(self.norms[0]) * gscale
            torch.div(self.clip, torch.clamp(gn, min=self.clip), out=self.clip_scale[0])
            clip_scale = self.clip_scale
            self.norms.zero_()
        ops.adam_tf(ps.master, ps.adam_m, ps.adam_v, ps.grad, ps.mirror, self.beta1, self.beta2, self.eps,
                    lr_t, gscale, clip_scale, self.norms)
        return loss

    def gradient_norm(self):
        """GNorm of main.py:336-346 (already averaged over towers and un-scaled)."""
        return float(torch.sqrt(self.norms[0]).item())

    def parameter_norm(self):
        return float(torch.sqrt(self.norms[1]).item())
