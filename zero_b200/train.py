"""Training step around the engine: what main.tower_train_graph + parallel.average_gradients +
cycle.create_train_op do in the reference (main.py:22-45, utils/parallel.py:134-208, utils/cycle.py:47-135).

One process per GPU.  Every rank runs forward/backward on its own batch (main.py:268-273 feeds one batch per
tower), the flat fp32 gradient arena is summed with ONE NCCL all-reduce over NVLink (the reference's per-variable
concat + reduce_mean, utils/parallel.py:184-196), and the 1/world average, 1/loss_scale and the
clip_by_global_norm factor are folded into the fused TF-semantics Adam kernel.
"""
from __future__ import annotations

import math
import os

import torch
import torch.distributed as dist

from . import lib as L
from . import ops
from .engine import Engine

f32 = torch.float32


def noam_lr(step, init_lr, warmup_steps, hidden_size, min_lr=0.0, max_lr=1.0):
    """lrs/noamlr.py:28-36 (+ the clamp of lrs/lr.py)."""
    step = float(step)
    decay = float(hidden_size) ** -0.5 * min((step + 1) * float(warmup_steps) ** -1.5, (step + 1) ** -0.5)
    return max(min(init_lr * decay, max_lr), min_lr)


class Trainer(object):
    """One optimizer step = `update_cycle` micro-batches per rank (utils/cycle.py:53-92: gradients and the loss are
    averaged over the cycle), one all-reduce, one fused Adam.  `lr_schedule` is a zero_b200.lrs schedule object
    (or None: Noam from the hyper-parameters, as the Transformer recipes use)."""

    MAX_GRAPHS = 24   # captured (shape, zero_grad) variants kept; further shapes run eagerly

    @staticmethod
    def bucketed(ids, multiple):
        """`ids` [B, L] zero-padded on the right to the next multiple of `multiple` columns (pad id 0: masked keys,
        zero-weight targets — the result of the step does not change, models/transformer.py:16,208-210).  Token-budget
        batches then fall into a few dozen (S, T) classes per batch size instead of a new shape every step, which is
        what lets captured graphs be replayed (opt-in: ZB_GRAPH_BUCKET=8 with use_graph)."""
        if multiple <= 1:
            return ids
        cols = ids.shape[1]
        want = (cols + multiple - 1) // multiple * multiple
        if want == cols:
            return ids
        out = ids.new_zeros((ids.shape[0], want))
        out[:, :cols] = ids
        return out

    def __init__(self, engine: Engine, hp, world_size=1, use_graph=False, side_stream=True, lr_schedule=None,
                 early_adam=False, shard_transport=None):
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
        self.loss_acc = torch.zeros(1, dtype=f32, device=dev)     # sum of the cycle's micro-batch losses
        self.global_step = 0
        self.beta1, self.beta2, self.eps = float(hp.beta1), float(hp.beta2), float(hp.epsilon)
        clip = getattr(hp, "clip_grad_norm", 0.0)
        self.clip = float(clip) if isinstance(clip or None, float) else None   # utils/cycle.py:98
        self.loss_scale = float(getattr(hp, "loss_scale", 1.0))
        self.cycle = max(1, int(getattr(hp, "update_cycle", 1)))
        self.lr_schedule = lr_schedule
        self.use_graph = use_graph
        self._graphs = {}
        self._graphs_gen = engine.ws.generation
        self.bucket = int(os.environ.get("ZB_GRAPH_BUCKET", "0") or 0)
        if self.bucket > 1:
            self.MAX_GRAPHS = 512      # the workspace is shared between shapes: a captured shape costs no buffers
        self._micro = 0
        self._pending = None
        # opt-in (ZB_SHARD_OPT=1; =p2p: without the multicast mappings): gradient aggregation + Adam + refresh of every
        # rank's compute copy as ONE kernel per rank over NVLink peer memory (zero_b200/shard_opt.py) instead of the
        # NCCL all-reduce + replicated Adam below.  safe_nan needs the summed gradients on the host BEFORE the update
        # (main.py:320-332), which the fused step never materialises: that mode keeps the all-reduce.
        self.shard = None
        # ZB_SHARD_OPT: "auto" (default) = the fused step, over unicast peer mappings on two GPUs and over the NVSwitch
        # multicast mapping from three on (measured at N = 2: 3.87 ms/step against 4.01 for the bucketed all-reduce; at
        # N = 8: 4.19 against 4.34 — profiles/r02_scaling_ab.log); "0" = NCCL all-reduce + replicated Adam; "1" / "p2p"
        # force a transport.  If the symmetric-memory set-up fails on ANY rank, every rank falls back to the all-reduce.
        mode = os.environ.get("ZB_SHARD_OPT", "auto")
        if mode == "auto":
            mode = "0" if self.world <= 1 else ("p2p" if self.world == 2 else "1")
        # ZB_SHARD_OVERLAP=1 (opt-in): two regions, the decoder side reduced / updated on a second stream UNDER the encoder
        # backward.  Validated (tests/test_shard_opt_gpu.py) and measured SLOWER than the flat step — 3.93 vs 3.85 ms/step
        # at N = 2, 4.30 vs 4.20 at N = 8 (profiles/r02_scaling_ab.log): the concurrent kernel costs the backward's GEMMs
        # more than the communication it hides, the same outcome as overlapping Adam (ZB_EARLY_ADAM) in round 1.
        self._shard_split = engine.ps.dec_offset if (self.clip is None and os.environ.get("ZB_SHARD_OVERLAP", "0") == "1"
                                                     and 0 < engine.ps.dec_offset < engine.ps.total) else None
        self._early = None
        if shard_transport is not None:
            from .shard_opt import ShardedStep
            self.shard = ShardedStep(engine, shard_transport, use_multicast=mode != "p2p", split=self._shard_split)
        elif self.world > 1 and mode in ("1", "p2p") and not bool(getattr(hp, "safe_nan", False)):
            self.shard = self._try_sharded_step(engine, mode)
        # exponential moving average of the parameters (utils/cycle.py:114-127), off unless ema_decay > 0
        self.ema_decay = float(getattr(hp, "ema_decay", -1.0))
        self.ema = engine.ps.master.clone() if self.ema_decay > 0.0 else None
        self._ema_backup = None
        # split optimizer step (see _step_split); ZB_EARLY_ADAM=0 keeps the single fused pass after the backward
        self.early_adam = (early_adam or os.environ.get("ZB_EARLY_ADAM", "0") == "1") and self.shard is None
        self._opt_stream = None
        # Data parallel without the sharded step: the encoder backward runs in `enc_groups` groups of layers (last
        # layers first) and every group's slice of the gradient arena is all-reduced while the next group's backward
        # runs; the source embedding + shared bias, whose gradients complete last, go out as their own bucket and are
        # reduced WHILE Adam already updates everything else (no global quantity is needed when clipping is off).
        # ZB_ENC_BUCKETS=1 restores the two-bucket scheme of round 1.
        # (measured: three groups win on two GPUs, 4.01 vs 4.11 ms/step; at N = 8 NCCL's per-call cost makes the
        # two-bucket scheme the faster one, 4.34 vs 4.40 ms/step)
        groups = os.environ.get("ZB_ENC_BUCKETS") or ("3" if self.world == 2 else "1")
        self.enc_groups = int(groups) if (self.world > 1 and self.shard is None) else 1
        self._late = None
        # NCCL's all-reduce kernels run next to the backward pass: the persistent GEMM / attention grids leave them a
        # few SMs instead of being split into two rounds (csrc/abi.cu num_sms_compute)
        # (measured at N = 8: no gain from reserving 8 / 16 / 32 SMs — 4.34 / 4.36 / 4.37 / 4.52 ms/step — so the default
        # is 0; ZB_SM_RESERVE keeps the experiment reproducible)
        if self.world > 1 and self.shard is None and os.environ.get("ZB_SM_RESERVE"):
            L.check(L.load().zb_set_sm_reserve(int(os.environ["ZB_SM_RESERVE"])), "zb_set_sm_reserve")

    def _try_sharded_step(self, engine, mode):
        """ShardedStep over symmetric memory, or None (on every rank) when any rank could not set it up."""
        from .shard_opt import ShardedStep, SymmMemTransport
        if not (dist.is_available() and dist.is_initialized()):
            return None
        shard, err = None, None
        keep = {name: getattr(engine.ps, name) for name in ("grad", "mirror", "master")}
        try:
            shard = ShardedStep(engine, SymmMemTransport(), use_multicast=mode != "p2p", split=self._shard_split)
        except Exception as e:     # noqa: BLE001 — whatever the plumbing raises, the all-reduce path still works
            err = e
        ok = torch.tensor([0 if shard is None else 1], dtype=torch.int32, device=engine.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            return shard
        if dist.get_rank() == 0:
            print("zero_b200: sharded optimizer step unavailable (%r); using the NCCL all-reduce" % (err,), flush=True)
        for name, t in keep.items():      # undo the rebinding of the arenas a partial set-up may have done
            if getattr(engine.ps, name) is not t:
                t.copy_(getattr(engine.ps, name))
                setattr(engine.ps, name, t)
        return None

    # ------------------------------------------------------------------------------------------ lr
    def lr(self):
        hp = self.hp
        if self.lr_schedule is not None:
            self.lr_schedule.step(self.global_step)
            return float(self.lr_schedule.get_lr())
        if getattr(hp, "lrate_strategy", "noam") == "noam":
            return noam_lr(self.global_step, hp.lrate, hp.warmup_steps, hp.hidden_size,
                           getattr(hp, "min_lrate", 0.0), getattr(hp, "max_lrate", 1.0))
        return float(hp.lrate)

    # ------------------------------------------------------------------------------------------ step
    def _phases(self, source, target, zero_grad=True):
        """Runs phase 1 (forward + decoder backward) and returns (loss, stages): the encoder backward as a list of
        callables, one per group of layers (Engine.encoder_buckets order), the last one including the embedding."""
        eng = self.eng
        eng.advance_dropout_seed()
        if self.use_graph and self.bucket > 1:
            source, target = self.bucketed(torch.as_tensor(source), self.bucket), \
                self.bucketed(torch.as_tensor(target), self.bucket)
        key = (tuple(source.shape), tuple(target.shape), bool(zero_grad))
        graphable = self.use_graph and source.shape[0] > 0      # an empty tower launches nothing (engine guard)
        if self._graphs_gen != eng.ws.generation:
            # a workspace buffer was replaced by a larger one since the capture: the graphs' pointers are stale
            self._graphs.clear()
            self._graphs_gen = eng.ws.generation
        entry = self._graphs.get(key) if graphable else None
        if graphable and entry is None and len(self._graphs) < self.MAX_GRAPHS:
            s_src = torch.empty(source.shape, dtype=torch.int32, device=eng.device)
            s_tgt = torch.empty(target.shape, dtype=torch.int32, device=eng.device)
            s_src.copy_(source)
            s_tgt.copy_(target)
            # warm-up on a side stream (allocates every workspace buffer), then capture the two phases; the warm-up
            # must not disturb gradients that a running accumulation cycle has already collected
            keep = eng.ps.grad.clone() if not zero_grad else None
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                eng.forward_backward(s_src, s_tgt, compact=False)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if keep is not None:
                eng.ps.grad.copy_(keep)
            if self._graphs_gen != eng.ws.generation:
                # the warm-up grew a workspace buffer: graphs captured for earlier (smaller) batches point into the
                # allocation it replaced
                self._graphs.clear()
                self._graphs_gen = eng.ws.generation
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                loss = eng.forward_backward_decoder(s_src, s_tgt, zero_grad=zero_grad, compact=False)
            g2 = []
            for stop in self._stage_stops():
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=g1.pool()):
                    eng.backward_encoder(stop_layer=stop)
                g2.append(g)
            if keep is not None:
                eng.ps.grad.copy_(keep)   # capture does not execute, but stay safe against future changes
            entry = self._graphs[key] = (g1, g2, s_src, s_tgt, loss)
        if entry is None:
            loss = eng.forward_backward_decoder(source, target, zero_grad=zero_grad,
                                                compact=not self.use_graph)
            return loss, [(lambda stop=stop: eng.backward_encoder(stop_layer=stop)) for stop in self._stage_stops()]
        g1, g2, s_src, s_tgt, loss = entry
        s_src.copy_(source, non_blocking=True)
        s_tgt.copy_(target, non_blocking=True)
        g1.replay()
        return loss, [g.replay for g in g2]

    def _stage_stops(self):
        """stop_layer of every encoder-backward stage (the last stage runs to layer 0 and the embedding)."""
        return [b[0] for b in self.eng.encoder_buckets(self.enc_groups)[:-1]]

    def step(self, source, target):
        """One micro-batch.  Returns the device loss tensor [1] of this micro-batch (no host sync).  The parameters
        are updated when the call completes an `update_cycle` (every call for update_cycle = 1).
        Data parallel: the decoder-side bucket of the flat gradient arena is all-reduced (NCCL, async) while the
        encoder backward is still running; the encoder-side bucket follows; both complete before Adam."""
        if self.early_adam and self.clip is None and self._micro == self.cycle - 1:
            return self._step_split(source, target)
        loss = self.compute(source, target)
        if self._pending:
            self.apply()
        return loss

    def _step_split(self, source, target):
        """Last micro-batch of a cycle without gradient clipping: the optimizer needs no global quantity, so the
        decoder-side parameters (their gradients are final after phase 1) are all-reduced and updated on a second
        stream WHILE the encoder backward runs; the encoder side follows.
        MEASURED (B200, config 2): slower than the single fused pass — 3.82 vs 3.73 ms/step at N = 1 and 4.31 vs
        4.19 ms at N = 2 — the Adam blocks delay the critical-path kernels more than they hide; off by default
        (Trainer(early_adam=True) / the ab_bench tools keep the experiment reproducible)."""
        eng, ps = self.eng, self.eng.ps
        first = self._micro == 0
        loss, stages = self._phases(source, target, zero_grad=first)

        def phase2():
            for st in stages:
                st()
        if self.cycle > 1:
            if first:
                self.loss_acc.zero_()
            self.loss_acc += loss
        self._micro = 0
        self._pending = False
        lr = self.lr()
        self.global_step += 1
        t = self.global_step
        lr_t = lr * math.sqrt(1.0 - self.beta2 ** t) / (1.0 - self.beta1 ** t)
        gscale = 1.0 / (self.world * self.cycle * self.loss_scale)
        self.norms.zero_()
        main = torch.cuda.current_stream()
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream(device=eng.device)
        opt = self._opt_stream
        opt.wait_stream(main)                      # phase 1 (and the zeroing of `norms`) precede the early update
        off = ps.dec_offset
        with torch.cuda.stream(opt):
            if self.world > 1:
                dist.all_reduce(ps.grad[off:], op=dist.ReduceOp.SUM, async_op=True).wait()
            ops.adam_tf(ps.master[off:], ps.adam_m[off:], ps.adam_v[off:], ps.grad[off:], ps.mirror[off:],
                        self.beta1, self.beta2, self.eps, lr_t, gscale, None, self.norms)
        phase2()
        if self.world > 1:
            dist.all_reduce(ps.grad[:off], op=dist.ReduceOp.SUM, async_op=True).wait()
        main.wait_stream(opt)
        ops.adam_tf(ps.master[:off], ps.adam_m[:off], ps.adam_v[:off], ps.grad[:off], ps.mirror[:off],
                    self.beta1, self.beta2, self.eps, lr_t, gscale, None, self.norms)
        self._ema_update()
        return loss

    def compute(self, source, target):
        """Forward + backward (+ all-reduce on the last micro-batch of the cycle); no parameter change.
        `self.cycle_ready()` tells whether `apply()` is due; `gradient_norm(ready=False)` reads GNorm first
        (safe_nan mode, main.py:320-332)."""
        eng, ps = self.eng, self.eng.ps
        first = self._micro == 0
        last = self._micro == self.cycle - 1
        loss, stages = self._phases(source, target, zero_grad=first)
        if self.cycle > 1:
            if first:
                self.loss_acc.zero_()
            self.loss_acc += loss
        works = []
        reduce_now = last and self.world > 1 and self.shard is None   # sharded step: the sum is taken in apply()
        if last and self.shard is not None and self.shard.split:
            self._shard_early()
        buckets = eng.encoder_buckets(self.enc_groups)
        if reduce_now:
            works.append(dist.all_reduce(ps.grad[ps.dec_offset:], op=dist.ReduceOp.SUM, async_op=True))
        for (stop, lo, hi), stage in zip(buckets[:-1], stages):
            stage()
            if reduce_now and self.enc_groups > 1:
                works.append(dist.all_reduce(ps.grad[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
        if reduce_now:
            if self.enc_groups > 1:
                # source embedding + shared bias: reduced while apply() updates the rest (when nothing global is needed)
                late = dist.all_reduce(ps.grad[:buckets[-1][2]], op=dist.ReduceOp.SUM, async_op=True)
                if self.clip is None and not bool(getattr(self.hp, "safe_nan", False)):
                    self._late = (late, buckets[-1][2])
                else:
                    works.append(late)
            else:
                works.append(dist.all_reduce(ps.grad[:ps.dec_offset], op=dist.ReduceOp.SUM, async_op=True))
            for w in works:
                w.wait()
        self._micro = 0 if last else self._micro + 1
        self._pending = last
        return loss

    def _step_scalars(self):
        """(lr_t, gscale) of the update that completes this cycle (utils/cycle.py:94-105; TF Adam's bias-corrected rate)."""
        t = self.global_step + 1
        lr_t = self.lr() * math.sqrt(1.0 - self.beta2 ** t) / (1.0 - self.beta1 ** t)
        return lr_t, 1.0 / (self.world * self.cycle * self.loss_scale)

    def _shard_early(self):
        """Decoder-side region of the fused step on a second stream, under the encoder backward (shard_opt.step_early)."""
        self._early = self._step_scalars()
        if self.eng.device.type != "cuda":           # host-only tests: no streams
            self.shard.step_early(self, *self._early)
            return
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream(device=self.eng.device)
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._opt_stream):
            self._opt_stream.wait_event(ev)
            self.shard.step_early(self, *self._early)

    def cycle_ready(self):
        return bool(self._pending)

    def cycle_loss(self):
        """Mean loss of the finished cycle (utils/cycle.py:90-92), device tensor."""
        return self.loss_acc / float(self.cycle)

    def mean_over_ranks(self, loss):
        """The tower-averaged loss of main.py:42 (one 4-byte all-reduce): every rank sees the same value, so the
        NaN / Inf decisions taken on it (main.py:316-332) agree across the job."""
        if self.world <= 1:
            return loss
        out = loss.detach().clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
        return out / float(self.world)

    def skip(self):
        """Drop the collected gradients without updating (safe_nan, main.py:326-330: the step is 'passed')."""
        self._pending = False
        if self._late is not None:
            self._late[0].wait()
            self._late = None

    def apply(self):
        """clip_by_global_norm + TF Adam on the averaged gradients (utils/cycle.py:94-105) + EMA."""
        ps = self.eng.ps
        self._pending = False
        if self.shard is not None:
            lr_t, gscale = self._early if self._early is not None else self._step_scalars()
            if self._early is not None and self._opt_stream is not None:
                torch.cuda.current_stream().wait_stream(self._opt_stream)
            self._early = None
            self.global_step += 1
            self.shard.step(self, lr_t, gscale)
            self._ema_update()
            return
        # tf.global_norm of gradients and parameters (utils/cycle.py:94-95): a separate pass only when the clip
        # factor needs the gradient norm before the update; otherwise fused into the Adam kernel
        self.norms.zero_()
        if self.clip is not None:
            ops.sumsq(ps.grad, self.norms[0:1])
        lr = self.lr()
        self.global_step += 1
        t = self.global_step
        lr_t = lr * math.sqrt(1.0 - self.beta2 ** t) / (1.0 - self.beta1 ** t)
        gscale = 1.0 / (self.world * self.cycle * self.loss_scale)
        clip_scale = None
        if self.clip is not None:
            # clip_by_global_norm: g * clip / max(norm, clip)   (device-side scalar ops, no host sync)
            gn = torch.sqrt(self.norms[0]) * gscale
            torch.div(self.clip, torch.clamp(gn, min=self.clip), out=self.clip_scale[0])
            clip_scale = self.clip_scale
            self.norms.zero_()
        if self._late is not None:
            work, cut = self._late
            self._late = None
            ops.adam_tf(ps.master[cut:], ps.adam_m[cut:], ps.adam_v[cut:], ps.grad[cut:], ps.mirror[cut:], self.beta1,
                        self.beta2, self.eps, lr_t, gscale, clip_scale, self.norms)
            work.wait()
            ops.adam_tf(ps.master[:cut], ps.adam_m[:cut], ps.adam_v[:cut], ps.grad[:cut], ps.mirror[:cut], self.beta1,
                        self.beta2, self.eps, lr_t, gscale, clip_scale, self.norms)
        else:
            ops.adam_tf(ps.master, ps.adam_m, ps.adam_v, ps.grad, ps.mirror, self.beta1, self.beta2, self.eps,
                        lr_t, gscale, clip_scale, self.norms)
        self._ema_update()

    def _ema_update(self):
        if self.ema is not None:
            # tf.train.ExponentialMovingAverage(decay, num_updates=global_step): decay' = min(decay, (1+n)/(10+n))
            n = float(self.global_step)
            d = min(self.ema_decay, (1.0 + n) / (10.0 + n))
            if self.shard is not None:   # only the own shards of the master are current
                for lo, n in self.shard.ranges:
                    self.ema[lo:lo + n].lerp_(self.eng.ps.master[lo:lo + n], 1.0 - d)
            else:
                self.ema.lerp_(self.eng.ps.master, 1.0 - d)

    def sync_full_state(self):
        """Collective, no-op unless the optimizer step is sharded: every rank's fp32 master, Adam slots and EMA shadow
        become whole again (before a checkpoint is written or the averaged parameters are swapped in)."""
        if self.shard is not None:
            self.shard.sync_full_state(extra=() if self.ema is None else (self.ema,))

    # ------------------------------------------------------------------------------------------ EMA swap (eval)
    def ema_assign(self):
        """ema_backup_op + ema_assign_op (main.py:368-370): evaluate with the averaged parameters."""
        if self.ema is None:
            return
        self.sync_full_state()
        ps = self.eng.ps
        self._ema_backup = ps.master.clone()
        ps.master.copy_(self.ema)
        ps.refresh_mirror()

    def ema_restore(self):
        if self.ema is None or self._ema_backup is None:
            return
        ps = self.eng.ps
        ps.master.copy_(self._ema_backup)
        self._ema_backup = None
        ps.refresh_mirror()

    def gradient_norm(self, before_apply=False):
        """GNorm of main.py:336-346 (averaged over towers / cycle, un-scaled).  With before_apply=True the norm of
        the collected, not yet applied gradients is computed by a separate pass (safe_nan mode)."""
        if before_apply:
            if self.shard is not None:
                raise L.ZeroB200Error("the sharded optimizer step never materialises the summed gradients")
            tmp = torch.zeros(1, dtype=f32, device=self.eng.device)
            ops.sumsq(self.eng.ps.grad, tmp)
            return float(torch.sqrt(tmp[0]).item()) / (self.world * self.cycle * self.loss_scale)
        return float(torch.sqrt(self._norms()[0]).item())

    def parameter_norm(self):
        return float(torch.sqrt(self._norms()[1]).item())

    def _norms(self):
        return self.norms if self.shard is None else self.shard.norms()
