"""Host logic of the sharded optimizer step (zero_b200/shard_opt.py, ZB_SHARD_OPT=1) on the CPU.

The product path is one CUDA kernel per rank over NVLink peer memory (zb_shard_adam) bracketed by symmetric-memory
barriers.  Here N ranks are N threads of one process: the transport is an in-process stand-in (shared CPU tensors as
"peer memory", threading.Barrier as the cross-rank barrier) and the kernel entry point is replaced by a torch
emulation of the semantics include/zero_b200.h documents.  What is checked is everything AROUND the kernel — shard
plan, arena rebinding, argument plumbing (addresses, offsets, flags), the two-pass clip flow, the norm exchange,
EMA on the own shard, whole-state sync — against the oracle's replicated Adam on the mean gradient
(utils/parallel.py:184-196 + main.py:178-181).  The kernel itself is the GPU tests' job."""
import threading

import pytest
import torch

from oracle import zero_oracle as zo
from zero_b200 import lib as L
from zero_b200 import shard_opt
from zero_b200.params import transformer_base

f32, bf16 = torch.float32, torch.bfloat16


# ------------------------------------------------------------------------------------------------ stand-ins
class World(object):
    """What N processes on one NVSwitch box share: symmetric allocations and a barrier."""

    def __init__(self, n):
        self.n = n
        self.bar = threading.Barrier(n)
        self.lock = threading.Lock()
        self.sym = []          # allocation index -> [tensor of rank 0, ..., tensor of rank n-1]
        self.addr = {}         # data_ptr -> tensor, for everything a raw address may point at
        self.box = {}
        self.barriers = [0] * n

    def register(self, t):
        with self.lock:
            self.addr[t.data_ptr()] = t

    def view(self, ptr, i, n):
        """Elements [i, i + n) of the array whose element 0 sits at raw address `ptr` (which may lie before the
        tensor that backs it: arena-indexed bases)."""
        with self.lock:
            for base, t in self.addr.items():
                d = ptr + i * t.element_size() - base
                if d % t.element_size() == 0 and 0 <= d and d + n * t.element_size() <= t.numel() * t.element_size():
                    return t[d // t.element_size(): d // t.element_size() + n]
        raise KeyError((ptr, i, n))


class FakeTransport(object):
    def __init__(self, world, rank):
        self.w, self.world, self.rank = world, world.n, rank
        self._next = 0

    def alloc(self, n, dtype, device):
        with self.w.lock:
            i = self._next
            self._next += 1
            if len(self.w.sym) <= i:
                self.w.sym.append([torch.zeros(n, dtype=dtype) for _ in range(self.world)])
                for t in self.w.sym[i]:
                    self.w.addr[t.data_ptr()] = t
        return self.w.sym[i][self.rank]

    def rendezvous(self, t):
        for bufs in self.w.sym:
            if bufs[self.rank] is t:
                return [b.data_ptr() for b in bufs], 0
        raise AssertionError("not a symmetric allocation")

    def barrier(self):
        self.w.barriers[self.rank] += 1
        self.w.bar.wait()

    def broadcast(self, t, src):
        key = ("bc", self.w.barriers[self.rank], t.numel(), src)
        if self.rank == src:
            self.w.box[key] = t.clone()
        self.w.bar.wait()
        if self.rank != src:
            t.copy_(self.w.box[key])
        self.w.bar.wait()


def _fake_shard_adam(world):
    """torch emulation of zb_shard_adam (include/zero_b200.h, K10) on raw addresses."""

    at = world.view

    def fn(lo, n, nranks, rank, grad_ptrs, mirror_ptrs, param, m, v, beta1, beta2, eps, lr_t, grad_scale, flags=0,
           grad_mc=0, mirror_mc=0, grad_out=None, clip_scale=None, norms=None, norm_parts_ptrs=None,
           done_counter=None, wide_mask=None, param_ptrs=None, param_mc=0, grad_sources=None):
        assert lo % 8 == 0 and n % 8 == 0 and not grad_mc and not mirror_mc and not param_mc
        src = nranks if grad_sources is None else grad_sources
        g = torch.zeros(n)
        for r in range(src):
            g = g + at(grad_ptrs[r], lo, n)
        if flags & L.ZB_SHARD_STORE_GRAD:
            at(grad_out, lo, n).copy_(g)
        sg = ((g * grad_scale) ** 2).sum()
        sp = torch.zeros(())
        if flags & L.ZB_SHARD_UPDATE:
            gs = grad_scale * (float(clip_scale[0]) if clip_scale is not None else 1.0)
            P, M, V = param[lo:lo + n], m[lo:lo + n], v[lo:lo + n]
            sp = (P * P).sum()
            gr = g * gs
            M.mul_(beta1).add_(gr, alpha=1 - beta1)
            V.mul_(beta2).add_(gr * gr, alpha=1 - beta2)
            P.sub_(lr_t * M / (V.sqrt() + eps))
            for r in range(nranks):
                at(mirror_ptrs[r], lo, n).copy_(P.to(bf16))
            if wide_mask is not None:
                sel = wide_mask[lo // 64:(lo + n) // 64].bool().repeat_interleave(64)
                for r in range(nranks):
                    if r != rank:
                        at(param_ptrs[r], lo, n)[sel] = P[sel]
        if norms is not None:
            if flags & L.ZB_SHARD_NORM_G:
                norms[0] += sg
            if flags & L.ZB_SHARD_NORM_P:
                norms[1] += sp
            if norm_parts_ptrs:
                for r in range(nranks):
                    at(norm_parts_ptrs[r], 2 * rank, 2).copy_(norms)
    return fn


def _hp(**over):
    return transformer_base(hidden_size=64, embed_size=64, filter_size=128, num_heads=2, num_encoder_layer=1,
                            num_decoder_layer=1, lrate=1.0, warmup_steps=10, beta1=0.9, beta2=0.98, epsilon=1e-8,
                            **over)


@pytest.fixture
def host_only(monkeypatch):
    import zero_b200.ops as ops
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(ops, "cast_f32_bf16", lambda src, dst: dst.copy_(src))


def _run_ranks(n, fn):
    errs = []

    def wrap(r):
        try:
            fn(r)
        except BaseException as e:       # noqa: BLE001 — re-raised in the main thread
            errs.append(e)
            raise
    ts = [threading.Thread(target=wrap, args=(r,)) for r in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(120)
    if errs:
        raise errs[0]


def _setup(world_size, monkeypatch, hp):
    import zero_b200.engine as E
    import zero_b200.ops as ops
    from zero_b200.train import Trainer
    world = World(world_size)
    monkeypatch.setattr(ops, "shard_adam", _fake_shard_adam(world))
    engines, trainers = [], []
    for r in range(world_size):
        eng = E.Engine(hp, 40, 40, device="cpu")
        eng.ps.init_random(7)                          # replicas start identical
        engines.append(eng)
    # constructors allocate symmetric memory collectively: build them in lock-step like N processes would
    slots = [None] * world_size

    def build(r):
        slots[r] = Trainer(engines[r], hp, world_size=world_size, use_graph=False, side_stream=False,
                           shard_transport=FakeTransport(world, r))
    _run_ranks(world_size, build)
    for tr in slots:
        tr.shard._reduced_base()                       # the clip flow's local buffer, so that its address resolves
        for t in (tr.eng.ps.adam_m, tr.eng.ps.adam_v, tr.norms, tr.shard.reduced):
            world.register(t)
    return world, engines, slots


# ------------------------------------------------------------------------------------------------ tests
def test_shard_plan_tiles_the_arena():
    for total, world in ((64 * 1000, 8), (64 * 7, 3), (64 * 2, 4), (76907008, 8), (64, 1)):
        plan = shard_opt.plan_shards(total, world)
        assert len(plan) == world
        pos = 0
        for lo, n in plan:
            assert lo == pos and lo % 64 == 0 and n % 64 == 0 and n >= 0
            pos += n
        assert pos == total
        assert max(n for _, n in plan) - min(n for _, n in plan if n) <= max(n for _, n in plan)   # equal but the tail
    assert shard_opt.plan_shards(76907008, 8)[0] == (0, 9613376)


def test_wide_mask_marks_exactly_the_fp32_read_variables(host_only):
    import zero_b200.engine as E
    eng = E.Engine(_hp(), 40, 40, device="cpu")
    ps = eng.ps
    mask = shard_opt.wide_slot_mask(ps)
    assert mask.numel() == ps.total // 64
    marks = torch.zeros(ps.total, dtype=torch.int32)
    for name, (off, shape) in ps.slots.items():
        if len(shape) == 1:
            marks[off:off + shape[0]] = 1
    # every element of a 1-D variable lies in a marked slot; no element of a matrix does
    per_elem = mask.repeat_interleave(64).int()
    assert bool((per_elem >= marks).all())
    for name, (off, shape) in ps.slots.items():
        if len(shape) == 2:
            size = shape[0] * shape[1]
            assert int(per_elem[off:off + size].sum()) == 0, name


@pytest.mark.parametrize("world_size,clip", [(2, 0.0), (3, 0.0), (3, 0.05), (4, 1e9)])
def test_sharded_step_equals_replicated_adam_on_the_mean_gradient(world_size, clip, monkeypatch, host_only):
    hp = _hp(clip_grad_norm=clip, ema_decay=0.9)
    monkeypatch.setenv("ZB_SHARD_OVERLAP", "1")     # the two-region plan wherever clipping allows it
    world, engines, trainers = _setup(world_size, monkeypatch, hp)
    total = engines[0].ps.total
    assert (trainers[0].clip is None) == (clip == 0.0)
    p_ref = engines[0].ps.master.clone()
    m_ref, v_ref = torch.zeros(total), torch.zeros(total)
    ema_ref = p_ref.clone()
    gen = torch.Generator().manual_seed(3)
    for step in (1, 2, 3):
        grads = [torch.randn(total, generator=gen) * 0.01 for _ in range(world_size)]
        for r in range(world_size):
            engines[r].ps.grad.copy_(grads[r])
            assert engines[r].ps.grad is world.sym[0][r]            # the arena moved into "symmetric memory"
        stale_before = [engines[r].ps.master.clone() for r in range(world_size)]

        def one(r):
            trainers[r]._pending = True
            if clip == 0.0 and step >= 2:
                # what Trainer.compute does after the decoder backward: the decoder-side region goes first (on the GPU
                # on a second stream, under the encoder backward), apply() then finishes with the encoder side
                trainers[r]._shard_early()
                assert trainers[r].shard._early_done
            trainers[r].apply()
            assert not trainers[r].shard._early_done
        _run_ranks(world_size, one)

        g = sum(grads) / world_size                                  # utils/parallel.py:196: mean over the towers
        gn = float(g.norm())
        if clip > 0.0:
            g_used = g * (clip / max(gn, clip))                      # tf.clip_by_global_norm
        else:
            g_used = g
        p_before = p_ref.clone()
        lr = trainers[0].lr_schedule.get_lr() if trainers[0].lr_schedule is not None else None
        from zero_b200.train import noam_lr
        lr = noam_lr(step - 1, hp.lrate, hp.warmup_steps, hp.hidden_size, 0.0, 1.0)
        p_ref, m_ref, v_ref = zo.adam_tf_step(p_ref, m_ref, v_ref, g_used, step, lr, 0.9, 0.98, 1e-8)
        d = min(0.9, (1.0 + step) / (10.0 + step))
        ema_ref = ema_ref + (p_ref - ema_ref) * (1.0 - d)
        wide = shard_opt.wide_slot_mask(engines[0].ps).bool().repeat_interleave(64)
        # the arena is sharded as one region (clipping) or two (decoder side / encoder side): assemble the owners' state
        owned = torch.zeros(total)
        for plan in trainers[0].shard.plans:
            for r, (lo, n) in enumerate(plan):
                owned[lo:lo + n] = engines[r].ps.master[lo:lo + n]
        assert (len(trainers[0].shard.plans) == 2) == (clip == 0.0)
        for r in range(world_size):
            ps, tr = engines[r].ps, trainers[r]
            mine = torch.zeros(total, dtype=torch.bool)
            for lo, n in tr.shard.ranges:
                mine[lo:lo + n] = True
            assert tr.global_step == step
            # every rank's compute copy is whole and identical; own shard of the fp32 state is current
            assert torch.equal(ps.mirror, owned.to(bf16)) and torch.equal(ps.mirror, engines[0].ps.mirror)
            torch.testing.assert_close(ps.mirror.float(), p_ref, atol=1e-2, rtol=1e-2)
            torch.testing.assert_close(ps.master[mine], p_ref[mine], atol=1e-6, rtol=1e-5)
            torch.testing.assert_close(ps.adam_m[mine], m_ref[mine], atol=1e-7, rtol=1e-5)
            # fp32-read variables are current everywhere, the rest of the foreign master is untouched (stale)
            torch.testing.assert_close(ps.master[wide], p_ref[wide], atol=1e-6, rtol=1e-5)
            foreign = ~mine
            assert torch.equal(ps.master[foreign & ~wide], stale_before[r][foreign & ~wide])
            # tf.global_norm of the averaged gradients / pre-update parameters, no extra collective
            assert abs(tr.gradient_norm() - gn) < 1e-4 * max(gn, 1.0)
            assert abs(tr.parameter_norm() - float(p_before.norm())) < 1e-3
        _run_ranks(world_size, lambda r: trainers[r].sync_full_state())
        for r in range(world_size):
            ps, tr = engines[r].ps, trainers[r]
            torch.testing.assert_close(ps.master, p_ref, atol=1e-6, rtol=1e-5)
            torch.testing.assert_close(ps.adam_v, v_ref, atol=1e-9, rtol=1e-5)
            torch.testing.assert_close(tr.ema, ema_ref, atol=1e-6, rtol=1e-5)
    # barriers per step: 2 (3 with the two-pass clip flow)
    per_step = 3 if clip > 0.0 else 2
    assert world.barriers[0] >= 3 * per_step


def test_ema_swap_and_checkpoint_see_the_whole_state(monkeypatch, host_only, tmp_path):
    from zero_b200.saver import Saver
    hp = _hp(clip_grad_norm=0.0, ema_decay=0.9)
    world, engines, trainers = _setup(2, monkeypatch, hp)
    total = engines[0].ps.total
    gen = torch.Generator().manual_seed(5)
    for r in range(2):
        engines[r].ps.grad.copy_(torch.randn(total, generator=gen) * 0.01)

    def one(r):
        trainers[r]._pending = True
        trainers[r].apply()
        trainers[r].ema_assign()                      # collective: syncs, then swaps the averages in
    _run_ranks(2, one)
    assert torch.equal(engines[0].ps.master, engines[1].ps.master)
    assert torch.equal(engines[0].ps.master, trainers[0].ema)
    assert torch.equal(engines[0].ps.mirror, engines[1].ps.mirror)
    _run_ranks(2, lambda r: trainers[r].ema_restore())
    assert torch.equal(engines[0].ps.master, engines[1].ps.master)
    Saver(output_dir=str(tmp_path)).save(engines[0], 1, trainer=trainers[0])
    arrays = Saver.state_of(engines[1], trainers[1])
    import numpy as np
    with np.load(Saver(output_dir=str(tmp_path)).latest()) as ck:
        for k in ck.files:
            np.testing.assert_array_equal(ck[k], arrays[k])


def test_safe_nan_keeps_the_all_reduce_path(monkeypatch, host_only):
    import zero_b200.engine as E
    from zero_b200.train import Trainer
    monkeypatch.setenv("ZB_SHARD_OPT", "1")
    hp = _hp(safe_nan=True)
    eng = E.Engine(hp, 40, 40, device="cpu")
    tr = Trainer(eng, hp, world_size=2, use_graph=False, side_stream=False)
    assert tr.shard is None
    monkeypatch.delenv("ZB_SHARD_OPT")
    tr = Trainer(eng, _hp(), world_size=2, use_graph=False, side_stream=False)
    assert tr.shard is None                            # "auto" needs an initialised process group to set it up
