"""Sharded optimizer step over NVLink / NVSwitch peer memory (opt-in: ZB_SHARD_OPT=1, world_size > 1).

The reference gathers every tower's gradients on one device, averages them variable by variable and runs Adam on
every variable (utils/parallel.py:134-208, main.py:42-43, :178-181).  The default data-parallel step here is the
direct restatement: NCCL all-reduce of the flat fp32 gradient arena, then the replicated Adam (zero_b200/train.py).
This module replaces both by ONE kernel per rank (`zb_shard_adam`, csrc/shard_opt.cu):

  * rank r owns the contiguous shard [lo_r, lo_r + n_r) of the arena (shards tile it, 64-element granular);
  * the kernel sums the shard's gradients over the ranks — in the NVSwitch through the multicast mapping of the
    symmetric gradient arena (multimem.ld_reduce) or from the peers' unicast mappings —, applies TF Adam to the
    shard's master / m / v, and stores the refreshed bf16 compute copy into every rank's mirror arena (multimem.st or
    one store per rank); the 1-D variables, which the forward pass reads in fp32, also go to every rank's master;
  * a cross-rank barrier before (all backwards finished) and after (all copies landed) is the only other
    communication of the step.  Per rank and step the Adam pass touches 1/N of the optimizer state (config 2 at
    N = 8: 2.3 GB -> 0.29 GB of HBM traffic) and the broadcast leg carries bf16; the gradient arena itself still
    crosses each GPU's link once on its way to the switch, as in any reduce-scatter.

fp32 master / Adam slots outside the own shard go stale; `sync_full_state()` (a collective: one broadcast per rank
and arena) makes them whole again before a checkpoint or an EMA swap.

`Transport` is the seam to the symmetric-memory plumbing (torch.distributed._symmetric_memory); the CPU tests drive
the same host logic through an in-process stand-in.
"""
from __future__ import annotations

import torch

from . import lib as L
from . import ops

f32, bf16 = torch.float32, torch.bfloat16
SLOT = 64   # ParamStore.ALIGN: every variable starts on a 64-element boundary


def plan_shards(total, world):
    """[(lo, n)] per rank: equal 64-element-granular shards tiling [0, total); trailing ranks may get less / none."""
    assert total % SLOT == 0 and world >= 1
    per = (total // SLOT + world - 1) // world * SLOT
    out = []
    for r in range(world):
        lo = min(r * per, total)
        out.append((lo, min(per, total - lo)))
    return out


def plan_region(lo, hi, world):
    """plan_shards over the arena range [lo, hi) (both 64-element granular)."""
    return [(lo + a, n) for a, n in plan_shards(hi - lo, world)]


def wide_slot_mask(ps):
    """uint8 per 64-element slot: 1 where the slot belongs to a 1-D variable (biases, LayerNorm scale / offset, the
    shared embedding bias, ReLA gates): engine code reads those through ParamStore.p(), i.e. from the fp32 master."""
    mask = torch.zeros(ps.total // SLOT, dtype=torch.uint8)
    for name, (off, shape) in ps.slots.items():
        if len(shape) == 1:
            size = (shape[0] + SLOT - 1) // SLOT
            mask[off // SLOT: off // SLOT + size] = 1
    return mask


class SymmMemTransport(object):
    """torch.distributed._symmetric_memory: allocation, peer / multicast addresses, stream-ordered barrier."""

    def __init__(self, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.symm = symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self._handles = []
        enable = getattr(symm, "enable_symm_mem_for_group", None)
        if enable is not None:          # a no-op where rendezvous() registers the group itself
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                enable(self.group.group_name)

    def alloc(self, n, dtype, device):
        t = self.symm.empty(n, dtype=dtype, device=device)
        t.zero_()
        return t

    def rendezvous(self, t):
        """-> (per-rank addresses of `t` as mapped in this process, multicast address or 0)."""
        h = self.symm.rendezvous(t, self.group)
        self._handles.append(h)
        ptrs = [h.get_buffer(r, tuple(t.shape), t.dtype, 0).data_ptr() for r in range(self.world)]
        if ptrs[self.rank] != t.data_ptr():
            raise L.ZeroB200Error("symmetric memory: the local mapping of the arena is not the arena itself")
        mc = int(getattr(h, "multicast_ptr", 0) or 0)
        if mc:
            # multicast_ptr follows the convention of buffer_ptrs (allocation base): same offset as the local tensor
            mc += t.data_ptr() - int(h.buffer_ptrs[self.rank])
        return ptrs, mc

    def barrier(self, channel=0):
        """Enqueued on the current stream: returns (on the device) once every rank has reached it.  Barriers that may
        be in flight on different streams at the same time use different channels."""
        self._handles[0].barrier(channel=channel)

    def broadcast(self, t, src):
        import torch.distributed as dist
        dist.broadcast(t, src=dist.get_global_rank(self.group, src), group=self.group)


class ShardedStep(object):
    """Owns the symmetric arenas of one rank and issues the fused step.  `trainer` supplies the hyper-parameters and
    the local scalars (norms, clip_scale)."""

    def __init__(self, engine, transport, use_multicast=True, split=None):
        """`split` (an arena offset, normally ParamStore.dec_offset): the arena is sharded as TWO regions,
        [split, total) — whose gradients are final after the decoder backward — and [0, split); `step_early` reduces
        and updates the first one on a second stream while the encoder backward still runs."""
        ps = engine.ps
        if ps.adam_m is None or ps.adam_v is None:
            raise L.ZeroB200Error("ShardedStep needs the Adam slots (create it from a Trainer)")
        self.ps, self.tp = ps, transport
        self.world, self.rank = transport.world, transport.rank
        if self.world > L.SHARD_MAX_WORLD:
            raise L.ZeroB200Error("sharded optimizer step: at most %d ranks" % L.SHARD_MAX_WORLD)
        dev = ps.device
        self.split = int(split) if split else None
        regions = [(0, ps.total)] if not self.split else [(self.split, ps.total), (0, self.split)]
        self.plans = [plan_region(a, b, self.world) for a, b in regions]
        self.ranges = [plan[self.rank] for plan in self.plans]          # this rank's (lo, n) per region
        self.shards = self.plans[0]                                      # single-region view (clip flow, tests)
        self.lo, self.n = self.ranges[0]
        self._early_done = False
        # the three arenas peers touch move into symmetric memory; views are taken from ps.* on every use
        # (ParamStore._view), so rebinding before the first captured step is enough
        for name, dtype in (("grad", f32), ("mirror", bf16), ("master", f32)):
            old = getattr(ps, name)
            new = transport.alloc(ps.total, dtype, dev)
            new.copy_(old)
            setattr(ps, name, new)
        self.parts = transport.alloc(2 * self.world, f32, dev)        # [world][2]: every rank's {sum g^2, sum p^2}
        self.grad_ptrs, self.grad_mc = transport.rendezvous(ps.grad)
        self.mirror_ptrs, self.mirror_mc = transport.rendezvous(ps.mirror)
        self.param_ptrs, self.param_mc = transport.rendezvous(ps.master)
        self.parts_ptrs, _ = transport.rendezvous(self.parts)
        if not use_multicast:
            self.grad_mc = self.mirror_mc = self.param_mc = 0
        self.wide = wide_slot_mask(ps).to(dev)
        self.done = torch.zeros(1, dtype=torch.int32, device=dev)
        self.reduced = None      # fp32 [n]: the shard's reduced gradients between the two passes of the clip flow
        self.steps = 0

    # ---------------------------------------------------------------------------------------------- the step
    def _launch(self, tr, lr_t, gscale, flags, clip_scale=None, local_grad=None, region=0):
        ps = self.ps
        lo, n = self.ranges[region]
        if local_grad is None:
            gptrs, gmc, sources = self.grad_ptrs, self.grad_mc, self.world
        else:
            # second pass of the clip flow: the summed gradients already sit in `local_grad` (indexed like the arena)
            gptrs, gmc, sources = [local_grad], 0, 1
        ops.shard_adam(lo, n, self.world, self.rank, gptrs, self.mirror_ptrs, ps.master, ps.adam_m, ps.adam_v,
                       tr.beta1, tr.beta2, tr.eps, lr_t, gscale, flags=flags, grad_mc=gmc, mirror_mc=self.mirror_mc,
                       grad_out=None if not (flags & L.ZB_SHARD_STORE_GRAD) else self._reduced_base(),
                       clip_scale=clip_scale, norms=tr.norms, norm_parts_ptrs=self.parts_ptrs, done_counter=self.done,
                       wide_mask=self.wide, param_ptrs=self.param_ptrs, param_mc=self.param_mc, grad_sources=sources)

    def _reduced_base(self):
        """Address of a buffer that the kernel may index with ARENA offsets: element `lo` is reduced[0]."""
        if self.reduced is None:
            self.reduced = torch.empty(max(self.n, 8), dtype=f32, device=self.ps.device)
        return self.reduced.data_ptr() - 4 * self.lo

    def step(self, tr, lr_t, gscale):
        """All ranks call this after their backward.  Without clipping: barrier, one kernel, barrier.  With
        clip_by_global_norm (utils/cycle.py:94-101) the factor needs the norm of the SUMMED gradients first:
        pass 1 reduces the shard into a local buffer and exchanges sum g^2, pass 2 updates from that buffer."""
        tp = self.tp
        full = L.ZB_SHARD_UPDATE | L.ZB_SHARD_NORM_G | L.ZB_SHARD_NORM_P
        if self.split:
            if tr.clip is not None:
                raise L.ZeroB200Error("the two-region sharded step has no clip flow (construct without split)")
            if not self._early_done:            # nobody ran step_early: both regions now
                tr.norms.zero_()
                tp.barrier()
                self._launch(tr, lr_t, gscale, full, region=0)
            else:
                tp.barrier()                    # every rank's encoder backward has finished
            self._launch(tr, lr_t, gscale, full, region=1)
            tp.barrier()
            self._early_done = False
            self.steps += 1
            return
        tr.norms.zero_()
        tp.barrier()
        if tr.clip is None:
            self._launch(tr, lr_t, gscale, full)
        else:
            self._launch(tr, lr_t, gscale, L.ZB_SHARD_STORE_GRAD | L.ZB_SHARD_NORM_G)
            tp.barrier()
            gn = torch.sqrt(self.parts.view(self.world, 2)[:, 0].sum())        # already scaled by gscale
            torch.div(tr.clip, torch.clamp(gn, min=tr.clip), out=tr.clip_scale[0])
            self._launch(tr, lr_t, gscale, L.ZB_SHARD_UPDATE | L.ZB_SHARD_NORM_P, clip_scale=tr.clip_scale,
                         local_grad=self._reduced_base())
        tp.barrier()
        self.steps += 1

    def step_early(self, tr, lr_t, gscale):
        """First region of a split plan (the decoder side).  The caller has made the current stream wait for this rank's
        decoder backward; the barrier makes it wait for every other rank's too — the kernel then reads their decoder-
        side gradients and overwrites their decoder-side weights while their encoder backward runs."""
        assert self.split and not self._early_done
        tr.norms.zero_()
        self.tp.barrier(1) if isinstance(self.tp, SymmMemTransport) else self.tp.barrier()
        self._launch(tr, lr_t, gscale, L.ZB_SHARD_UPDATE | L.ZB_SHARD_NORM_G | L.ZB_SHARD_NORM_P, region=0)
        self._early_done = True

    def norms(self):
        """fp32 [2] device tensor: {sum (g * gscale)^2, sum p^2} over the whole arena (all ranks' shards)."""
        return self.parts.view(self.world, 2).sum(0)

    # ---------------------------------------------------------------------------------------------- whole state
    def sync_full_state(self, extra=()):
        """Collective.  Every rank receives every shard of the fp32 master and the Adam slots (and of the arenas in
        `extra`, e.g. the EMA shadow) from its owner; afterwards the local copies are complete and identical."""
        ps = self.ps
        for arena in (ps.master, ps.adam_m, ps.adam_v) + tuple(extra):
            for plan in self.plans:
                for r, (lo, n) in enumerate(plan):
                    if n:
                        self.tp.broadcast(arena[lo:lo + n], r)
