"""zb_shard_adam (csrc/shard_opt.cu) on the GPU; the product reaches it under ZB_SHARD_OPT=1.

1. One device, N emulated ranks: peer pointers are just device addresses, so N gradient / mirror / master arenas on the
   same GPU exercise the whole unicast path of the kernel (reduce in rank order, Adam, bf16 + fp32 fan-out, norm
   exchange, the two-pass clip flow) against the oracle's replicated Adam on the mean gradient.
2. Two devices (skipped on a one-GPU box), two processes, NCCL + symmetric memory: three training steps of the real
   engine with the fused step (multicast and unicast) against the all-reduce + replicated Adam default."""
import math
import os
import socket

import pytest
import torch

pytestmark = [pytest.mark.gpu]

f32, bf16 = torch.float32, torch.bfloat16


@pytest.mark.parametrize("world,clip", [(1, None), (2, None), (3, None), (8, None), (3, 0.02)])
def test_shard_adam_kernel_emulated_ranks(world, clip):
    from oracle import zero_oracle as zo
    from zero_b200 import lib as L
    from zero_b200 import ops
    from zero_b200.shard_opt import plan_shards
    dev = torch.device("cuda")
    total = 64 * 1237
    gen = torch.Generator().manual_seed(11)
    p0 = torch.randn(total, generator=gen)
    wide = (torch.rand(total // 64, generator=gen) < 0.2).to(torch.uint8)
    wide_e = wide.bool().repeat_interleave(64)
    master = [p0.clone().to(dev) for _ in range(world)]
    m = [torch.zeros(total, device=dev) for _ in range(world)]
    v = [torch.zeros(total, device=dev) for _ in range(world)]
    mirror = [torch.zeros(total, dtype=bf16, device=dev) for _ in range(world)]
    grad = [torch.empty(total, device=dev) for _ in range(world)]
    parts = [torch.zeros(world, 2, device=dev) for _ in range(world)]
    norms = [torch.zeros(2, device=dev) for _ in range(world)]
    done = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(world)]
    shards = plan_shards(total, world)
    wide_d = wide.to(dev)
    p_ref, m_ref, v_ref = p0.clone(), torch.zeros(total), torch.zeros(total)
    gscale = 1.0 / world
    for step in (1, 2, 3):
        gs = [torch.randn(total, generator=gen) * 0.01 for _ in range(world)]
        for r in range(world):
            grad[r].copy_(gs[r])
            norms[r].zero_()
        lr_t = 0.01 * math.sqrt(1 - 0.98 ** step) / (1 - 0.9 ** step)
        common = dict(mirror_ptrs=[t.data_ptr() for t in mirror], norm_parts_ptrs=[t.data_ptr() for t in parts],
                      wide_mask=wide_d, param_ptrs=[t.data_ptr() for t in master])
        g_mean = sum(gs) / world
        gn = float(g_mean.norm())
        if clip is None:
            for r, (lo, n) in enumerate(shards):
                ops.shard_adam(lo, n, world, r, [t.data_ptr() for t in grad], param=master[r], m=m[r], v=v[r],
                               beta1=0.9, beta2=0.98, eps=1e-8, lr_t=lr_t, grad_scale=gscale, norms=norms[r],
                               done_counter=done[r], **common)
            g_used = g_mean
        else:
            reduced = [torch.zeros(max(n, 8), device=dev) for _, n in shards]
            for r, (lo, n) in enumerate(shards):
                ops.shard_adam(lo, n, world, r, [t.data_ptr() for t in grad], param=master[r], m=m[r], v=v[r],
                               beta1=0.9, beta2=0.98, eps=1e-8, lr_t=lr_t, grad_scale=gscale, norms=norms[r],
                               done_counter=done[r], flags=L.ZB_SHARD_STORE_GRAD | L.ZB_SHARD_NORM_G,
                               grad_out=reduced[r].data_ptr() - 4 * lo, **common)
            torch.cuda.synchronize()
            got_gn = float(torch.sqrt(parts[0][:, 0].sum()))
            assert abs(got_gn - gn) < 1e-4 * gn
            factor = clip / max(gn, clip)
            cs = torch.tensor([factor], device=dev)
            for r, (lo, n) in enumerate(shards):
                torch.testing.assert_close(reduced[r][:n].cpu(), sum(gs)[lo:lo + n], atol=1e-7, rtol=1e-5)
                ops.shard_adam(lo, n, world, r, [reduced[r].data_ptr() - 4 * lo], param=master[r], m=m[r], v=v[r],
                               beta1=0.9, beta2=0.98, eps=1e-8, lr_t=lr_t, grad_scale=gscale, norms=norms[r],
                               done_counter=done[r], flags=L.ZB_SHARD_UPDATE | L.ZB_SHARD_NORM_P, clip_scale=cs,
                               grad_sources=1, **common)
            g_used = g_mean * factor
        torch.cuda.synchronize()
        p_before = p_ref.clone()
        p_ref, m_ref, v_ref = zo.adam_tf_step(p_ref, m_ref, v_ref, g_used, step, 0.01, 0.9, 0.98, 1e-8)
        owned = torch.cat([master[r][lo:lo + n] for r, (lo, n) in enumerate(shards)]).cpu()
        torch.testing.assert_close(owned, p_ref, atol=1e-6, rtol=1e-5)
        for r, (lo, n) in enumerate(shards):
            assert torch.equal(mirror[r].cpu(), owned.to(bf16)), "rank %d compute copy" % r
            torch.testing.assert_close(m[r][lo:lo + n].cpu(), m_ref[lo:lo + n], atol=1e-8, rtol=1e-5)
            torch.testing.assert_close(v[r][lo:lo + n].cpu(), v_ref[lo:lo + n], atol=1e-10, rtol=1e-5)
            torch.testing.assert_close(master[r].cpu()[wide_e], p_ref[wide_e], atol=1e-6, rtol=1e-5)
            tot = parts[r].sum(0).cpu()
            assert abs(float(tot[0].sqrt()) - gn) < 1e-4 * gn
            assert abs(float(tot[1].sqrt()) - float(p_before.norm())) < 1e-3
            assert int(done[r]) == 0
            # state outside the shard and outside the fp32-read slots is not touched
            foreign = torch.ones(total, dtype=torch.bool)
            foreign[lo:lo + n] = False
            assert float(m[r].cpu()[foreign].abs().max() if foreign.any() else 0.0) == 0.0


def test_shard_adam_rejects_bad_arguments():
    from zero_b200 import lib as L
    from zero_b200 import ops
    t = torch.zeros(128, device="cuda")
    mir = torch.zeros(128, dtype=bf16, device="cuda")
    with pytest.raises(L.ZeroB200Error):      # lo not a multiple of 8
        ops.shard_adam(4, 64, 1, 0, [t.data_ptr()], [mir.data_ptr()], t, t, t, 0.9, 0.98, 1e-8, 0.1, 1.0, flags=1)
    with pytest.raises(L.ZeroB200Error):      # rank outside the world
        ops.shard_adam(0, 64, 2, 2, [t.data_ptr()] * 2, [mir.data_ptr()] * 2, t, t, t, 0.9, 0.98, 1e-8, 0.1, 1.0, flags=1)
    with pytest.raises(L.ZeroB200Error):      # norm flags without the accumulators
        ops.shard_adam(0, 64, 1, 0, [t.data_ptr()], [mir.data_ptr()], t, t, t, 0.9, 0.98, 1e-8, 0.1, 1.0, flags=1 | 4)


# ------------------------------------------------------------------------------------------------ two devices
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    # "allreduce": bucketed NCCL all-reduce + replicated Adam; "1" / "p2p": the fused step over the multicast / unicast
    # mappings in two regions (decoder side under the encoder backward); "p2p-flat": one region after the backward
    os.environ["ZB_SHARD_OPT"] = "0" if mode == "allreduce" else mode.split("-")[0]
    os.environ["ZB_SHARD_OVERLAP"] = "0" if mode.endswith("-flat") else "1"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from tests.golden_util import load_golden
    from zero_b200.engine import Engine
    from zero_b200.train import Trainer
    z, hp, variables, grads, vs, vt = load_golden("transformer")
    hp.override_from_dict(dict(lrate=1.0, warmup_steps=10, beta1=0.9, beta2=0.98, epsilon=1e-8, clip_grad_norm=0.0,
                               lrate_strategy="noam"))
    eng = Engine(hp, vs, vt)
    eng.ps.load_state_dict(variables)
    tr = Trainer(eng, hp, world_size=world, use_graph=False)
    assert (tr.shard is not None) == (mode != "allreduce")
    if tr.shard is not None:
        assert (tr.shard.split is not None) == (not mode.endswith("-flat"))
    src, tgt = torch.from_numpy(z["source"]), torch.from_numpy(z["target"])
    if rank == 1:                                   # a different batch per tower (main.py:268-273)
        src, tgt = torch.flip(src, [0]), torch.flip(tgt, [0])
    losses = []
    for _ in range(3):
        losses.append(float(tr.step(src, tgt)[0]))
    tr.sync_full_state()
    torch.cuda.synchronize()
    torch.save({"losses": losses, "master": eng.ps.master.cpu(), "mirror": eng.ps.mirror.cpu(),
                "m": eng.ps.adam_m.cpu(), "gnorm": tr.gradient_norm(), "pnorm": tr.parameter_norm(),
                "mc": None if tr.shard is None else bool(tr.shard.grad_mc)},
               os.path.join(out_dir, "%s_%d.pt" % (mode, rank)))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one NVLink box")
def test_two_rank_training_fused_step_equals_allreduce_path(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path)
    for mode in ("allreduce", "1", "p2p", "p2p-flat"):
        mp.spawn(_worker, args=(2, _free_port(), mode, out), nprocs=2, join=True)
    ref = [torch.load(os.path.join(out, "allreduce_%d.pt" % r)) for r in range(2)]
    assert torch.equal(ref[0]["master"], ref[1]["master"])
    for mode in ("1", "p2p", "p2p-flat"):
        got = [torch.load(os.path.join(out, "%s_%d.pt" % (mode, r))) for r in range(2)]
        assert torch.equal(got[0]["master"], got[1]["master"]) and torch.equal(got[0]["mirror"], got[1]["mirror"])
        assert torch.equal(got[0]["mirror"], got[0]["master"].to(bf16))
        for r in range(2):
            for a, b in zip(got[r]["losses"], ref[r]["losses"]):
                assert abs(a - b) < 2e-2, (mode, r, got[r]["losses"], ref[r]["losses"])
            assert abs(got[r]["gnorm"] - ref[r]["gnorm"]) < 2e-2 * ref[r]["gnorm"]
            assert abs(got[r]["pnorm"] - ref[r]["pnorm"]) < 1e-3 * ref[r]["pnorm"]
        # same arithmetic up to the summation order of two addends (exact) and bf16 forward noise between runs
        d = (got[0]["master"] - ref[0]["master"]).abs().max()
        assert float(d) < 5e-2, float(d)
        assert (mode.startswith("p2p") and got[0]["mc"] is False) or mode == "1"
