"""world_size-2 check of the data-parallel host logic on CPU (gloo): per-rank gradients packed into the flat
gradient arena (ParamStore layout, fused k|v views included), ONE all-reduce(sum), 1/N scale == gradient of the
mean of the per-tower losses (main.py:42-43, utils/parallel.py:184-196; SURVEY.md section 4, identity 4)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import zero_oracle as zo
from zero_b200.engine import ModelConfig, ParamStore
from zero_b200.params import transformer_base


def _hp():
    return transformer_base(hidden_size=64, embed_size=64, filter_size=128, num_heads=2, num_encoder_layer=1,
                            num_decoder_layer=1)


def _batch(rank):
    g = torch.Generator().manual_seed(100 + rank)
    src = torch.randint(3, 96, (3, 7), generator=g)
    tgt = torch.randint(3, 96, (3, 6), generator=g)
    src[0, 4:] = 0
    tgt[1, 3:] = 0
    return src, tgt


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    hp = _hp()
    cfg = ModelConfig(hp, 96, 96)
    c = zo.Cfg(hp, 96, 96)
    ps = ParamStore(cfg, torch.device("cpu"))
    P = {k: v.requires_grad_(True) for k, v in zo.init_params(c, seed=5).items()}
    assert set(P) == set(ps.tf_names())
    src, tgt = _batch(rank)
    loss = zo.train_loss(c, P, src, tgt)[0]
    grads = torch.autograd.grad(loss, [P[k] for k in ps.tf_names()])
    for k, g in zip(ps.tf_names(), grads):
        ps.tf_view(ps.grad, k).copy_(g)
    dist.all_reduce(ps.grad, op=dist.ReduceOp.SUM)      # the single collective of the training step
    ps.grad.mul_(1.0 / world)                           # folded into zb_adam_tf's grad_scale on the GPU
    if rank == 0:
        torch.save({k: ps.tf_view(ps.grad, k).clone() for k in ps.tf_names()}, out_path)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(180)
def test_flat_arena_allreduce_equals_gradient_of_mean_loss(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    hp = _hp()
    c = zo.Cfg(hp, 96, 96)
    P = {k: v.requires_grad_(True) for k, v in zo.init_params(c, seed=5).items()}
    losses = [zo.train_loss(c, P, *_batch(r))[0] for r in range(2)]
    mean_loss = (losses[0] + losses[1]) / 2
    names = sorted(P)
    ref = torch.autograd.grad(mean_loss, [P[k] for k in names])
    for k, g in zip(names, ref):
        torch.testing.assert_close(got[k], g, atol=1e-6, rtol=1e-5, msg=k)


def test_param_store_layout_is_tma_legal_and_name_complete():
    hp = transformer_base()
    cfg = ModelConfig(hp, 32000, 32000)
    ps = ParamStore.__new__(ParamStore)
    from collections import OrderedDict
    ps.cfg, ps.slots, ps.alias, ps.tf_views = cfg, OrderedDict(), {}, OrderedDict()
    ps._plan()
    c = zo.Cfg(hp, 32000, 32000)
    shapes = zo.param_shapes(c)
    assert set(ps.tf_views) == set(shapes)
    assert len(shapes) == 207                      # 207 variables at config 2
    assert sum(int(torch.tensor(s).prod()) for s in shapes.values()) == 76907008
    for name, (off, shape) in ps.slots.items():
        assert off % 64 == 0, name                 # 128-byte aligned bf16 views


def test_batched_memory_projection_layout(monkeypatch):
    """ZB_BATCH_MEM_PROJ=1: the k_map | v_map weights of all decoder layers are windows of ONE [d, ndec * 2d] matrix.
    Same TF names, same shapes, every parameter element owned by exactly one TF variable, per-layer windows strided
    with a 16-byte-granular pitch (TMA-legal), and the TF-name order (hence the seeded initialisation) unchanged."""
    from collections import OrderedDict
    hp = transformer_base()

    def plan():
        cfg = ModelConfig(hp, 32000, 32000)
        ps = ParamStore.__new__(ParamStore)
        ps.cfg, ps.slots, ps.alias, ps.tf_views = cfg, OrderedDict(), {}, OrderedDict()
        ps._plan()
        return cfg, ps
    monkeypatch.setenv("ZB_BATCH_MEM_PROJ", "0")
    cfg0, ps0 = plan()
    monkeypatch.setenv("ZB_BATCH_MEM_PROJ", "1")
    cfg1, ps1 = plan()
    assert not cfg0.batch_mem and cfg1.batch_mem
    assert list(ps1.tf_views) == list(ps0.tf_views)
    d, nd = cfg1.d, cfg1.ndec
    arena = torch.zeros(ps1.total, dtype=torch.int32)
    for k in ps1.tf_views:
        v0, v1 = ps0.tf_view(torch.zeros(ps0.total), k), ps1.tf_view(arena, k)
        assert tuple(v0.shape) == tuple(v1.shape), k
        v1 += 1
    assert int(arena.sum()) == 76907008 and int(arena.max()) == 1      # disjoint windows, nothing counted twice
    for l in range(nd):
        w = ps1._view(arena, "dec%d.cross.kv.W" % l)
        assert tuple(w.shape) == (d, 2 * d) and w.stride() == (nd * 2 * d, 1)
        assert w.data_ptr() == ps1._view(arena, "dec.kvall.W").data_ptr() + l * 2 * d * arena.element_size()
        k_map = ps1.tf_view(arena, "transformer/decoder/layer_%d/cross_attention/dot_attention/k_map/W_0_0" % l)
        assert k_map.data_ptr() == w.data_ptr() and tuple(k_map.shape) == (d, d)
    assert ps1.slots["dec.kvall.W"][0] % 64 == 0 and (nd * 2 * d) % 8 == 0
    assert ps1.slots["dec.kvall.W"][0] >= ps1.dec_offset               # stays in the decoder-side all-reduce bucket


def _gather_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zero_b200 import evalu
    local = ([["r%d" % rank, "a"], ["b"]] if rank == 0 else [["c", "d", "e"]],
             [0.5, 0.25] if rank == 0 else [0.75], [0, 2] if rank == 0 else [1], 5 + rank, 0.1 * (rank + 1))
    got = evalu.gather_decoded(local, world)
    torch.save(got, os.path.join(out_dir, "g%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_rank_sharded_decoding_results_reach_every_rank(tmp_path):
    """evalu.decoding with world_size 2: rank r decodes every second batch, both ranks end up with all hypotheses
    (rank order; `indices` restore the corpus order) — so BLEU-driven decisions agree across ranks."""
    from zero_b200 import evalu
    mp.spawn(_gather_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    assert a == b
    trans, scores, indices, tokens, seconds = a
    assert trans == [["r0", "a"], ["b"], ["c", "d", "e"]] and scores == [0.5, 0.25, 0.75] and indices == [0, 2, 1]
    assert tokens == 11 and abs(seconds - 0.2) < 1e-12
    assert evalu.in_corpus_order(trans, indices) == [["r0", "a"], ["c", "d", "e"], ["b"]]
    assert evalu.gather_decoded((["x"], [1.0], [0], 1, 0.5), 1) == (["x"], [1.0], [0], 1, 0.5)
