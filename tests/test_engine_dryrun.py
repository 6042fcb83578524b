"""Dry run of the engine's host-side schedule on the CPU: every kernel entry point of zero_b200.ops is replaced by a
recorder, buffers live on the CPU, and forward_backward / score / one cached decode step are driven through the
real Python control flow.  Nothing is computed — what is checked is that the schedule itself is sound for every
model family and every host-level switch: names resolve, views of the planned buffers have the shapes and strides
the argument builders (ops.attention_args, ops.gemm_args: the real ones) accept, and the launch sequence is the
expected one.  The numerics of the same paths are the GPU tests' job (tests/test_model_gpu.py)."""
import collections

import pytest
import torch

from zero_b200.params import transformer_base

BUILDERS = ("attention_args", "gemm_args", "wgrad_args", "beam_args")
PURE = ("vocab_topk_supported",)          # shape predicates that launch nothing: kept real


def _mock_ops(monkeypatch):
    """Replace every launching function of zero_b200.ops by a recorder; returns the list of recorded names."""
    import zero_b200.ops as ops
    calls = []
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    for name in dir(ops):
        fn = getattr(ops, name)
        if callable(fn) and not isinstance(fn, type) and not name.startswith("_") \
                and getattr(fn, "__module__", "") == ops.__name__ and name not in BUILDERS and name not in PURE:
            monkeypatch.setattr(ops, name, (lambda n: (lambda *a, **k: calls.append(n)))(name))
    return calls


def _engine(**over):
    import zero_b200.engine as E
    hp = transformer_base(hidden_size=128, embed_size=128, filter_size=256, num_heads=2, num_encoder_layer=2,
                          num_decoder_layer=3, **over)
    eng = E.Engine(hp, 208, 208, device="cpu")
    eng.ps.init_random(3)
    return eng


def _dry_engine(monkeypatch, **over):
    calls = _mock_ops(monkeypatch)
    return _engine(**over), calls


def _batch():
    g = torch.Generator().manual_seed(0)
    src = torch.randint(3, 208, (4, 9), generator=g)
    tgt = torch.randint(3, 208, (4, 7), generator=g)
    src[1, 5:] = 0
    tgt[2, 4:] = 0
    return src, tgt


FAMILIES = [dict(model_name="transformer", scope_name="transformer"),
            dict(model_name="transformer_rpr", scope_name="transformer_rpr", max_relative_position=4),
            dict(model_name="transformer_rela", scope_name="transformer_rela"),
            dict(model_name="transformer_aan", scope_name="transformer_aan"),
            dict(model_name="transformer_aan", scope_name="transformer_aan", use_ffn=True, aan_mask=False),
            dict(model_name="transformer_fuse", scope_name="transformer_fuse")]


@pytest.mark.parametrize("family", FAMILIES, ids=lambda f: f["model_name"] + ("+ffn" if f.get("use_ffn") else ""))
@pytest.mark.parametrize("switch", [None, "ZB_BATCH_MEM_PROJ", "ZB_DECODE_FUSED_SMALL"])
def test_schedule_runs_for_every_family_and_switch(family, switch, monkeypatch):
    # both switches are on by default since round 2; "None" is the schedule with both turned off
    for name in ("ZB_BATCH_MEM_PROJ", "ZB_DECODE_FUSED_SMALL"):
        monkeypatch.setenv(name, "1" if name == switch else "0")
    eng, calls = _dry_engine(monkeypatch, **family)
    src, tgt = _batch()
    loss = eng.forward_backward(src, tgt)
    assert tuple(loss.shape) == (1,)
    n = collections.Counter(calls)
    assert n["softmax_ce"] == 1 and n["embed_fwd"] == 2 and n["embed_bwd"] == 2
    assert n["attention_fwd"] == n["attention_bwd"] > 0
    score = eng.score(src, tgt)
    assert tuple(score.shape) == (4,)
    # one cached decode step per family (beam 2): the per-sentence memories, the per-beam caches / running sums
    del calls[:]
    state = eng.encoding_fn(src)
    state.begin_search(2, cap=12)
    tok = torch.zeros(8, 1, dtype=torch.int32)
    for t in range(2):
        logits, state = eng.decoding_fn(tok, state, t)
        state.reorder(torch.arange(8, dtype=torch.int32), t)
    assert tuple(logits.shape) == (8, 208)
    n = collections.Counter(calls)
    if switch == "ZB_DECODE_FUSED_SMALL" and family["model_name"] == "transformer_aan" and not family.get("use_ffn"):
        assert n["aan_cat_step"] == 2 * 3 and n["aan_gate_ln"] == 2 * 3 and n["aan_step"] == 0
    elif family["model_name"] == "transformer_aan":
        assert n["aan_step"] == 2 * 3 and n["aan_gate_fwd"] == 2 * 3
    if switch == "ZB_BATCH_MEM_PROJ":
        batched = family["model_name"] not in ("transformer_aan", "transformer_fuse")
        assert eng.cfg.batch_mem == batched and ("dec.kvall.W" in eng.ps.slots) == batched


def test_batched_memory_projection_changes_the_launch_count(monkeypatch):
    """One memory-projection GEMM forward and one dgrad backward for all decoder layers instead of one per layer."""
    src, tgt = _batch()
    calls = _mock_ops(monkeypatch)
    monkeypatch.setenv("ZB_BATCH_MEM_PROJ", "0")
    eng0 = _engine()
    del calls[:]
    eng0.forward_backward(src, tgt)
    n0 = collections.Counter(calls)
    monkeypatch.setenv("ZB_BATCH_MEM_PROJ", "1")
    eng1 = _engine()
    del calls[:]
    eng1.forward_backward(src, tgt)
    n1 = collections.Counter(calls)
    ndec = eng1.cfg.ndec
    assert n0["linear_fwd"] - n1["linear_fwd"] == ndec - 1
    assert n0["linear_dgrad"] - n1["linear_dgrad"] == ndec - 1
    assert n0["attention_fwd"] == n1["attention_fwd"]


def test_empty_tower_contributes_zero_loss_and_no_launches(monkeypatch):
    """The reference's zero-shape guard (models/transformer.py:213-216): a tower that receives no sentence returns
    loss 0; here it also launches nothing and leaves the (zeroed) gradient arena alone."""
    eng, calls = _dry_engine(monkeypatch, model_name="transformer", scope_name="transformer")
    eng.ps.grad.fill_(1.0)
    del calls[:]
    loss = eng.forward_backward(torch.zeros(0, 5, dtype=torch.int64), torch.zeros(0, 4, dtype=torch.int64))
    assert tuple(loss.shape) == (1,) and float(loss) == 0.0
    assert calls == [] and float(eng.ps.grad.abs().sum()) == 0.0
    assert tuple(eng.score(torch.zeros(0, 5, dtype=torch.int64), torch.zeros(0, 4, dtype=torch.int64)).shape) == (0,)
    # the next real batch runs the normal schedule again
    src, tgt = _batch()
    eng.forward_backward(src, tgt)
    assert collections.Counter(calls)["softmax_ce"] == 1


@pytest.mark.parametrize("family", [FAMILIES[0], FAMILIES[3]], ids=["transformer", "transformer_aan"])
def test_vocabulary_sizes_need_not_be_multiples_of_8(family, monkeypatch):
    """Real vocabularies are 3 specials + N words.  Logit rows get a pitch rounded up to 8 elements (TMA), the kernels
    see n = V columns (real argument builders: strided views accepted), decoding_fn still returns dense [rows, V]."""
    import zero_b200.engine as E
    import zero_b200.ops as ops
    calls = _mock_ops(monkeypatch)
    seen = []
    real = ops.gemm_args
    monkeypatch.setattr(ops, "gemm", lambda *a, **k: seen.append(real(*a, **k)))
    hp = transformer_base(hidden_size=128, embed_size=128, filter_size=256, num_heads=2, num_encoder_layer=1,
                          num_decoder_layer=1, **family)
    eng = E.Engine(hp, 203, 205, device="cpu")
    eng.ps.init_random(3)
    assert eng.cfg.vt == 205 and eng.cfg.vt_pitch == 208
    src, tgt = _batch()
    src, tgt = src.clamp(max=202), tgt.clamp(max=204)
    eng.forward_backward(src, tgt)
    vocab_gemms = [a for a in seen if 205 in (a.m, a.n, a.k)]
    assert len(vocab_gemms) == 3                                   # logits, dE (m = V), dfeat (k = V)
    fwd = next(a for a in vocab_gemms if a.n == 205)
    assert fwd.ldd == 208 and fwd.k == 128
    assert any(a.k == 205 and a.lda == 208 for a in vocab_gemms)   # d_logits read with its padded pitch
    assert any(a.m == 205 and a.lda == 208 for a in vocab_gemms)
    state = eng.encoding_fn(src)
    state.begin_search(2, cap=12)
    logits, state = eng.decoding_fn(torch.zeros(8, 1, dtype=torch.int32), state, 0)
    assert tuple(logits.shape) == (8, 205) and logits.is_contiguous()
    assert tuple(eng.score(src, tgt).shape) == (4,)


def test_workspace_footprint_is_the_largest_batch_not_the_sum_of_shapes(monkeypatch):
    """Token-budget batching gives a new (B, S, T) almost every step (data.py:67-117): the planned buffers must be
    reused across shapes.  Static addresses per name for a repeated shape (CUDA graphs), growth only for a larger
    batch, and a generation counter that tells graph holders when an allocation they point into was replaced."""
    import zero_b200.engine as E
    ws = E.Workspace(torch.device("cpu"))
    a = ws.get("x", (4, 8))
    assert tuple(a.shape) == (4, 8) and a.is_contiguous() and a.dtype == torch.bfloat16
    assert ws.get("x", (4, 8)).data_ptr() == a.data_ptr() and ws.get("x", (2, 16)).data_ptr() == a.data_ptr()
    assert ws.get("x", (3, 5)).data_ptr() == a.data_ptr()              # smaller: a view of the same storage
    assert ws.get("x", (4, 8), torch.float32).data_ptr() != a.data_ptr()   # one pool per (name, dtype)
    before, gen = ws.nbytes(), ws.generation
    assert gen == 0                                                    # first allocations do not count
    b = ws.get("x", (8, 8))                                            # larger: replaced, graph holders are told
    assert b.data_ptr() != a.data_ptr() and ws.nbytes() == before + (64 - 32) * 2 and ws.generation == gen + 1
    c = ws.get("x", (9, 8))                                            # geometric growth: room for 80 elements
    assert ws.generation == gen + 2 and ws.get("x", (10, 8)).data_ptr() == c.data_ptr() and ws.generation == gen + 2
    # a real schedule over many batch shapes: the footprint stops growing once the largest has been seen
    eng, calls = _dry_engine(monkeypatch, model_name="transformer", scope_name="transformer")
    g = torch.Generator().manual_seed(1)

    def batch(b_, s_, t_):
        return torch.randint(3, 208, (b_, s_), generator=g), torch.randint(3, 208, (b_, t_), generator=g)
    eng.forward_backward(*batch(6, 12, 11))
    peak = eng.ws.nbytes()
    for shape in ((2, 5, 7), (5, 12, 3), (6, 11, 11), (3, 9, 10), (4, 12, 11), (1, 1, 1)):
        eng.forward_backward(*batch(*shape))
    assert eng.ws.nbytes() <= peak * 1.3, (eng.ws.nbytes(), peak)


@pytest.mark.parametrize("family", FAMILIES, ids=lambda f: f["model_name"] + ("+ffn" if f.get("use_ffn") else ""))
def test_fixed_shapes_never_replace_a_workspace_buffer(family, monkeypatch):
    """Captured CUDA graphs are dropped when the workspace generation moves, so a fixed batch shape must never move it:
    no buffer name is requested with a growing shape along the decode steps, across searches, or across training /
    scoring steps of the same shape."""
    eng, calls = _dry_engine(monkeypatch, **family)
    src, tgt = _batch()
    for _ in range(2):
        state = eng.encoding_fn(src)
        state.begin_search(2, cap=12)
        tok = torch.zeros(8, 1, dtype=torch.int32)
        for t in range(8):
            _, state = eng.decoding_fn(tok, state, t)
            state.reorder(torch.arange(8, dtype=torch.int32), t)
    for _ in range(2):
        eng.forward_backward(src, tgt)
        eng.score(src, tgt)
    assert eng.ws.generation == 0
    # a larger batch does move it (and only then)
    big = torch.cat([src, src]), torch.cat([tgt, tgt])
    eng.forward_backward(*big)
    assert eng.ws.generation > 0
    g = eng.ws.generation
    eng.forward_backward(src, tgt)
    eng.forward_backward(*big)
    assert eng.ws.generation == g


@pytest.mark.parametrize("case", ["default", "ZB_BEAM_FUSED=0", "noise", "wrapped", "beam5", "dev", "aan"])
def test_search_hands_candidates_to_the_beam_step_only_on_its_own_plain_path(case, monkeypatch):
    """search.beam_search asks the engine's decode step for beam candidates instead of logits (K8 fused:
    ops.vocab_topk -> zb_beam_step(cand)) exactly when nothing else needs the logits: the engine's own cached
    decoding_fn, no Gumbel noise, beam <= 4.  Every other case keeps the logits path (zb_gemm -> zb_beam_step(logits))."""
    import zero_b200.ops as ops
    from zero_b200 import search
    from zero_b200.params import SimpleVocab
    family = dict(model_name="transformer_aan", scope_name="transformer_aan") if case == "aan" else \
        dict(model_name="transformer", scope_name="transformer")
    eng, calls = _dry_engine(monkeypatch, **family)
    hp = transformer_base(hidden_size=128, embed_size=128, filter_size=256, num_heads=2, num_encoder_layer=2,
                          num_decoder_layer=3, beam_size=5 if case == "beam5" else 4, decode_length=3, **family)
    hp.add_hparam("tgt_vocab", SimpleVocab(208))
    hp.add_hparam("decode_graph", False)
    if case == "noise":
        hp.enable_noise_beam_search = True
    if case == "dev":
        hp.search_mode = "dev"
    monkeypatch.setenv("ZB_BEAM_FUSED", "0" if case == "ZB_BEAM_FUSED=0" else "1")
    steps = []

    def fake_topk(feat, table, workspace, skip_col=-1, temperature=1.0):
        calls.append("vocab_topk")
        return ops.BeamCandidates(torch.zeros(4), feat.shape[0], table.shape[0], 2, skip_col, temperature)

    def fake_beam_step(a):
        steps.append((bool(a.cand), bool(a.logits), int(a.time)))

    monkeypatch.setattr(ops, "vocab_topk", fake_topk)
    monkeypatch.setattr(ops, "beam_step", fake_beam_step)
    # the loop condition lives on the device: here three steps, then stop
    monkeypatch.setattr(search.BeamState, "cond_async", lambda self, t: t if t < 3 else None)
    monkeypatch.setattr(search.BeamState, "cond_wait", lambda self, slot: True)
    src, _ = _batch()
    decoding_fn = eng.decoding_fn_dev if case == "dev" else eng.decoding_fn
    if case == "wrapped":
        inner = decoding_fn
        decoding_fn = lambda tok, state, t: inner(tok, state, t)     # noqa: E731  (a user's hook around the step)
    out = search.beam_search({"source": src}, eng.encoding_fn, decoding_fn, hp)
    assert tuple(out["seq"].shape[:2]) == (4, int(hp.beam_size))
    fused = case in ("default", "aan")
    assert [s[2] for s in steps] == [0, 1, 2]
    assert all(s[0] == fused and s[1] == (not fused) for s in steps), steps
    n = collections.Counter(calls)
    assert n["vocab_topk"] == (3 if fused else 0)
    assert n["gumbel_add"] == (3 if case == "noise" else 0)


@pytest.mark.parametrize("name", ["transformer_shared_emb", "transformer_softmax_emb", "transformer_aan_shared_emb"])
def test_embedding_sharing_switches_plan_the_references_variables(name, monkeypatch):
    """shared_source_target_embedding / shared_target_softmax_embedding away from their defaults
    (models/transformer.py:21-22, 96-97, 186-189): the parameter store holds exactly the variables the reference
    created (by TF name and shape, from the reference-executed golden), loads them, and the schedule — forward,
    backward, score, one cached decode step — runs with the embedding gradients routed to the right tables."""
    import zero_b200.engine as E
    import zero_b200.ops as ops
    from tests.golden_util import load_golden
    calls = _mock_ops(monkeypatch)
    z, hp, variables, grads, vs, vt = load_golden(name)
    eng = E.Engine(hp, vs, vt, device="cpu")
    assert set(eng.ps.tf_views) == set(variables), sorted(set(eng.ps.tf_views) ^ set(variables))
    for k, v in variables.items():
        assert tuple(eng.ps.tf_view(eng.ps.master, k).shape) == tuple(v.shape), k
    eng.ps.load_state_dict(variables)
    for k, v in variables.items():
        assert torch.equal(eng.ps.tf_view(eng.ps.master, k), v), k
    shared = bool(hp.shared_source_target_embedding)
    assert eng._tgt_table() == ("src_emb" if shared else "tgt_emb")
    assert eng._softmax_table() == ("src_emb" if shared else ("tgt_emb" if hp.shared_target_softmax_embedding
                                                             else "softmax_emb"))
    tables = []
    real_p = ops._p

    def embed_bwd(ids, dy, d_table, *a, **k):
        calls.append("embed_bwd")
        tables.append(d_table.data_ptr())
    monkeypatch.setattr(ops, "embed_bwd", embed_bwd)
    src, tgt = torch.from_numpy(z["source"]), torch.from_numpy(z["target"])
    loss = eng.forward_backward(src, tgt)
    assert tuple(loss.shape) == (1,) and real_p is ops._p
    # target embedding gradient first (decoder backward), source embedding gradient last
    want = [eng.ps.g(eng._tgt_table()).data_ptr(), eng.ps.g("src_emb").data_ptr()]
    assert tables == want and (tables[0] == tables[1]) == shared
    assert tuple(eng.score(src, tgt).shape) == (src.shape[0],)
    state = eng.encoding_fn(src)
    state.begin_search(2, cap=20)
    logits, state = eng.decoding_fn(torch.zeros(2 * src.shape[0], 1, dtype=torch.int32), state, 0)
    assert tuple(logits.shape) == (2 * src.shape[0], vt)


def test_configurations_outside_the_path_fail_loudly_and_there_is_no_cpu_fallback(monkeypatch):
    """No silent degradation: a model family, head size, width or dropout rate the kernels do not serve raises a
    ZeroB200Error that names the limit; without a CUDA device the engine refuses to exist (there is no CPU path)."""
    import zero_b200.engine as E
    from zero_b200.lib import ZeroB200Error
    small = dict(hidden_size=128, embed_size=128, filter_size=256, num_heads=2, num_encoder_layer=1, num_decoder_layer=1)
    assert not torch.cuda.is_available()
    with pytest.raises(ZeroB200Error, match="no CPU path"):
        E.Engine(transformer_base(**small), 50, 50)
    for over, msg in [(dict(embed_size=64), "embed_size must equal hidden_size"),
                      (dict(num_heads=16), "head size 8 unsupported"),
                      (dict(num_heads=3), "head size"),
                      (dict(filter_size=250), "multiples of 8"),
                      (dict(model_name="rnnsearch"), "outside the hot path"),
                      (dict(model_name="transformer_l0drop"), "outside the hot path")]:
        with pytest.raises(ZeroB200Error, match=msg):
            E.ModelConfig(transformer_base(**dict(small, **over)), 50, 50)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    with pytest.raises(ZeroB200Error, match="dropout rate"):
        E.Engine(transformer_base(attention_dropout=1.0, **small), 50, 50, device="cpu")
    with pytest.raises(ZeroB200Error, match="default_dtype"):
        E.Engine(transformer_base(default_dtype="float8", **small), 50, 50, device="cpu")
    eng = E.Engine(transformer_base(**small), 50, 50, device="cpu")
    sd = eng.ps.state_dict()
    sd.pop(next(iter(sd)))
    with pytest.raises(ZeroB200Error, match="state dict mismatch"):
        eng.ps.load_state_dict(sd)
