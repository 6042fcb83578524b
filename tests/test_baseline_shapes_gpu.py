"""Parity of the CUDA path against the CPU oracle AT THE SHAPES OF BASELINE.json's configs (the model-level golden
vectors of tests/golden are 2-layer, d = 128 models with short sentences; these tests run the real 6+6, d = 512,
V = 32k models, reduced only in batch size so the fp32 oracle finishes in seconds on the host cores):

  C2  configs[1]  transformer 6+6, S = T = 64, B = 8 with padded rows  loss, logits, every gradient, scores
  C4  configs[3]  transformer_rpr 6+6, S = T = 128, k = 16, B = 2      loss, logits, every gradient
  C3  configs[2]  transformer_aan 6+6, beam 4, B = 8, src len 64       step logits, beams bit-exact on the same logits
  C5  configs[4]  DS-Init encoder, S = 1024 (4 layers of the 24, B = 2) encoder output, loss, logits, gradients

Tolerance (BASELINE.json north_star): logits within 1e-2 * (logit range) in bf16 compute with fp32 accumulation; the
achieved error is printed (run with -s) and asserted at exactly that bound, not a multiple of it.  Beam indices are
bit-exact when both sides consume the same logits.  Reference lines: models/transformer.py:15-218,
models/transformer_rpr.py:47-170, models/transformer_aan.py:92-260, search.py:19-275.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

V = 32000


def _batch(seed, B, S, T, pads):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(3, V, (B, S), generator=g)
    tgt = torch.randint(3, V, (B, T), generator=g)
    src[:, -1] = 2
    tgt[:, -1] = 2
    for which, row, n in pads:
        t = src if which == "s" else tgt
        t[row, n - 1] = 2
        t[row, n:] = 0
    return src, tgt


def _oracle_and_engine(hp, seed):
    from oracle import zero_oracle as zo
    from zero_b200.engine import Engine
    eng = Engine(hp, V, V)
    eng.ps.init_random(seed)
    # the bias / LN parameters start at exactly 0 / 1: move them so that their part of the computation is not vacuous
    g = torch.Generator().manual_seed(seed + 1)
    sd = eng.ps.state_dict()
    for k in sd:
        leaf = k.rsplit("/", 1)[1]
        if leaf in ("b_0", "offset", "bias", "scale", "gate"):
            sd[k] = sd[k] + 0.05 * torch.randn(sd[k].shape, generator=g)
    eng.ps.load_state_dict(sd)
    c = zo.Cfg(hp, V, V)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    return zo, c, P, eng


def _check_train(tag, zo, c, P, eng, src, tgt, min_cos=0.99, max_rel=0.10, check_score=True):
    loss = eng.forward_backward(src, tgt)
    torch.cuda.synchronize()
    got_grads = eng.ps.grad_dict()
    _, per_sample, logits = eng.train_loss(src, tgt)
    logits, per_sample = logits.float().cpu().clone(), per_sample.cpu().clone()
    want_loss, want_logits, want_ps, _ = zo.train_loss(c, P, src, tgt)
    names = sorted(P)
    want_grads = dict(zip(names, torch.autograd.grad(want_loss, [P[n] for n in names], allow_unused=True)))
    want_logits = want_logits.detach().reshape(-1, want_logits.shape[-1])
    scale = max(1.0, float(want_logits.abs().max()))
    err = float((logits - want_logits).abs().max())
    rel = float((logits - want_logits).double().norm() / want_logits.double().norm())
    print("%s: loss %.5f (oracle %.5f)  logits max-abs err %.4f = %.2e of the logit range %.2f (bound 1e-2), "
          "rel-Frobenius %.2e" % (tag, float(loss[0]), float(want_loss), err, err / scale, scale, rel))
    assert abs(float(loss[0]) - float(want_loss)) < 1e-2, (float(loss[0]), float(want_loss))
    assert err <= 1e-2 * scale, "logits max abs err %.4f exceeds 1e-2 x the logit range %.2f" % (err, scale)
    assert rel < 1e-2
    np.testing.assert_allclose(per_sample.numpy(), want_ps.detach().numpy(), atol=2e-2, rtol=5e-3)
    worst_cos, worst_rel = ("", 1.0), ("", 0.0)
    for k in names:
        g = want_grads[k]
        if g is None or float(g.abs().max()) < 1e-7:
            assert float(got_grads[k].abs().max()) < 2e-3, k    # analytically zero (e.g. k_map/b_0)
            continue
        a, b = got_grads[k].double().flatten(), g.double().flatten()
        cos = float(torch.nn.functional.cosine_similarity(a, b, dim=0))
        r = float((a - b).norm() / (b.norm() + 1e-30))
        if cos < worst_cos[1]:
            worst_cos = (k, cos)
        if r > worst_rel[1]:
            worst_rel = (k, r)
    print("%s: %d gradients, worst cosine %.5f (%s), worst relative error %.4f (%s)" % (
        tag, len(names), worst_cos[1], worst_cos[0], worst_rel[1], worst_rel[0]))
    assert worst_cos[1] > min_cos, worst_cos
    assert worst_rel[1] < max_rel, worst_rel
    if check_score:
        with torch.no_grad():
            want_sc = zo.score(c, {k: v.detach() for k, v in P.items()}, src, tgt)
        np.testing.assert_allclose(eng.score(src, tgt).cpu().numpy(), want_sc.numpy(), atol=5e-2, rtol=5e-3)


def test_c2_transformer_base_64x64_against_the_oracle():
    """BASELINE configs[1] at its own layer count, width, vocabulary and sentence length (B = 8 instead of 64): the
    64-token attention kernels, the 32k-way vocabulary projection + smoothed CE and every GEMM shape of the
    benchmarked step, against the fp32 oracle.  Rows 1 / 2 are padded (key-length masks, loss masks)."""
    from zero_b200.params import transformer_base
    hp = transformer_base()
    zo, c, P, eng = _oracle_and_engine(hp, 11)
    src, tgt = _batch(5, 8, 64, 64, [("s", 1, 40), ("t", 2, 50), ("s", 5, 17), ("t", 5, 23)])
    _check_train("C2", zo, c, P, eng, src, tgt)


def test_c4_relative_positions_len128_against_the_oracle():
    """BASELINE configs[3]: transformer_rpr 6+6, d = 512, src / tgt length 128, max_relative_position 16, B = 2
    (modules/rpr.py:10-75 on encoder self, decoder self AND decoder cross attention)."""
    from zero_b200.params import transformer_base
    hp = transformer_base(model_name="transformer_rpr", scope_name="transformer_rpr", max_relative_position=16)
    zo, c, P, eng = _oracle_and_engine(hp, 12)
    src, tgt = _batch(6, 2, 128, 128, [("s", 1, 90), ("t", 1, 77)])
    _check_train("C4", zo, c, P, eng, src, tgt)
    for l in range(6):
        for key in ("enc%d.self" % l, "dec%d.self" % l, "dec%d.cross" % l):
            assert float(eng.ps.g(key + ".rpr_k").abs().sum()) > 0, key


def test_c5_deep_init_encoder_len1024_against_the_oracle():
    """BASELINE configs[4] reduced to what the CPU oracle does in seconds: depth-scaled initialisation
    (models/transformer.py:38-45), source length 1024 (multi-tile attention with online softmax), 4 of the 24 encoder
    layers + 2 decoder layers, target length 64, B = 2 with one source padded to 700 tokens."""
    from zero_b200.params import transformer_base
    hp = transformer_base(num_encoder_layer=4, num_decoder_layer=2, deep_transformer_init=True,
                          initializer="uniform_unit_scaling", initializer_gain=1.0)
    zo, c, P, eng = _oracle_and_engine(hp, 13)
    src, tgt = _batch(7, 2, 1024, 64, [("s", 1, 700), ("t", 0, 41)])
    _check_train("C5", zo, c, P, eng, src, tgt, check_score=False)
    # encoder output on its own (the part configs[4] is about)
    enc, _ = eng.encode(eng._prep_ids(src, eng.device))
    enc = enc.float().cpu().view(2, 1024, 512)
    with torch.no_grad():
        want = zo.encoder(c, {k: v.detach() for k, v in P.items()}, src)["encodes"]
    mask = (src != 0)
    err = float((enc - want)[mask].abs().max())
    print("C5: encoder output max-abs err %.4f (values up to %.2f)" % (err, float(want.abs().max())))
    # measured 0.060 on values up to 4.94 (1.2e-2 of the range, profiles/r02_golden_logit_errors.log): the encoder output
    # is a bf16 tensor, two of its ulps at that magnitude; the bound is 1.5x the measurement rather than the 1e-2 that
    # the logits of the same run meet (2.4e-3)
    assert err < 1.8e-2 * max(1.0, float(want.abs().max()))


def test_c3_average_attention_beam4_vocab32k_against_the_oracle():
    """BASELINE configs[2]: transformer_aan 6+6, beam 4, B = 8, source length 64 (two padded rows), V = 32k.
    (1) the first step's logits against the oracle's decoding_fn; (2) the oracle's beam search replayed on the
    logits the CUDA path produced returns bit-identical sequences — the fused beam-step kernel at the real
    vocabulary size against search.py:115-238, step for step; (3) scores to 1e-4."""
    from oracle import zero_oracle as zo
    from zero_b200 import search
    from zero_b200.params import SimpleVocab, transformer_base
    hp = transformer_base(model_name="transformer_aan", scope_name="transformer_aan", use_ffn=False, aan_mask=True,
                          beam_size=4, decode_length=6, decode_alpha=0.6)
    zo, c, P, eng = _oracle_and_engine(hp, 14)
    hp.add_hparam("src_vocab", SimpleVocab(V))
    hp.add_hparam("tgt_vocab", SimpleVocab(V))
    hp.add_hparam("decode_graph", False)
    src, _ = _batch(8, 8, 64, 8, [("s", 1, 40), ("s", 6, 9)])
    eng.decode_length = hp.decode_length
    recorded = []

    def dec_fn(tok, state, t):
        lg, st = eng.decoding_fn(tok, state, t)
        recorded.append(lg.detach().float().cpu().clone())
        return lg, st

    out = search.beam_search({"source": src}, eng.encoding_fn, dec_fn, hp)
    torch.cuda.synchronize()
    Pd = {k: v.detach() for k, v in P.items()}
    enc_fn, dec_oracle = zo.make_infer_fns(c, Pd)
    steps = {}
    with torch.no_grad():
        zo.beam_search(c, src[:2], enc_fn, dec_oracle, logits_hook=lambda t, lg: steps.setdefault(t, lg.clone()))
    want0 = steps[0]                                   # [2 * beam, V]: at t = 0 every beam holds the same prefix
    got0 = recorded[0][:want0.shape[0]]
    scale = max(1.0, float(want0.abs().max()))
    err = float((got0 - want0).abs().max())
    print("C3: step-0 logits max-abs err %.4f = %.2e of the logit range %.2f (bound 1e-2)" % (err, err / scale, scale))
    assert err <= 1e-2 * scale
    calls = {"n": 0}

    def dec_replay(tok, state, time):
        i = max(calls["n"] - 1, 0)
        calls["n"] += 1
        return recorded[min(i, len(recorded) - 1)], {"dummy": state["dummy"], "decoder": {"state": {}}}

    want = zo.beam_search(c, src, lambda s: {"dummy": torch.zeros(s.shape[0], 1)}, dec_replay)
    assert want["steps"] == len(recorded)
    np.testing.assert_array_equal(out["seq"].cpu().numpy(), want["seq"].numpy())
    np.testing.assert_allclose(out["score"].cpu().numpy(), want["score"].numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["transformer", "transformer_len40", "transformer_aan", "transformer_rpr_len40"])
def test_cuda_beams_against_the_reference_run_beams(name):
    """The beams the reference's own search.py produced (tests/golden, fp32) against the beams of the CUDA path (bf16
    logits): near-ties may legitimately resolve differently, so the comparison is on what is robust to that — the
    score of every sentence's best hypothesis (1e-2 of its magnitude) — and the share of identical top-1 sequences
    is printed and must be a clear majority."""
    from tests.golden_util import load_golden
    from zero_b200 import search
    from zero_b200.engine import Engine
    from zero_b200.params import SimpleVocab
    z, hp, variables, grads, vs, vt = load_golden(name)
    eng = Engine(hp, vs, vt)
    eng.ps.load_state_dict(variables)
    hp.add_hparam("src_vocab", SimpleVocab(vs))
    hp.add_hparam("tgt_vocab", SimpleVocab(vt))
    hp.add_hparam("decode_graph", False)
    eng.decode_length = hp.decode_length
    out = search.beam_search({"source": torch.from_numpy(z["source"])}, eng.encoding_fn, eng.decoding_fn, hp)
    seq, score = out["seq"].cpu().numpy(), out["score"].cpu().numpy()
    want_seq, want_score = z["beam_seq"], z["beam_score"]
    L = min(seq.shape[-1], want_seq.shape[-1])
    same = [bool(np.array_equal(seq[b, 0, :L], want_seq[b, 0, :L])) for b in range(seq.shape[0])]
    print("%s: %d / %d top-1 hypotheses identical to the reference's; top-1 scores %s vs %s" % (
        name, sum(same), len(same), np.round(score[:, 0], 3), np.round(want_score[:, 0], 3)))
    np.testing.assert_allclose(score[:, 0], want_score[:, 0], atol=3e-2, rtol=1e-2)
    assert sum(same) * 2 > len(same)
