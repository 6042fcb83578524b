"""The reference's built-in self-consistency paths and the independent cross-checks of section 8c,
run on the oracle in float64 — results that must hold whatever TensorFlow computes, so they pin the restatement from a
second side (the first is the reference-executed goldens of tests/test_oracle_golden.py):

  1. cached incremental decoding == the full decoder on the prefix (search_mode "cache" vs "dev",
     search.py:129-142, models/transformer.py:272-281), every model family;
  2. average attention three ways: mask matmul (aan_mask=True) == cumulative sum / count (False) == the running sum of
     the decode cache (models/transformer_aan.py:99-112, func.py:390-398);
  3. score_fn's per-sentence NLL == the mean over target positions of -log_softmax(step logits)[gold]
     (models/transformer.py:235-249 vs search.py:148);
  4. the mean of the towers' gradients == the gradient of the mean of the towers' losses (main.py:42-43,
     utils/parallel.py:196);
  5. attention, LayerNorm and the smoothed cross-entropy against torch's own operators (a second, independently
     written CPU path).
CPU only; sizes of the golden models.
"""
import math

import pytest
import torch

from oracle import zero_oracle as zo
from zero_b200.params import transformer_base

F64 = torch.float64
FAMILIES = {
    "transformer": dict(),
    "transformer_rpr": dict(max_relative_position=3),
    "transformer_rela": dict(),
    "transformer_aan": dict(use_ffn=False, aan_mask=True),
    "transformer_aan+ffn": dict(use_ffn=True, aan_mask=False),
    "transformer_fuse": dict(),
}


def _cfg(name, V=53, **over):
    model = name.split("+")[0]
    hp = transformer_base(hidden_size=32, embed_size=32, filter_size=48, num_heads=2, num_encoder_layer=2,
                          num_decoder_layer=2, model_name=model, scope_name=model, **dict(FAMILIES[name], **over))
    return zo.Cfg(hp, V, V)


def _params(c, seed=5):
    P = {k: v.to(F64) for k, v in zo.init_params(c, seed=seed).items()}
    g = torch.Generator().manual_seed(seed + 1)
    for k, v in P.items():                       # biases / LN offsets are zero-initialised: make every term count
        if v.dim() == 1:
            v += 0.1 * torch.randn(v.shape, generator=g, dtype=F64)
    return P


def _batch(V=53, B=4, S=9, T=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(3, V, (B, S), generator=g)
    tgt = torch.randint(3, V, (B, T), generator=g)
    src[:, -1] = 2
    tgt[:, -1] = 2
    src[1, 5:] = 0
    src[1, 4] = 2
    tgt[2, 4:] = 0
    tgt[2, 3] = 2
    return src, tgt


def _teacher_forced_steps(c, P, src, tgt):
    """Cached decoding driven with the gold prefix: step t is fed the gold token t - 1 (zeros at t = 0,
    models/transformer.py:113-116) and returns the logits of position t."""
    enc_fn, dec_fn = zo.make_infer_fns(c, P, F64)
    state = enc_fn(src)
    steps = []
    for t in range(tgt.shape[1]):
        prev = torch.zeros(tgt.shape[0], 1, dtype=tgt.dtype) if t == 0 else tgt[:, t - 1:t]
        logits, state = dec_fn(prev, state, t)
        steps.append(logits)
    return torch.stack(steps, 1)                 # [B, T, V]


@pytest.mark.parametrize("name", sorted(FAMILIES))
def test_cached_decoding_equals_the_full_decoder_on_the_prefix(name):
    c = _cfg(name)
    P = _params(c)
    src, tgt = _batch()
    full = tgt.clone()
    full[2, 4:] = 7                              # no padding inside the compared prefix (pads change the AAN counts)
    _, logits, _, _ = zo.train_loss(c, P, src, full, F64)
    steps = _teacher_forced_steps(c, P, src, full)
    torch.testing.assert_close(steps, logits.reshape(steps.shape), atol=1e-9, rtol=1e-9)


def test_average_attention_three_ways():
    src, tgt = _batch()
    c_mask, c_sum = _cfg("transformer_aan", aan_mask=True), _cfg("transformer_aan", aan_mask=False)
    P = _params(c_mask)
    l1, g1, _, _ = zo.train_loss(c_mask, P, src, tgt, F64)
    l2, g2, _, _ = zo.train_loss(c_sum, P, src, tgt, F64)
    valid = (tgt != 0).reshape(-1)
    # func.attention_bias("aan") zeroes the rows of padded positions, the cumulative form keeps averaging there: the
    # two agree wherever the loss looks (non-pad targets), which is all the model ever uses
    torch.testing.assert_close(g1[valid], g2[valid], atol=1e-7, rtol=1e-7)   # the mask's softmax of -1e8: exp(-1e8) == 0
    assert abs(float(l1) - float(l2)) < 1e-9
    full = tgt.clone()
    full[2, 4:] = 7
    _, gm, _, _ = zo.train_loss(c_mask, P, src, full, F64)
    steps = _teacher_forced_steps(c_sum, P, src, full)                        # running sum / (t + 1)
    torch.testing.assert_close(steps, gm.reshape(steps.shape), atol=1e-7, rtol=1e-7)
    # the averaging matrix itself: row i = 1 / (i + 1) over columns 0..i (func.py:390-398)
    m = zo.aan_matrix(torch.ones(1, 5, dtype=F64), 1e8)[0]
    want = torch.tril(torch.ones(5, 5, dtype=F64)) / torch.arange(1, 6, dtype=F64)[:, None]
    torch.testing.assert_close(m, want, atol=1e-12, rtol=0)


@pytest.mark.parametrize("name", ["transformer", "transformer_rpr", "transformer_aan", "transformer_fuse"])
def test_score_fn_equals_the_summed_step_log_probabilities(name):
    c = _cfg(name)
    P = _params(c)
    src, tgt = _batch()
    full = tgt.clone()
    full[2, 4:] = 7
    nll = zo.score(c, P, src, full, F64)
    lp = torch.log_softmax(_teacher_forced_steps(c, P, src, full), -1)
    gold = lp.gather(-1, full[:, :, None]).squeeze(-1)
    torch.testing.assert_close(nll, -gold.mean(1), atol=1e-9, rtol=1e-9)
    # with padding: the padded positions carry no weight (models/transformer.py:208-210)
    nll_pad = zo.score(c, P, src, tgt, F64)
    if name in ("transformer", "transformer_rpr"):       # causal self-attention: the prefix does not see the padding
        torch.testing.assert_close(nll_pad[2], -gold[2, :4].sum() / 4.0 + (gold[2, 3] - lp[2, 3, 2]) / 4.0,
                                   atol=1e-9, rtol=1e-9)


def test_mean_of_tower_gradients_is_the_gradient_of_the_mean_loss():
    c = _cfg("transformer")
    P = {k: v.requires_grad_(True) for k, v in _params(c).items()}
    towers = [_batch(seed=s, B=3 + s) for s in range(3)]
    per_tower = []
    for src, tgt in towers:
        loss = zo.train_loss(c, P, src, tgt, F64)[0]
        per_tower.append(torch.autograd.grad(loss, list(P.values())))
    mean_loss = sum(zo.train_loss(c, P, s, t, F64)[0] for s, t in towers) / len(towers)
    joint = torch.autograd.grad(mean_loss, list(P.values()))
    for k, j, *gs in zip(P, joint, *per_tower):
        torch.testing.assert_close(j, sum(gs) / len(gs), atol=1e-12, rtol=1e-9, msg=k)


def test_attention_core_equals_torch_scaled_dot_product_attention():
    c = _cfg("transformer")
    g = torch.Generator().manual_seed(2)
    B, h, L, dh = 3, c.h, 7, c.d // c.h
    q, k, v = (torch.randn(B, h, L, dh, generator=g, dtype=F64) for _ in range(3))
    lens = torch.tensor([7, 4, 1])
    keep = (torch.arange(L)[None, :] < lens[:, None])
    bias = zo.mask_bias(keep.to(F64), c.inf)                          # [B, 1, 1, L] additive -1e8
    out, weights = zo.attention_core(c, {}, "unused", q, k, v, bias)
    want = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=keep[:, None, None, :])
    torch.testing.assert_close(out, zo.heads_merge(want), atol=1e-9, rtol=1e-9)
    assert float(weights[1, :, :, 4:].abs().max()) == 0.0                # exp(-1e8) is exactly 0: masked keys weigh nothing
    causal = zo.causal_bias(L, c.inf, F64)
    out, _ = zo.attention_core(c, {}, "unused", q, k, v, causal)
    want = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True)
    torch.testing.assert_close(out, zo.heads_merge(want), atol=1e-9, rtol=1e-9)


def test_layer_norm_equals_torch_layer_norm():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 6, 32, generator=g, dtype=F64) * 3 + 1
    P = {"ln/scale": torch.randn(32, generator=g, dtype=F64), "ln/offset": torch.randn(32, generator=g, dtype=F64)}
    got = zo.layer_norm(P, "ln", x, 1e-8)
    want = torch.nn.functional.layer_norm(x, (32,), P["ln/scale"], P["ln/offset"], eps=1e-8)   # biased variance
    torch.testing.assert_close(got, want, atol=1e-10, rtol=1e-10)


@pytest.mark.parametrize("V,eps", [(53, 0.1), (1000, 0.1), (32000, 0.1), (53, 0.3)])
def test_smoothed_cross_entropy_equals_torch_label_smoothing_after_the_change_of_variable(V, eps):
    """Zero smooths with p = 1 - eps on the gold class and q = eps / (V - 1) elsewhere (utils/util.py:88-103); torch
    spreads alpha / V over ALL classes.  alpha = eps V / (V - 1) makes the two targets identical; Zero then subtracts
    the entropy of the target so that a perfect prediction scores 0 (models/transformer.py:203-205)."""
    g = torch.Generator().manual_seed(4)
    logits = torch.randn(11, V, generator=g, dtype=F64) * 2
    labels = torch.randint(0, V, (11,), generator=g)
    got = zo.smoothed_ce(logits, labels, eps)
    alpha = eps * V / (V - 1.0)
    xent = torch.nn.functional.cross_entropy(logits, labels, reduction="none", label_smoothing=alpha)
    p, q = 1.0 - eps, eps / (V - 1.0)
    normaliser = -(p * math.log(p) + (V - 1.0) * q * math.log(q + 1e-20))
    torch.testing.assert_close(got, xent - normaliser, atol=1e-10, rtol=1e-10)
    # the normaliser is the loss of the perfect prediction: logits = log(soft target) gives exactly 0
    soft = torch.full((1, V), q, dtype=F64)
    soft[0, 5] = p
    assert abs(float(zo.smoothed_ce(torch.log(soft), torch.tensor([5]), eps))) < 1e-9
    torch.testing.assert_close(zo.smoothed_ce(logits, labels, 0.0),
                               torch.nn.functional.cross_entropy(logits, labels, reduction="none"), atol=1e-10, rtol=1e-10)


@pytest.mark.parametrize("name", ["transformer", "transformer_aan", "transformer_fuse"])
def test_empty_tower_has_zero_loss_and_zero_gradients(name):
    """models/transformer.py:213-216: a tower that received no sentence (fewer batches than devices at the end of an
    epoch, main.py:287-294) contributes a loss of exactly 0 — and therefore nothing to the averaged gradients."""
    c = _cfg(name)
    P = {k: v.requires_grad_(True) for k, v in _params(c).items()}
    src = torch.zeros(0, 5, dtype=torch.long)
    tgt = torch.zeros(0, 4, dtype=torch.long)
    loss, logits, per_sample, _ = zo.train_loss(c, P, src, tgt, F64)
    assert float(loss.detach()) == 0.0 and per_sample.numel() == 0 and logits.shape[0] == 0
    grads = torch.autograd.grad(loss, list(P.values()), allow_unused=True)
    assert all(g is None or float(g.abs().max()) == 0.0 for g in grads)
