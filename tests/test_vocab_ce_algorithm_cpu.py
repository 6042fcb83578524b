"""K6 (csrc/vocab_ce.cu + the CE epilogues of csrc/gemm2_tcgen05.cu) as an ALGORITHM, on the CPU, against the oracle.

The CUDA path never writes the [tokens, V] logits: pass 1 reduces every 128-column part of a row to
{max, sum exp(x - max), sum x, x[gold]}, the merge kernel folds the parts into the row's log-sum-exp and its smoothed
NLL  nll = lse - p x[gold] - q (sum x - x[gold]) - normaliser,  and pass 2 recomputes the logits and stores
d_logits = (exp(x - lse) - soft_label) * weight  with  weight = mask / (len_b * batch) * loss_scale.
This file restates exactly that decomposition in float32 numpy (same part width, same empty-part convention, same
formulas as the kernels' source) and checks it against oracle.smoothed_ce and torch autograd of the reference's loss
(models/transformer.py:198-211, utils/util.py:88-103) — the parity of the kernels themselves is
tests/test_kernels_gpu.py::test_vocab_ce_fused_matches_gemm_plus_softmax_ce.
"""
import math

import numpy as np
import pytest
import torch

from oracle import zero_oracle as zo

PART = 128


def part_statistics(x, labels):
    """Pass 1: [rows, V] fp32 logits -> [parts, rows, 4] {max, sum exp(x - max), sum x, x[gold] or 0}; parts is rounded
    up to whole 256-column tiles (two halves each), a half that lies past V holds {-inf, 0, 0, 0}."""
    rows, V = x.shape
    parts = 2 * ((V + 255) // 256)
    out = np.zeros((parts, rows, 4), np.float32)
    out[:, :, 0] = -np.inf
    for p in range(parts):
        lo, hi = p * PART, min(V, (p + 1) * PART)
        if lo >= hi:
            continue
        blk = x[:, lo:hi]
        m = blk.max(1)
        out[p, :, 0] = m
        out[p, :, 1] = np.exp(blk - m[:, None], dtype=np.float32).sum(1, dtype=np.float32)
        out[p, :, 2] = blk.sum(1, dtype=np.float32)
        inside = (labels >= lo) & (labels < hi)
        out[p, inside, 3] = x[inside, labels[inside]]
    return out


def merge(stats, labels, batch, seq_len, vocab, smooth, loss_scale):
    """vocab_ce_merge_kernel: running (max, sum) fold that skips empty parts, then lse / nll / row weight."""
    parts, rows, _ = stats.shape
    m = np.full(rows, -np.inf, np.float32)
    s = np.zeros(rows, np.float32)
    tot = np.zeros(rows, np.float32)
    gold = np.zeros(rows, np.float32)
    for p in range(parts):
        t = stats[p]
        live = t[:, 0] > -np.inf
        nm = np.where(live, np.maximum(m, t[:, 0]), m)
        with np.errstate(invalid="ignore"):
            s = np.where(live, s * np.exp(np.where(np.isinf(m), -np.inf, m - nm), dtype=np.float32)
                         + t[:, 1] * np.exp(t[:, 0] - nm, dtype=np.float32), s).astype(np.float32)
        m = nm
        tot += t[:, 2]
        gold += t[:, 3]
    lse = m + np.log(s, dtype=np.float32)
    lg = gold - lse
    if 0.0 < smooth < 1.0:
        n = np.float32(vocab - 1)
        p_, q_ = np.float32(1.0 - smooth), np.float32(smooth) / n
        norm = -(p_ * np.log(p_) + n * q_ * np.log(q_ + np.float32(1e-20)))
        nll = -(p_ * lg + q_ * ((tot - np.float32(vocab) * lse) - lg)) - norm
    else:
        nll = -lg
    lab = labels.reshape(batch, seq_len)
    lens = (lab != 0).sum(1)
    w = np.where((lab != 0) & (lens[:, None] > 0), loss_scale / (np.maximum(lens, 1)[:, None] * float(batch)), 0.0)
    return nll.astype(np.float32), lse.astype(np.float32), w.reshape(-1).astype(np.float32)


def d_logits(x, labels, lse, w, vocab, smooth):
    """Pass 2's epilogue: (softmax - soft label) * row weight."""
    p_, q_ = (1.0 - smooth, smooth / (vocab - 1)) if 0.0 < smooth < 1.0 else (1.0, 0.0)
    soft = np.full_like(x, q_)
    soft[np.arange(x.shape[0]), labels] = p_
    return (np.exp(x - lse[:, None]) - soft) * w[:, None]


@pytest.mark.parametrize("B,T,V,smooth,scale", [(3, 5, 128, 0.1, 1.0), (4, 7, 208, 0.1, 1.0), (2, 6, 1003, 0.1, 128.0),
                                                (2, 4, 32000, 0.1, 1.0), (3, 5, 300, 0.0, 1.0), (2, 3, 129, 0.3, 1.0)])
def test_part_statistics_merge_and_gradient_equal_the_reference_loss(B, T, V, smooth, scale):
    g = torch.Generator().manual_seed(V + B)
    logits = (torch.randn(B * T, V, generator=g) * 3.0).requires_grad_(True)
    labels = torch.randint(1, V, (B, T), generator=g)
    labels[0, T - 2:] = 0                               # padded tail of one sentence
    labels[1, 0] = V - 1                                # the last column of the last (ragged) part is a real class
    if B > 2:
        labels[2, :] = 0                                # a sentence without any target token (weight 0, no NaN)
    flat = labels.reshape(-1)
    x = logits.detach().numpy()
    stats = part_statistics(x, flat.numpy())
    assert stats.shape[0] * PART >= V and stats.shape[0] % 2 == 0
    nll, lse, w = merge(stats, flat.numpy(), B, T, V, smooth, scale)
    # the oracle: per-token smoothed NLL, per-sentence masked mean, batch mean (models/transformer.py:198-211)
    ce = zo.smoothed_ce(logits, flat, smooth)
    np.testing.assert_allclose(nll, ce.detach().numpy(), atol=2e-5 * max(1.0, float(ce.detach().abs().max())), rtol=2e-5)
    np.testing.assert_allclose(lse, torch.logsumexp(logits.detach(), -1).numpy(), rtol=1e-6, atol=1e-6)
    mask = (labels != 0).float()
    denom = mask.sum(1)
    keep = denom > 0
    per_sentence = (ce.reshape(B, T) * mask).sum(1)[keep] / denom[keep]
    loss = per_sentence.sum() / B * scale               # an all-pad sentence contributes 0 to the batch mean's sum
    (want,) = torch.autograd.grad(loss, logits)
    got = d_logits(x, flat.numpy(), lse, w, V, smooth)
    np.testing.assert_allclose(got, want.numpy(), atol=1e-6 * scale, rtol=1e-4)
    assert np.isfinite(got).all() and float(np.abs(got[flat.numpy() == 0]).max()) == 0.0
    # the loss itself from the merged rows
    np.testing.assert_allclose(float((nll * w).sum()), float(loss), rtol=1e-5)


def test_parts_past_the_vocabulary_are_empty_and_leave_the_fold_untouched():
    x = np.random.default_rng(0).standard_normal((4, 130)).astype(np.float32)
    labels = np.array([1, 129, 5, 128])
    stats = part_statistics(x, labels)
    assert stats.shape[0] == 2                          # one 256-column tile: a full half and a 2-column half
    wide = np.concatenate([stats, np.tile(np.array([-np.inf, 0, 0, 0], np.float32), (2, 4, 1))])   # two empty halves more
    a = merge(stats, labels, 1, 4, 130, 0.1, 1.0)
    b = merge(wide, labels, 1, 4, 130, 0.1, 1.0)
    for u, v in zip(a, b):
        np.testing.assert_array_equal(u, v)
    # gold in the ragged half is found there and only there
    assert stats[1, 1, 3] == x[1, 129] and stats[0, 1, 3] == 0.0 and stats[1, 3, 3] == x[3, 128]


def test_normaliser_is_the_entropy_of_the_smoothed_target():
    for V, eps in [(32000, 0.1), (208, 0.1), (129, 0.3)]:
        p, q = 1.0 - eps, eps / (V - 1)
        norm = -(p * math.log(p) + (V - 1) * q * math.log(q + 1e-20))
        soft = np.full(V, q)
        soft[3] = p
        assert abs(norm + float((soft * np.log(soft)).sum())) < 1e-9
