"""The algorithm behind K8 fused (csrc/vocab_topk.cu + beam_cand_kernel in csrc/beam.cu), checked on the CPU against the
oracle's restatement of search.py:147-176: reducing every 128-column part of a row's logits to {max, sum exp, top-8
(logit, column)} loses nothing the beam step needs — the 2 * beam best continuations of a sentence, with tf.nn.top_k's
order, come out of the candidates exactly as they come out of all beam * V scores.  The torch model below follows the
two kernels step by step (same formulas, same special cases); the GPU tests compare the kernels themselves with the
logits path (tests/test_kernels_gpu.py::test_beam_step_from_candidates_equals_the_logits_step)."""
import math

import numpy as np
import pytest
import torch

F32_MIN = float(np.finfo(np.float32).min)
NEG_INF = -float("inf")


def reduce_parts(x, skip_col, temperature):
    """gemm2_tcgen05.cu ce_mode 3, per row: parts = 2 * ceil(V / 256) pieces of 128 columns."""
    R, V = x.shape
    if temperature != 1.0:
        x = x / temperature
    parts = 2 * ((V + 255) // 256)
    stats = torch.zeros(parts, R, 2)
    vals = torch.full((parts, R, 8), NEG_INF)
    cols = torch.full((parts, R, 8), 0x7fffffff, dtype=torch.int64)
    for p in range(parts):
        lo, hi = p * 128, min(V, (p + 1) * 128)
        if lo >= hi:
            stats[p, :, 0] = NEG_INF
            continue
        piece = x[:, lo:hi]
        m = piece.max(1).values
        stats[p, :, 0], stats[p, :, 1] = m, torch.exp(piece - m[:, None]).sum(1)
        key = piece.clone()
        if lo <= skip_col < hi:
            key[:, skip_col - lo] = NEG_INF          # the statistics keep the column, the candidates do not
        v, i = torch.sort(key, dim=1, descending=True, stable=True)
        n = min(8, hi - lo)
        v, i = v[:, :n], i[:, :n] + lo
        keep = v > NEG_INF
        vals[p, :, :n] = torch.where(keep, v, torch.full_like(v, NEG_INF))
        cols[p, :, :n] = torch.where(keep, i, torch.full_like(i, 0x7fffffff))
    return stats, vals, cols


def step_from_candidates(stats, vals, cols, logp_prev, pen, V):
    """beam_cand_kernel + the merge of beam_sentence_tail: (scores, flat indices) [B, 2K] of each sentence."""
    B, K = logp_prev.shape
    m, s = stats[:, :, 0], stats[:, :, 1]
    lse = torch.logsumexp(torch.where(m > NEG_INF, m + torch.log(s), m), 0)      # [R]
    out_s, out_i = [], []
    for b in range(B):
        cand = []
        for k in range(K):
            row, prev = b * K + k, float(logp_prev[b, k])
            if not prev > F32_MIN:                   # not really alive: every continuation ties at prev / pen
                sc = torch.tensor(prev, dtype=torch.float32) / pen
                cand += [(float(sc), k * V + c) for c in range(2 * K)]
                continue
            v, c = vals[:, row].reshape(-1), cols[:, row].reshape(-1)
            ok = v > NEG_INF
            sc = (torch.tensor(prev, dtype=torch.float32) + (v[ok] - lse[row])) / pen
            cand += list(zip(sc.tolist(), (k * V + c[ok]).tolist()))
        cand.sort(key=lambda e: (-e[0], e[1]))
        out_s.append([e[0] for e in cand[:2 * K]])
        out_i.append([e[1] for e in cand[:2 * K]])
    return torch.tensor(out_s, dtype=torch.float32), torch.tensor(out_i)


def step_from_logits(x, logp_prev, pen, t, eos, inf, temperature):
    """search.py:147-176 as oracle/zero_oracle.py:beam_search restates it."""
    from oracle import zero_oracle as zo
    B, K = logp_prev.shape
    V = x.shape[1]
    logits = x / temperature
    lp = logits - torch.logsumexp(logits, -1, keepdim=True)
    if t < 1:
        lp = lp + (torch.arange(V) == eos).float()[None, :] * -inf
    cs = (logp_prev[:, :, None] + lp.reshape(B, K, V)) / pen
    return zo.top_k(cs.reshape(B, K * V), 2 * K)


def _logp(B, K, kind, gen):
    if kind == "first":                              # search.py:46-47
        lp = torch.full((B, K), F32_MIN)
        lp[:, 0] = 0.0
        return lp
    lp = -torch.rand(B, K, generator=gen) * 8.0
    lp = torch.sort(lp, 1, descending=True).values
    if kind == "dead":                               # sentence 0 ran past its max_len: a_s * pen overflowed
        lp[0] = NEG_INF
    return lp


@pytest.mark.parametrize("V", [208, 1000, 4099])
@pytest.mark.parametrize("K", [1, 2, 4])
@pytest.mark.parametrize("kind,t,temperature", [("first", 0, 1.0), ("alive", 3, 1.0), ("alive", 2, 0.7), ("dead", 9, 1.0)])
def test_candidates_give_the_sentences_top_2k_exactly(V, K, kind, t, temperature):
    gen = torch.Generator().manual_seed(V * 31 + K * 7 + t)
    B, eos, inf = 3, 2, 1e8
    x = torch.randn(B * K, V, generator=gen) * 3.0
    logp = _logp(B, K, kind, gen)
    pen = torch.pow(torch.tensor((5.0 + float(t + 1)) / 6.0), 0.6)
    want_s, want_i = step_from_logits(x, logp, pen, t, eos, inf, temperature)
    stats, vals, cols = reduce_parts(x, eos if t < 1 else -1, temperature)
    got_s, got_i = step_from_candidates(stats, vals, cols, logp, pen, V)
    assert torch.equal(got_i, want_i)
    torch.testing.assert_close(got_s, want_s, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("t", [0, 1])
def test_a_frequency_sorted_vocabulary_keeps_its_whole_top_2k_in_one_part(t):
    """Real BPE vocabularies are sorted by frequency: the best first words all sit in the first 128 columns, next to
    EOS (id 2).  At t = 0 EOS scores high as a logit and must not take one of the part's 8 candidate slots."""
    gen = torch.Generator().manual_seed(5)
    B, K, V, eos = 2, 4, 1000, 2
    x = torch.randn(B * K, V, generator=gen)
    x[:, :12] += 12.0                                # the 2K = 8 best are among columns 0..11 of every row
    x[:, eos] += 5.0                                 # ... and EOS is the best logit of all
    logp = _logp(B, K, "first" if t == 0 else "alive", gen)
    pen = torch.pow(torch.tensor((5.0 + float(t + 1)) / 6.0), 0.6)
    want_s, want_i = step_from_logits(x, logp, pen, t, eos, 1e8, 1.0)
    stats, vals, cols = reduce_parts(x, eos if t < 1 else -1, 1.0)
    got_s, got_i = step_from_candidates(stats, vals, cols, logp, pen, V)
    assert torch.equal(got_i, want_i)
    assert (t == 0) == (not bool((want_i % V == eos).any()))
    torch.testing.assert_close(got_s, want_s, rtol=1e-5, atol=1e-5)


def test_equal_logits_keep_top_ks_lower_column_first():
    gen = torch.Generator().manual_seed(9)
    B, K, V = 1, 2, 300
    x = torch.round(torch.randn(B * K, V, generator=gen) * 2.0)       # many exact ties
    logp = torch.tensor([[-0.5, -1.5]])
    pen = torch.tensor(1.0)
    want_s, want_i = step_from_logits(x, logp, pen, 4, 2, 1e8, 1.0)
    got_s, got_i = step_from_candidates(*reduce_parts(x, -1, 1.0), logp, pen, V)
    assert torch.equal(got_i, want_i)
    assert math.isclose(float(got_s[0, 0]), float(want_s[0, 0]), rel_tol=1e-6)
