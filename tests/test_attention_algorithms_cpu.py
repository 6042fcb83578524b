"""The reformulations csrc/attention_tc.cu rests on, as ALGORITHMS on the CPU against the oracle's literal forms.

The tcgen05 attention kernels do not compute func.dot_attention (func.py:218-256) the way it is written: keys are walked
in 128-wide blocks with a base-2 online softmax, the backward recomputes the weights from the saved log-sum-exp, two
64-token heads share one 128 x 128 MMA, relative positions (modules/rpr.py:10-75) become GEMMs against the 2k + 1 rows
of the tables with skewed reads / bucket sums instead of a gathered [Lq, Lk, dh] tensor, and the masks are exploited
("a masked key weighs exactly 0").  Each of those steps is restated here in torch (float64 unless the claim is about
float32) and compared with oracle.attention_core and torch autograd of it.  The kernels' own parity is
tests/test_kernels_gpu.py::test_attention_tcgen05*; the oracle is pinned by the reference-executed goldens.
"""
import math

import pytest
import torch

from oracle import zero_oracle as zo
from zero_b200.params import transformer_base

F64 = torch.float64
LOG2E = 1.4426950408889634
BLK = 128


def _cfg(model="transformer", **kw):
    hp = transformer_base(hidden_size=128, embed_size=128, filter_size=128, num_heads=2, num_encoder_layer=1,
                          num_decoder_layer=1, model_name=model, scope_name=model, **kw)
    return zo.Cfg(hp, 50, 50)


def _bias(B, lq, lk, key_len, causal, inf, q_offset=0):
    bias = torch.zeros(B, 1, lq, lk, dtype=F64)
    if key_len is not None:
        bias = bias + zo.mask_bias((torch.arange(lk)[None, :] < key_len[:, None]).to(F64), inf)
    if causal:
        i = torch.arange(lq)[:, None] + q_offset
        j = torch.arange(lk)[None, :]
        bias = bias + (j > i).to(F64)[None, None] * (-inf)
    return bias


def online_softmax_attention(q, k, v, key_len, causal, inf, q_offset=0, dtype=F64):
    """The forward walk of fwd_tc_kernel for one head set: 128-key blocks, running base-2 maximum and sum, the
    accumulator rescaled when the maximum moves, blocks wholly above the causal diagonal never visited."""
    B, H, lq, dh = q.shape
    lk = k.shape[2]
    scale2 = dh ** -0.5 * LOG2E
    m = torch.full((B, H, lq), -float("inf"), dtype=dtype)
    l = torch.zeros(B, H, lq, dtype=dtype)
    acc = torch.zeros(B, H, lq, dh, dtype=dtype)
    i = (torch.arange(lq) + q_offset)[:, None]
    visited = 0
    for k0 in range(0, lk, BLK):
        j = torch.arange(k0, min(lk, k0 + BLK))[None, :]
        if causal and k0 > q_offset + lq - 1:
            continue                                   # Walk::load: inner_end stops at the diagonal block
        visited += 1
        t = (q.to(dtype) @ k[:, :, k0:k0 + BLK].to(dtype).transpose(-1, -2)) * scale2
        add = torch.zeros(B, 1, lq, j.shape[1], dtype=dtype)
        if key_len is not None:
            add = add + (j[None, None] >= key_len[:, None, None, None]).to(dtype) * (-inf * LOG2E)
        if causal:
            add = add + (j > i).to(dtype)[None, None] * (-inf * LOG2E)
        t = t + add
        m_new = torch.maximum(m, t.max(-1).values)
        p = torch.exp2(t - m_new[..., None])
        corr = torch.exp2(m - m_new)
        l = l * corr + p.sum(-1)
        acc = acc * corr[..., None] + p @ v[:, :, k0:k0 + BLK].to(dtype)
        m = m_new
    lse2 = m + torch.log2(l)                           # what the kernel saves (base 2); natural lse = lse2 / log2(e)
    return acc / l[..., None], lse2, visited


@pytest.mark.parametrize("lq,lk,causal,lens", [(64, 64, False, [64, 17, 1]), (128, 128, True, None), (100, 300, False, [300, 129, 5]),
                                               (300, 300, True, [300, 300, 200]), (1, 200, False, [200, 77, 128])])
def test_blockwise_base2_online_softmax_equals_the_literal_attention(lq, lk, causal, lens):
    c = _cfg()
    g = torch.Generator().manual_seed(lq + lk)
    B, H, dh = 3, 2, 64
    q, k, v = (torch.randn(B, H, n, dh, generator=g, dtype=F64) for n in (lq, lk, lk))
    key_len = torch.tensor(lens) if lens is not None else None
    want, w = zo.attention_core(c, {}, "", q, k, v, _bias(B, lq, lk, key_len, causal, c.inf))
    got, lse2, visited = online_softmax_attention(q, k, v, key_len, causal, c.inf)
    torch.testing.assert_close(zo.heads_merge(got), want, atol=1e-10, rtol=1e-10)
    logits = (q * dh ** -0.5) @ k.transpose(-1, -2) + _bias(B, lq, lk, key_len, causal, c.inf)
    torch.testing.assert_close(lse2 / LOG2E, torch.logsumexp(logits, -1), atol=1e-9, rtol=1e-12)
    nblocks = (lk + BLK - 1) // BLK
    assert visited == (min(nblocks, (lq - 1) // BLK + 1) if causal else nblocks)


def test_a_masked_key_weighs_exactly_zero_in_float32_and_an_all_masked_row_is_uniform():
    """func.attention_bias adds -1e8 (func.py:372-388).  In fp32, for any row that sees at least one key, the weight of
    a masked key is exp(x - 1e8 - m) == 0 exactly — which is what lets fully visible blocks skip the mask arithmetic
    and fully masked blocks be skipped altogether.  A row with EVERY key masked keeps its softmax over x - 1e8, which
    in fp32 is a softmax over logits rounded to multiples of 8 (ulp of 1e8): the kernels take the literal path there."""
    c = _cfg()
    f32 = torch.float32
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(2, 2, 16, 64, generator=g, dtype=f32) for _ in range(3))
    key_len = torch.tensor([9, 0])
    bias = zo.mask_bias((torch.arange(16)[None, :] < key_len[:, None]).to(f32), c.inf)
    _, w = zo.attention_core(c, {}, "", q, k, v, bias)
    assert float(w[0, :, :, 9:].abs().max()) == 0.0 and abs(float(w[0].sum(-1).mean()) - 1.0) < 1e-6
    assert float(torch.exp2(torch.tensor(30.0 - 1e8 * LOG2E - (-30.0), dtype=f32))) == 0.0     # the base-2 form, worst case
    # sentence 1: no key visible -> every logit becomes -1e8 exactly (|x| < 4 = half an ulp): uniform weights
    assert float((q[1] * 0.125 @ k[1].transpose(-1, -2)).abs().max()) < 4.0
    torch.testing.assert_close(w[1], torch.full_like(w[1], 1.0 / 16), atol=0, rtol=0)
    got, _, _ = online_softmax_attention(q, k, v, key_len, False, c.inf, dtype=f32)
    torch.testing.assert_close(got[1], v[1].mean(-2, keepdim=True).expand_as(got[1]), atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("causal", [False, True])
def test_backward_from_the_saved_log_sum_exp_equals_autograd(causal):
    """bwd_tc_kernel: P = 2^(t - lse2), delta = rowsum(dO * O), dS = P (dP - delta), dQ = scale dS K, dK = scale dS^T Q,
    dV = P^T dO — per 128-key block, dK / dV accumulated over the query blocks, dQ summed over the key blocks."""
    c = _cfg()
    g = torch.Generator().manual_seed(7)
    B, H, lq, lk, dh = 2, 2, 200, 260, 64
    if causal:
        lk = lq
    q, k, v = (torch.randn(B, H, n, dh, generator=g, dtype=F64).requires_grad_(True) for n in (lq, lk, lk))
    key_len = torch.tensor([lk, lk - 70])
    bias = _bias(B, lq, lk, key_len, causal, c.inf)
    out, _ = zo.attention_core(c, {}, "", q, k, v, bias)
    d_out = torch.randn(out.shape, generator=g, dtype=F64)
    want = torch.autograd.grad(out, (q, k, v), d_out)
    with torch.no_grad():
        o, lse2, _ = online_softmax_attention(q, k, v, key_len, causal, c.inf)
        dO = zo.heads_split(d_out, H)
        delta = (dO * o).sum(-1)
        scale = dh ** -0.5
        dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
        for k0 in range(0, lk, BLK):
            ks = slice(k0, min(lk, k0 + BLK))
            for q0 in range(0, lq, BLK):
                if causal and q0 + BLK - 1 < k0:
                    continue                           # Walk::load: inner_begin starts at the diagonal block
                qs = slice(q0, min(lq, q0 + BLK))
                t = (q[:, :, qs] @ k[:, :, ks].transpose(-1, -2)) * (scale * LOG2E) + bias[:, :, qs, ks] * LOG2E
                p = torch.exp2(t - lse2[:, :, qs, None])
                dp = dO[:, :, qs] @ v[:, :, ks].transpose(-1, -2)
                ds = p * (dp - delta[:, :, qs, None])
                dq[:, :, qs] += scale * ds @ k[:, :, ks]
                dk[:, :, ks] += scale * ds.transpose(-1, -2) @ q[:, :, qs]
                dv[:, :, ks] += p.transpose(-1, -2) @ dO[:, :, qs]
    for name, a, b in zip("qkv", (dq, dk, dv), want):
        torch.testing.assert_close(a, b, atol=1e-10, rtol=1e-9, msg=name)


def test_two_stacked_heads_in_one_128_row_block_equal_two_separate_heads():
    c = _cfg()
    g = torch.Generator().manual_seed(3)
    B, dh, lq, lk = 2, 64, 50, 64
    q, k, v = (torch.randn(B, 2, n, dh, generator=g, dtype=F64) for n in (lq, lk, lk))
    key_len = torch.tensor([64, 20])
    want, _ = zo.attention_core(c, {}, "", q, k, v, _bias(B, lq, lk, key_len, False, c.inf))
    # rows 0..63 = head h, 64..127 = head h + 1; positions past the sequence end arrive as zeros (TMA)
    Q, K, V = (torch.zeros(B, 128, dh, dtype=F64) for _ in range(3))
    for h in range(2):
        Q[:, 64 * h:64 * h + lq], K[:, 64 * h:64 * h + lk], V[:, 64 * h:64 * h + lk] = q[:, h], k[:, h], v[:, h]
    S = Q @ K.transpose(-1, -2) * dh ** -0.5           # ONE 128 x 128 product
    row_head = torch.arange(128)[:, None] // 64
    col_head = torch.arange(128)[None, :] // 64
    col_pos = torch.arange(128)[None, :] % 64
    keep = (row_head == col_head)[None] & (col_pos[None] < key_len[:, None, None])
    S = torch.where(keep, S, torch.full_like(S, -float("inf")))
    P = torch.softmax(S, -1)
    assert float(P[:, :64, 64:].abs().max()) == 0.0 and float(P[:, 64:, :64].abs().max()) == 0.0   # written as zeros
    O = P @ V                                          # needs no special case
    got = torch.cat([O[:, :lq], O[:, 64:64 + lq]], -1)
    torch.testing.assert_close(got, want, atol=1e-12, rtol=1e-10)


def _rel_setup(lq, lk, kmax, q_offset=0, seed=5):
    c = _cfg("transformer_rpr", max_relative_position=kmax)
    g = torch.Generator().manual_seed(seed)
    B, H, dh = 2, 2, 64
    q, k, v = (torch.randn(B, H, n, dh, generator=g, dtype=F64).requires_grad_(True) for n in (lq, lk, lk))
    P = {"a/rpr_keys/embeddings": torch.randn(2 * kmax + 1, dh, generator=g, dtype=F64).requires_grad_(True),
         "a/rpr_values/embeddings": torch.randn(2 * kmax + 1, dh, generator=g, dtype=F64).requires_grad_(True)}
    return c, g, q, k, v, P


@pytest.mark.parametrize("lq,lk,kmax,causal", [(40, 40, 4, False), (128, 128, 16, True), (70, 100, 16, False), (33, 33, 1, True)])
def test_relative_positions_as_bucket_gemms_forward_and_backward(lq, lk, kmax, causal):
    """modules/rpr.py gathers E[u(i, j)] into [Lq, Lk, dh] and contracts it twice.  The kernels never build that tensor:
       S  += (Q E_k^T)[i, u(i, j)]               one [Lq, 2k + 1] GEMM, read with the skew u(i, j) = clip(i - j, -k, k) + k
       O   = P V + W E_v,  W[i, u] = sum_{j: u(i, j) = u} P[i, j]      (bucket sums; the inner buckets hold one key each)
       dP += (dO E_v^T)[i, u(i, j)],   dQ += DSb E_k,   dE_k = DSb^T Q (scaled),   dE_v = W^T dO,   DSb = bucket sums of dS."""
    c, g, q, k, v, P = _rel_setup(lq, lk, kmax)
    key_len = torch.tensor([lk, max(1, lk - 13)])
    bias = _bias(2, lq, lk, key_len, causal, c.inf)
    out, w_ref = zo.attention_core(c, P, "a", q, k, v, bias)
    d_out = torch.randn(out.shape, generator=g, dtype=F64)
    Ek, Ev = P["a/rpr_keys/embeddings"], P["a/rpr_values/embeddings"]
    want = torch.autograd.grad(out, (q, k, v, Ek, Ev), d_out)
    nb = 2 * kmax + 1
    with torch.no_grad():
        u = zo.rel_index(lq, lk, kmax)                                   # [lq, lk] bucket of every (query, key) pair
        onehot = torch.nn.functional.one_hot(u, nb).to(F64)              # [lq, lk, nb]: the skew as a selector
        scale = 64 ** -0.5
        qs = q * scale
        QE = qs @ Ek.t()                                                 # [B, H, lq, nb]: the extra GEMM
        S = qs @ k.transpose(-1, -2) + torch.gather(QE, -1, u.expand(2, 2, lq, lk)) + bias
        Pw = torch.softmax(S, -1)
        torch.testing.assert_close(Pw, w_ref, atol=1e-12, rtol=1e-10)
        W = torch.einsum("bhij,iju->bhiu", Pw, onehot)                   # bucket sums of the weights
        # the inner buckets 1 .. 2k - 1 hold exactly one key each (a skewed copy), the two clipped ones are real sums
        inner = onehot[:, :, 1:nb - 1].sum(1)
        assert float(inner.max()) <= 1.0
        O = Pw @ v + W @ Ev
        torch.testing.assert_close(zo.heads_merge(O), out, atol=1e-11, rtol=1e-10)
        dO = zo.heads_split(d_out, 2)
        dP = dO @ v.transpose(-1, -2) + torch.gather(dO @ Ev.t(), -1, u.expand(2, 2, lq, lk))
        delta = (dO * O).sum(-1, keepdim=True)
        dS = Pw * (dP - delta)
        DSb = torch.einsum("bhij,iju->bhiu", dS, onehot)
        dq = scale * (dS @ k + DSb @ Ek)
        dk = dS.transpose(-1, -2) @ qs
        dv = Pw.transpose(-1, -2) @ dO
        dEk = torch.einsum("bhiu,bhid->ud", DSb, qs)
        dEv = torch.einsum("bhiu,bhid->ud", W, dO)
    for name, a, b in zip(("dq", "dk", "dv", "dE_k", "dE_v"), (dq, dk, dv, dEk, dEv), want):
        torch.testing.assert_close(a, b, atol=1e-10, rtol=1e-9, msg=name)


def test_relative_positions_of_a_cached_decode_step_use_the_absolute_query_position():
    """modules/rpr.py:53-54 (`last` row): at decode step t the single query sits at position t, keys at 0 .. t."""
    c, g, q, k, v, P = _rel_setup(1, 9, 3)
    full_c, _, fq, fk, fv, _ = _rel_setup(9, 9, 3)
    with torch.no_grad():
        fq[:, :, 8:9] = q
        fk.copy_(k)
        fv.copy_(v)
        step, _ = zo.attention_core(c, P, "a", q, k, v, None, q_offset=8)
        full, _ = zo.attention_core(full_c, P, "a", fq, fk, fv, _bias(2, 9, 9, None, True, c.inf))
    torch.testing.assert_close(step[:, 0], full[:, 8], atol=1e-12, rtol=1e-10)
    assert zo.rel_index(1, 9, 3, q_offset=8).tolist() == [[6, 6, 6, 6, 6, 6, 5, 4, 3]]


def test_rela_is_the_same_walk_without_maximum_and_normaliser():
    """modules/rela.py:52-75: weights = relu(logits * keep) — every key block contributes independently, no running
    statistics; masked keys contribute exactly 0 (the product with keep, not an additive -inf)."""
    c = _cfg("transformer_rela")
    g = torch.Generator().manual_seed(9)
    B, H, lq, lk, dh = 2, 2, 150, 280, 64
    q, k, v = (torch.randn(B, H, n, dh, generator=g, dtype=F64) for n in (lq, lk, lk))
    key_len = torch.tensor([280, 131])
    bias = _bias(B, lq, lk, key_len, False, c.inf)
    P = {"a/post/scale": torch.ones(128, dtype=F64), "a/post/gate": torch.zeros(128, dtype=F64)}
    _, w = zo.attention_core(c, P, "a", q, k, v, bias)
    acc = torch.zeros(B, H, lq, dh, dtype=F64)
    for k0 in range(0, lk, BLK):
        j = torch.arange(k0, min(lk, k0 + BLK))
        keep = (j[None, :] < key_len[:, None]).to(F64)[:, None, None, :]
        p = torch.relu((q * dh ** -0.5) @ k[:, :, k0:k0 + BLK].transpose(-1, -2) * keep)
        torch.testing.assert_close(p, w[:, :, :, k0:k0 + BLK], atol=1e-12, rtol=1e-10)
        acc += p @ v[:, :, k0:k0 + BLK]
    torch.testing.assert_close(acc, w @ v, atol=1e-10, rtol=1e-10)
    assert float(w[1, :, :, 131:].abs().max()) == 0.0


def test_average_attention_prefix_mean_and_its_gradient_as_scans():
    """models/transformer_aan.py:99-108: y_i = mean(x_0 .. x_i) (prefix_mean_fwd); the gradient is the suffix scan
    dx_j = sum_{i >= j} dy_i / (i + 1) (prefix_mean_bwd) — O(T d) each instead of the reference's [T, T] matmul."""
    g = torch.Generator().manual_seed(11)
    B, T, d = 3, 37, 16
    x = torch.randn(B, T, d, generator=g, dtype=F64).requires_grad_(True)
    y = zo.aan_matrix(torch.ones(B, T, dtype=F64), 1e8) @ x
    dy = torch.randn(B, T, d, generator=g, dtype=F64)
    (want,) = torch.autograd.grad(y, x, dy)
    with torch.no_grad():
        cnt = torch.arange(1, T + 1, dtype=F64)[None, :, None]
        torch.testing.assert_close(torch.cumsum(x, 1) / cnt, y, atol=1e-12, rtol=1e-10)
        dx = torch.flip(torch.cumsum(torch.flip(dy / cnt, [1]), 1), [1])
    torch.testing.assert_close(dx, want, atol=1e-12, rtol=1e-10)
    assert math.isclose(float(y.detach()[0, 0, 0]), float(x.detach()[0, 0, 0]))
