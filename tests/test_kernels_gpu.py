"""Per-kernel parity on the GPU, through the C ABI, against plain torch fp32 restatements of the same op
(inputs rounded to bf16 first so the comparison isolates the kernel's own arithmetic)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

bf16, f32 = torch.bfloat16, torch.float32


def dev():
    return torch.device("cuda")


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(bf16).to(dev())


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (200, 136, 72), (4096, 1536, 512), (1000, 1000, 520)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_gemm_layouts(m, n, k, a_mn, b_mn):
    from zero_b200 import ops
    A = rnd(m, k, seed=1)
    B = rnd(n, k, seed=2)
    a_st = A.t().contiguous() if a_mn else A
    b_st = B.t().contiguous() if b_mn else B
    # pad pitches to multiples of 8 elements
    def pad(t):
        c = (t.shape[1] + 7) // 8 * 8
        buf = torch.zeros(t.shape[0], c, dtype=t.dtype, device=t.device)
        buf[:, :t.shape[1]] = t
        return buf[:, :t.shape[1]]
    a_st, b_st = pad(a_st), pad(b_st)
    out = pad(torch.zeros(m, n, dtype=f32, device=dev()))
    ops.gemm(a_st, b_st, out, a_mn, b_mn, m=m, n=n, k=k)
    ref = A.float() @ B.float().t()
    torch.testing.assert_close(out, ref, atol=2e-3, rtol=2e-3)


@pytest.mark.parametrize("m,n,k", [(256, 512, 512), (256, 2048, 512), (256, 512, 2048), (256, 1024, 1024),
                                   (100, 128, 64), (1, 64, 192), (384, 1536, 512), (37, 384, 128)])
@pytest.mark.parametrize("b_mn", [1, 0])
@pytest.mark.parametrize("split", ["1", "0"])
def test_gemm_skinny_rows(m, n, k, b_mn, split, monkeypatch):
    """The opt-in decode-step GEMM (gemm_skinny.cu: m <= 384 rows, 64 x 64 tiles, mma.sync, k split over a
    thread-block cluster with a DSMEM reduction) on the projections of one cached decode step at configs[2]
    (256 rows) and on ragged row counts: plain fp32 output, then the bias + relu epilogue into a strided bf16
    destination (func.linear writing into the [rows, cap, 3d] cache)."""
    from zero_b200 import ops
    import zero_b200.lib as L
    monkeypatch.setenv("ZB_SKINNY_GEMM", "1")
    monkeypatch.setenv("ZB_SKINNY_SPLIT", split)
    before = L.path_launch_count("gemm_skinny")
    x = rnd(m, k, seed=11)
    w = rnd(k, n, scale=0.05, seed=12)            # [in, out] like func.linear's W
    b_st = w if b_mn else w.t().contiguous()      # MN-major = [k][n]; K-major = [n][k]
    out = torch.zeros(m, n, dtype=f32, device=dev())
    ops.gemm(x, b_st, out, L.ZB_K_MAJOR, L.ZB_MN_MAJOR if b_mn else L.ZB_K_MAJOR)
    ref = x.float() @ w.float()
    torch.testing.assert_close(out, ref, atol=2e-3, rtol=2e-3)
    if b_mn:
        bias = torch.randn(n, device=dev())
        buf = torch.full((m, 3, n + 64), 7.0, dtype=bf16, device=dev())
        ops.linear_fwd(x, w, bias, buf[:, 1, :n], relu=True)
        torch.testing.assert_close(buf[:, 1, :n].float(), torch.relu(ref + bias), atol=3e-2, rtol=2e-2)
        assert float((buf[:, 0] - 7).abs().max()) == 0 and float((buf[:, 2] - 7).abs().max()) == 0
        assert float((buf[:, 1, n:] - 7).abs().max()) == 0
    assert L.path_launch_count("gemm_skinny") > before      # the switch really routed the problem to this kernel


def test_gemm_epilogues_and_splitk():
    from zero_b200 import ops
    m, n, k = 512, 1536, 4096
    x, dy = rnd(k, m, seed=3), rnd(k, n, seed=4)
    acc = torch.randn(m, n, device=dev())
    want = acc + x.float().t() @ dy.float()
    ops.linear_wgrad(x, dy, acc)
    torch.testing.assert_close(acc, want, atol=2e-2, rtol=2e-3)
    # bias + relu, bf16 out
    x2, w, b = rnd(300, 512, seed=5), rnd(512, 2048, scale=0.05, seed=6), torch.randn(2048, device=dev())
    out = torch.empty(300, 2048, dtype=bf16, device=dev())
    ops.linear_fwd(x2, w, b, out, relu=True)
    ref = torch.relu(x2.float() @ w.float() + b)
    torch.testing.assert_close(out.float(), ref, atol=3e-2, rtol=2e-2)
    # relu-masked dgrad
    dh = torch.empty(300, 512, dtype=bf16, device=dev())
    mask = rnd(300, 512, seed=7)
    ops.linear_dgrad(out, w, dh, relu_mask=mask)  # dh = (dy @ W^T) * (mask > 0), dy := out
    ref2 = (out.float() @ w.float().t()) * (mask.float() > 0)
    torch.testing.assert_close(dh.float(), ref2, atol=0.15, rtol=2e-2)


@pytest.mark.parametrize("tokens", [4096, 1000])
def test_gemm_grouped_and_colsum_grouped(tokens):
    """The weight gradients + bias gradients of one decoder layer through the grouped entry points (one launch
    each) equal the fp32 reference and the per-problem launches; a group that does not qualify (small / bf16
    problems mixed in) falls back to per-problem launches with the same results."""
    from zero_b200 import ops
    shapes = [(512, 1536), (512, 512), (512, 512), (512, 1024), (512, 512), (512, 2048), (2048, 512)]
    xs = [rnd(tokens, i, seed=10 + j) for j, (i, o) in enumerate(shapes)]
    dys = [rnd(tokens, o, seed=30 + j) for j, (i, o) in enumerate(shapes)]
    accs = [torch.randn(i, o, device=dev()) for i, o in shapes]
    want = [a + x.float().t() @ dy.float() for a, x, dy in zip(accs, xs, dys)]
    single = [a.clone() for a in accs]
    for s, x, dy in zip(single, xs, dys):
        ops.linear_wgrad(x, dy, s)
    ops.gemm_grouped([ops.wgrad_args(x, dy, a) for x, dy, a in zip(xs, dys, accs)])
    for got, one, ref in zip(accs, single, want):
        torch.testing.assert_close(got, ref, atol=3e-2, rtol=3e-3)
        torch.testing.assert_close(got, one, atol=3e-2, rtol=3e-3)
    # bias gradients
    dbs = [torch.randn(o, device=dev()) for _, o in shapes]
    want_b = [b + dy.float().sum(0) for b, dy in zip(dbs, dys)]
    ops.colsum_grouped(list(zip(dys, dbs)))
    for got, ref in zip(dbs, want_b):
        torch.testing.assert_close(got, ref, atol=5e-2, rtol=3e-3)
    # non-qualifying group (a 64-wide problem): per-problem fallback, same numbers
    x2, dy2 = rnd(tokens, 64, seed=77), rnd(tokens, 72, seed=78)
    a2, a3 = torch.zeros(64, 72, device=dev()), torch.zeros(512, 1536, device=dev())
    ops.gemm_grouped([ops.wgrad_args(x2, dy2, a2), ops.wgrad_args(xs[0], dys[0], a3)])
    torch.testing.assert_close(a2, x2.float().t() @ dy2.float(), atol=3e-2, rtol=3e-3)
    torch.testing.assert_close(a3, xs[0].float().t() @ dys[0].float(), atol=3e-2, rtol=3e-3)


# ------------------------------------------------------------------------------------------------ add + LN
@pytest.mark.parametrize("rows,cols", [(37, 64), (4096, 512), (100, 1024), (5000, 512), (300, 128)])
def test_add_ln_fwd_bwd(rows, cols):
    from zero_b200 import ops
    x, y = rnd(rows, cols, seed=1), rnd(rows, cols, seed=2)
    scale = (1 + 0.1 * torch.randn(cols)).to(dev())
    offset = (0.1 * torch.randn(cols)).to(dev())
    out = torch.empty_like(x)
    mean = torch.empty(rows, device=dev())
    rstd = torch.empty(rows, device=dev())
    ops.add_ln_fwd(x, y, out, scale, offset, mean, rstd, 1e-8)
    xr = (x.float() + y.float()).requires_grad_(True)
    sc, of = scale.clone().requires_grad_(True), offset.clone().requires_grad_(True)
    mu = xr.mean(-1, keepdim=True)
    var = ((xr - mu) ** 2).mean(-1, keepdim=True)
    ref = sc * (xr - mu) * torch.rsqrt(var + 1e-8) + of
    torch.testing.assert_close(out.float(), ref.detach(), atol=2e-2, rtol=2e-2)
    d1, d2 = rnd(rows, cols, seed=3), rnd(rows, cols, seed=4)
    ref.backward(d1.float() + d2.float())
    ds = torch.empty_like(x)
    dscale = torch.zeros(cols, device=dev())
    doffset = torch.zeros(cols, device=dev())
    dbias = torch.zeros(cols, device=dev())
    ops.add_ln_bwd(x, y, d1, d2, mean, rstd, scale, ds, dscale, doffset, dbias)
    torch.testing.assert_close(ds.float(), xr.grad, atol=3e-2, rtol=2e-2)
    torch.testing.assert_close(dbias, xr.grad.sum(0), atol=2e-2 * math.sqrt(rows), rtol=1e-2)
    torch.testing.assert_close(dscale, sc.grad, atol=2e-2 * math.sqrt(rows), rtol=1e-2)
    torch.testing.assert_close(doffset, of.grad, atol=2e-2 * math.sqrt(rows), rtol=1e-2)
    # without the bias-gradient rider and with a single gradient addend
    ds2 = torch.empty_like(x)
    dscale2, doffset2 = torch.zeros(cols, device=dev()), torch.zeros(cols, device=dev())
    d12 = (d1.float() + d2.float()).to(torch.bfloat16)
    ops.add_ln_bwd(x, y, d12, None, mean, rstd, scale, ds2, dscale2, doffset2, None)
    torch.testing.assert_close(ds2.float(), xr.grad, atol=6e-2, rtol=3e-2)
    torch.testing.assert_close(doffset2, d12.float().sum(0), atol=2e-2 * math.sqrt(rows), rtol=1e-2)


# ------------------------------------------------------------------------------------------------ embedding
def test_embed_fwd_bwd():
    from oracle import zero_oracle as zo
    from zero_b200 import ops
    V, d, B, Lq = 96, 64, 3, 7
    table = rnd(V, d, seed=1)
    bias = (0.1 * torch.randn(d)).to(dev())
    ids = torch.randint(0, V, (B, Lq), dtype=torch.int32, device=dev())
    for shift in (0, 1):
        out = torch.empty(B, Lq, d, dtype=bf16, device=dev())
        ops.embed_fwd(ids, table, bias, out, mult=d ** 0.5, shift=shift)
        x = table.float()[ids.long()] * d ** 0.5 + bias
        if shift:
            x = torch.nn.functional.pad(x, (0, 0, 1, 0))[:, :-1]
        ref = x + zo.timing_signal(Lq, d, torch.float32).to(dev())
        torch.testing.assert_close(out.float(), ref, atol=3e-2, rtol=2e-2)
        d_out, d_out2 = rnd(B, Lq, d, seed=5), rnd(B, Lq, d, seed=6)
        d_table = torch.zeros(V, d, device=dev())
        d_bias = torch.zeros(d, device=dev())
        ops.embed_bwd(ids, d_out, d_table, d_bias, mult=d ** 0.5, shift=shift, d_out2=d_out2)
        g = (d_out.float() + d_out2.float())
        if shift:
            g = g[:, 1:]
            idx = ids[:, :-1]
        else:
            idx = ids
        want_t = torch.zeros(V, d, device=dev()).index_add_(0, idx.reshape(-1).long(), g.reshape(-1, d) * d ** 0.5)
        torch.testing.assert_close(d_table, want_t, atol=2e-2, rtol=1e-3)
        torch.testing.assert_close(d_bias, g.reshape(-1, d).sum(0), atol=2e-2, rtol=1e-3)
    # cached decode: every row at position `time`, zeroed when all ids are pad (models/transformer.py:113-117)
    ids0 = torch.zeros(4, 1, dtype=torch.int32, device=dev())
    out = torch.empty(4, 1, d, dtype=bf16, device=dev())
    ops.embed_fwd(ids0, table, bias, out, mult=d ** 0.5, zero_if_all_pad=True, time=5)
    ref = zo.timing_signal(1, d, torch.float32, time=5).to(dev()).expand(4, 1, d)
    torch.testing.assert_close(out.float(), ref, atol=1e-2, rtol=1e-2)


# ------------------------------------------------------------------------------------------------ CE
@pytest.mark.parametrize("V", [200, 1002, 32000, 40000])
def test_softmax_ce(V):
    from oracle import zero_oracle as zo
    from zero_b200 import ops
    B, T = 4, 9
    logits = (torch.randn(B * T, V) * 2).to(dev())
    labels = torch.randint(3, V, (B, T), dtype=torch.int32, device=dev())
    labels[1, 5:] = 0
    labels[2, 2:] = 0
    nll = torch.empty(B * T, device=dev())
    per = torch.empty(B, device=dev())
    loss = torch.empty(1, device=dev())
    dl = torch.empty(B * T, V, dtype=bf16, device=dev())
    ops.softmax_ce(logits, labels, nll, 0.1, d_logits=dl, per_sample=per, loss=loss)
    lg = logits.clone().requires_grad_(True)
    ce = zo.smoothed_ce(lg, labels.reshape(-1), 0.1).reshape(B, T)
    m = (labels != 0).float()
    ps = (ce * m).sum(-1) / m.sum(-1)
    ref_loss = ps.mean()
    ref_loss.backward()
    torch.testing.assert_close(nll, ce.detach().reshape(-1), atol=2e-4, rtol=1e-4)
    torch.testing.assert_close(per, ps.detach(), atol=2e-4, rtol=1e-4)
    torch.testing.assert_close(loss[0], ref_loss.detach(), atol=2e-4, rtol=1e-4)
    torch.testing.assert_close(dl.float(), lg.grad, atol=2e-4, rtol=1e-2)
    # score mode: no smoothing, no gradient
    ops.softmax_ce(logits, labels, nll, 0.0)
    ce0 = zo.smoothed_ce(logits, labels.reshape(-1), 0.0)
    torch.testing.assert_close(nll, ce0, atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("B,T,V,d,smooth", [(64, 64, 32000, 512, 0.1), (5, 9, 1000, 128, 0.1), (3, 7, 208, 64, 0.1),
                                            (11, 64, 40000, 512, 0.0), (2, 300, 1304, 256, 0.1), (4, 9, 128, 64, 0.1)])
def test_vocab_ce_fused_matches_gemm_plus_softmax_ce(B, T, V, d, smooth):
    """K6 (zb_vocab_ce: the vocabulary projection with the label-smoothed CE in the GEMM epilogue, logits never
    written) against the two-kernel path it replaces (zb_gemm into fp32 logits + zb_softmax_ce) and against the oracle
    (util.label_smooth + softmax_cross_entropy_with_logits_v2, models/transformer.py:186-211): per-token NLL, per-sentence
    and batch means, d_logits.  Covers the BASELINE configs[1] shape, ragged row / vocabulary tiles (45 rows, V = 1000,
    1304, 208 — the last tile partly or wholly past V), the minimum vocabulary, smoothing off (score_fn)."""
    from oracle import zero_oracle as zo
    from zero_b200 import ops
    import zero_b200.lib as L
    N = B * T
    feat = rnd(N, d, seed=81)
    table = rnd(V, d, scale=d ** -0.5 * 2, seed=82)
    labels = torch.randint(3, V, (B, T), dtype=torch.int32, device=dev())
    labels[1, T // 2:] = 0
    labels[-1, 2:] = 0
    labels[0, 0] = V - 1
    pitch = (V + 7) // 8 * 8
    outs = []
    for fused in (False, True):
        nll, per, loss = torch.empty(N, device=dev()), torch.empty(B, device=dev()), torch.empty(1, device=dev())
        dl = torch.zeros(N, pitch, dtype=bf16, device=dev())[:, :V]
        if fused:
            assert ops.vocab_ce_supported(N, d, V, feat, table, dl)
            ops.vocab_ce(feat, table, labels, nll, smooth, lambda nbytes: torch.empty(nbytes, dtype=torch.uint8, device=dev()),
                         d_logits=dl, per_sample=per, loss=loss, loss_scale=2.0)
        else:
            logits = torch.empty(N, pitch, device=dev())[:, :V]
            ops.gemm(feat, table, logits, L.ZB_K_MAJOR, L.ZB_K_MAJOR)
            ops.softmax_ce(logits, labels, nll, smooth, d_logits=dl, per_sample=per, loss=loss, loss_scale=2.0)
        outs.append((nll, per, loss, dl))
    (nll0, per0, loss0, dl0), (nll1, per1, loss1, dl1) = outs
    torch.testing.assert_close(nll1, nll0, atol=2e-3, rtol=1e-4)
    torch.testing.assert_close(per1, per0, atol=2e-3, rtol=1e-4)
    torch.testing.assert_close(loss1, loss0, atol=2e-3, rtol=1e-4)
    scale = float(dl0.float().abs().max())
    assert float((dl1.float() - dl0.float()).abs().max()) <= 2e-2 * scale + 1e-7
    # ... and against the oracle on the same bf16-rounded operands (fp32 logits)
    lg = (feat.float() @ table.float().t()).requires_grad_(True)
    ce = zo.smoothed_ce(lg, labels.reshape(-1), smooth).reshape(B, T)
    m = (labels != 0).float()
    ps = (ce * m).sum(-1) / m.sum(-1)
    (ps.mean() * 2.0).backward()
    torch.testing.assert_close(nll1, ce.detach().reshape(-1), atol=2e-3, rtol=1e-4)
    torch.testing.assert_close(loss1[0], ps.mean().detach(), atol=2e-3, rtol=1e-4)
    assert float((dl1.float() - lg.grad).abs().max()) <= 2e-2 * float(lg.grad.abs().max()) + 1e-7
    # forward only (score_fn): no d_logits
    nll2 = torch.empty(N, device=dev())
    ops.vocab_ce(feat, table, labels, nll2, 0.0, lambda nbytes: torch.empty(nbytes, dtype=torch.uint8, device=dev()))
    torch.testing.assert_close(nll2, zo.smoothed_ce(lg.detach(), labels.reshape(-1), 0.0), atol=2e-3, rtol=1e-4)


@pytest.mark.parametrize("rows,k,n", [(256, 2048, 512), (256, 512, 512), (37, 1024, 512), (8, 256, 128)])
def test_decode_split_k_projection_into_add_ln(rows, k, n):
    """The decode step's out = LN(res + inp @ W + b) with the k dimension split over CTAs (fp32 partial tiles reduce-added
    into one accumulator, consumed and CLEARED by zb_add_ln_fwd's y32 path) against the unsplit projection + bf16 y."""
    from zero_b200 import ops
    import zero_b200.lib as L
    inp, w, b = rnd(rows, k, seed=91), rnd(k, n, scale=k ** -0.5, seed=92), rnd(n, seed=93).float()
    res = rnd(rows, n, seed=94)
    scale, offset = (1 + 0.1 * rnd(n, seed=95).float()), 0.1 * rnd(n, seed=96).float()
    y = torch.empty(rows, n, dtype=bf16, device=dev())
    ops.linear_fwd(inp, w, b, y)
    want = torch.empty(rows, n, dtype=bf16, device=dev())
    ops.add_ln_fwd(res, y, want, scale, offset, eps=1e-8)
    y32 = torch.zeros(rows, n, device=dev())
    for _ in range(2):            # twice: the second pass relies on the accumulator having been cleared
        got = torch.empty(rows, n, dtype=bf16, device=dev())
        ops.gemm(inp, w, y32, L.ZB_K_MAJOR, L.ZB_MN_MAJOR, accum=True, split_k=max(2, min(8, k // 128)))
        ops.add_ln_fwd(res, None, got, scale, offset, eps=1e-8, y32=y32, ybias=b)
        assert float(y32.abs().max()) == 0.0
        torch.testing.assert_close(got.float(), want.float(), atol=3e-2, rtol=2e-2)
        assert float((got.float() - want.float()).abs().mean()) < 2e-3


# ------------------------------------------------------------------------------------------------ attention
def _attn_ref(q, k, v, heads, key_len, causal, q_off, inf, ek, ev, max_rel, relu):
    from oracle import zero_oracle as zo
    B, Lq, D = q.shape
    Lk = k.shape[1]
    dh = D // heads
    qh, kh, vh = [zo.heads_split(t, heads) for t in (q, k, v)]
    qh = qh * dh ** -0.5
    logits = qh @ kh.transpose(-1, -2)
    if ek is not None:
        idx = zo.rel_index(Lq, Lk, max_rel, q_off).to(q.device)
        logits = logits + torch.einsum("bhid,ijd->bhij", qh, ek[idx])
    bias = torch.zeros(B, 1, Lq, Lk, device=q.device)
    if key_len is not None:
        pad = torch.arange(Lk, device=q.device)[None, :] >= key_len[:, None].long()
        bias = bias + pad[:, None, None, :].float() * -inf
    if causal:
        i = torch.arange(Lq, device=q.device)[:, None] + q_off
        j = torch.arange(Lk, device=q.device)[None, :]
        bias = bias + (j > i).float()[None, None] * -inf
    if relu:
        w = torch.relu(logits * (bias == 0).float())
    else:
        w = torch.softmax(logits + bias, -1)
    o = w @ vh
    if ev is not None:
        o = o + torch.einsum("bhij,ijd->bhid", w, ev[idx])
    return zo.heads_merge(o)


@pytest.mark.parametrize("cfg", [
    dict(B=3, h=2, Lq=11, Lk=11, dh=64, causal=False, klen=True),
    dict(B=2, h=4, Lq=9, Lk=9, dh=32, causal=True, klen=False),
    dict(B=2, h=4, Lq=40, Lk=70, dh=16, causal=False, klen=True),
    dict(B=2, h=2, Lq=33, Lk=33, dh=32, causal=True, klen=False, rpr=4),
    dict(B=2, h=2, Lq=20, Lk=45, dh=64, causal=False, klen=True, rpr=16),
    dict(B=2, h=2, Lq=17, Lk=17, dh=32, causal=True, klen=False, relu=True),
    dict(B=2, h=2, Lq=12, Lk=30, dh=64, causal=False, klen=True, relu=True),
    dict(B=4, h=8, Lq=64, Lk=64, dh=64, causal=False, klen=True),
    dict(B=4, h=8, Lq=64, Lk=64, dh=64, causal=True, klen=False),
    dict(B=2, h=8, Lq=128, Lk=128, dh=64, causal=True, klen=False),
    dict(B=3, h=2, Lq=40, Lk=70, dh=64, causal=False, klen=True),
    dict(B=2, h=4, Lq=100, Lk=100, dh=64, causal=True, klen=False),
    dict(B=2, h=2, Lq=17, Lk=200, dh=64, causal=False, klen=True),
])
def test_attention_fwd_bwd(cfg):
    from zero_b200 import ops
    B, h, Lq, Lk, dh = cfg["B"], cfg["h"], cfg["Lq"], cfg["Lk"], cfg["dh"]
    D = h * dh
    rpr, relu = cfg.get("rpr", 0), cfg.get("relu", False)
    q, k, v = rnd(B, Lq, D, seed=1), rnd(B, Lk, D, seed=2), rnd(B, Lk, D, seed=3)
    key_len = None
    if cfg["klen"]:
        key_len = torch.randint(1, Lk + 1, (B,), dtype=torch.int32, device=dev())
        key_len[0] = Lk
    ek = ev = None
    if rpr:
        ek, ev = rnd(2 * rpr + 1, dh, scale=0.5, seed=4), rnd(2 * rpr + 1, dh, scale=0.5, seed=5)
    o = torch.empty(B, Lq, D, dtype=bf16, device=dev())
    lse = torch.empty(B, h, Lq, device=dev())
    a = ops.attention_args(q, k, v, o, h, key_len=key_len, causal=cfg["causal"], lse=lse, rpr_k=ek, rpr_v=ev,
                           max_rel=rpr, relu_attn=relu)
    ops.attention_fwd(a)
    leaves = [t.float().requires_grad_(True) for t in (q, k, v)]
    ekf = ek.float().requires_grad_(True) if rpr else None
    evf = ev.float().requires_grad_(True) if rpr else None
    ref = _attn_ref(leaves[0], leaves[1], leaves[2], h, key_len, cfg["causal"], 0, 1e8, ekf, evf, rpr, relu)
    torch.testing.assert_close(o.float(), ref.detach(), atol=3e-2, rtol=3e-2)
    d_o = rnd(B, Lq, D, seed=6)
    ref.backward(d_o.float())
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    delta = torch.empty(B, h, Lq, device=dev())
    dek = torch.zeros(2 * rpr + 1, dh, device=dev()) if rpr else None
    dev_ = torch.zeros(2 * rpr + 1, dh, device=dev()) if rpr else None
    ops.attention_bwd(a, d_o, dq, dk, dv, delta, dek, dev_)
    tol = dict(atol=6e-2, rtol=5e-2)
    torch.testing.assert_close(dq.float(), leaves[0].grad, **tol)
    torch.testing.assert_close(dk.float(), leaves[1].grad, **tol)
    torch.testing.assert_close(dv.float(), leaves[2].grad, **tol)
    if rpr:
        torch.testing.assert_close(dek, ekf.grad, atol=0.15, rtol=5e-2)
        torch.testing.assert_close(dev_, evf.grad, atol=0.15, rtol=5e-2)


def test_attention_cached_decode_step():
    """lq = 1 against a growing key cache, kv_group sharing of per-sentence memory."""
    from zero_b200 import ops
    B, K, h, dh, S = 3, 4, 2, 32, 13
    D = h * dh
    q = rnd(B * K, 1, D, seed=1)
    mem_k, mem_v = rnd(B, S, D, seed=2), rnd(B, S, D, seed=3)
    key_len = torch.tensor([13, 7, 2], dtype=torch.int32, device=dev())
    o = torch.empty(B * K, 1, D, dtype=bf16, device=dev())
    a = ops.attention_args(q, mem_k, mem_v, o, h, key_len=key_len, kv_group=K)
    ops.attention_fwd(a)
    ref = _attn_ref(q.float(), mem_k.float().repeat_interleave(K, 0), mem_v.float().repeat_interleave(K, 0), h,
                    key_len.repeat_interleave(K), False, 0, 1e8, None, None, 0, False)
    torch.testing.assert_close(o.float(), ref, atol=3e-2, rtol=3e-2)


# ------------------------------------------------------------------------------------------------ optimizer / misc
@pytest.mark.parametrize("cfg", [
    dict(B=64, K=4, h=8, dh=64, S=64, relu=False),     # BASELINE configs[2] cross-attention shape
    dict(B=5, K=3, h=2, dh=16, S=70, relu=False),      # ragged key tile, odd group size
    dict(B=2, K=12, h=2, dh=32, S=33, relu=False),     # more rows per memory than warps in the CTA
    dict(B=3, K=1, h=4, dh=64, S=100, relu=True),      # ReLA decode (modules/rela.py:52-75)
])
def test_attention_decode_kernel(cfg):
    """The lq = 1 kernel (one warp per (row, head)): per-sentence memories shared by the beams, key-length masks,
    self-attention against a strided, partly filled cache with the causal offset, and the log-sum-exp output."""
    from zero_b200 import ops
    import zero_b200.lib as L
    B, K, h, dh, S = cfg["B"], cfg["K"], cfg["h"], cfg["dh"], cfg["S"]
    D = h * dh
    before = L.path_launch_count("attn_decode")
    q = rnd(B * K, 1, D, seed=1)
    mem = rnd(B, S, 2 * D, seed=2)                      # fused [k | v] memory like the engine's
    key_len = torch.randint(1, S + 1, (B,), dtype=torch.int32, device=dev())
    key_len[0] = S
    o = torch.empty(B * K, 1, D, dtype=bf16, device=dev())
    lse = torch.empty(B * K, h, 1, device=dev())
    a = ops.attention_args(q, mem[:, :, :D], mem[:, :, D:], o, h, key_len=key_len, kv_group=K, lse=lse,
                           relu_attn=cfg["relu"])
    ops.attention_fwd(a)
    kf, vf = mem[:, :, :D].float().repeat_interleave(K, 0), mem[:, :, D:].float().repeat_interleave(K, 0)
    ref = _attn_ref(q.float(), kf, vf, h, key_len.repeat_interleave(K), False, 0, 1e8, None, None, 0, cfg["relu"])
    torch.testing.assert_close(o.float(), ref, atol=3e-2, rtol=3e-2)
    if not cfg["relu"]:
        from oracle import zero_oracle as zo
        qh, kh = zo.heads_split(q.float(), h) * dh ** -0.5, zo.heads_split(kf, h)
        pad = torch.arange(S, device=dev())[None, :] >= key_len.repeat_interleave(K)[:, None].long()
        want = torch.logsumexp(qh @ kh.transpose(-1, -2) + pad[:, None, None, :].float() * -1e8, -1)
        torch.testing.assert_close(lse, want, atol=2e-2, rtol=2e-2)
    # self-attention step t against a [rows, cap, 3D] cache: keys 0..t valid, q_offset = t
    R, cap, t = B * K, S + 3, S - 1
    cache = rnd(R, cap, 3 * D, seed=5)
    o2 = torch.empty(R, 1, D, dtype=bf16, device=dev())
    a = ops.attention_args(cache[:, t:t + 1, :D], cache[:, :t + 1, D:2 * D], cache[:, :t + 1, 2 * D:], o2, h,
                           q_offset=t, causal=True, relu_attn=cfg["relu"])
    ops.attention_fwd(a)
    ref2 = _attn_ref(cache[:, t:t + 1, :D].float(), cache[:, :t + 1, D:2 * D].float(), cache[:, :t + 1, 2 * D:].float(),
                     h, None, True, t, 1e8, None, None, 0, cfg["relu"])
    torch.testing.assert_close(o2.float(), ref2, atol=3e-2, rtol=3e-2)
    assert L.path_launch_count("attn_decode") == before + 2


def test_adam_tf_and_sumsq_and_colsum():
    from oracle import zero_oracle as zo
    from zero_b200 import ops
    n = 10007
    p, g = torch.randn(n, device=dev()), torch.randn(n, device=dev())
    m, v = torch.zeros(n, device=dev()), torch.zeros(n, device=dev())
    pb = torch.empty(n, dtype=bf16, device=dev())
    p_ref, m_ref, v_ref = p.clone(), m.clone(), v.clone()
    for step in (1, 2, 3):
        lr_t = 0.01 * math.sqrt(1 - 0.98 ** step) / (1 - 0.9 ** step)
        ops.adam_tf(p, m, v, g, pb, 0.9, 0.98, 1e-8, lr_t, 0.25, torch.tensor([2.0], device=dev()))
        p_ref, m_ref, v_ref = zo.adam_tf_step(p_ref, m_ref, v_ref, g * 0.5, step, 0.01, 0.9, 0.98, 1e-8)
    torch.testing.assert_close(p, p_ref, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(pb.float(), p_ref, atol=1e-2, rtol=1e-2)
    out = torch.zeros(1, device=dev())
    ops.sumsq(g, out)
    torch.testing.assert_close(out[0], (g * g).sum(), rtol=1e-4, atol=1e-3)
    x = rnd(777, 130, seed=9)
    cs = torch.zeros(130, device=dev())
    ops.colsum(x, cs)
    torch.testing.assert_close(cs, x.float().sum(0), atol=5e-2, rtol=1e-3)
    for m_, n_ in ((1000, 512), (4096, 1536), (77, 2048)):   # 16-byte vectorised path, incl. a strided view
        xb = rnd(m_, n_ + 64, seed=10)[:, :n_]
        cs2 = torch.zeros(n_, device=dev())
        ops.colsum(xb, cs2)
        torch.testing.assert_close(cs2, xb.float().sum(0), atol=8e-2, rtol=1e-3)


def test_prefix_mean_and_gather_rows():
    from zero_b200 import ops
    x = rnd(3, 9, 64, seed=1)
    y = torch.empty_like(x)
    ops.prefix_mean_fwd(x, y)
    ref = torch.cumsum(x.float(), 1) / torch.arange(1, 10, device=dev())[None, :, None]
    torch.testing.assert_close(y.float(), ref, atol=2e-2, rtol=2e-2)
    dy = rnd(3, 9, 64, seed=2)
    dx = torch.empty_like(x)
    ops.prefix_mean_bwd(dy, dx)
    w = dy.float() / torch.arange(1, 10, device=dev())[None, :, None]
    ref_dx = torch.flip(torch.cumsum(torch.flip(w, [1]), 1), [1])
    torch.testing.assert_close(dx.float(), ref_dx, atol=2e-2, rtol=2e-2)
    src = rnd(8, 5, 64, seed=3)
    idx = torch.tensor([3, 3, 0, 7, 1, 1, 2, 6], dtype=torch.int32, device=dev())
    dst = torch.zeros_like(src)
    ops.gather_rows(src, idx, dst, row_elems=2 * 64)
    assert torch.equal(dst[:, :2], src[idx.long()][:, :2]) and float(dst[:, 2:].abs().sum()) == 0.0


# ------------------------------------------------------------------------------------------------ beam step
def test_beam_step_matches_oracle_step_for_step():
    """K8 on the oracle's own logits: indices must be bit-exact (search.py:141-228)."""
    from oracle import zero_oracle as zo
    from zero_b200.search import BeamState
    torch.manual_seed(0)
    B, K, V, steps = 5, 4, 208, 9

    class C:
        beam, alpha, decode_length, temperature, inf = K, 0.6, 6, 1.0, 1e8
    src = torch.randint(3, 50, (B, 7))
    src[1, 4:] = 0
    src[3, 2:] = 0
    logits_seq = [torch.randn(B * K, V) * 3 for _ in range(32)]
    # make EOS likely sometimes so the finished-merge path is exercised
    for t, lg in enumerate(logits_seq):
        lg[:, 2] += 2.0 if t % 3 == 2 else -1.0

    calls = {"n": 0}

    def enc_fn(source):
        return {"dummy": torch.zeros(source.shape[0], 1)}

    def dec_fn(tok, state, time):
        lg = logits_seq[calls["n"] - 1] if calls["n"] > 0 else logits_seq[0]
        calls["n"] += 1
        return lg, {"dummy": state["dummy"], "decoder": {"state": {}}}

    want = zo.beam_search(C, src, enc_fn, dec_fn)
    st = BeamState(B, K, V, src.to(dev()), C.decode_length, C.alpha, C.temperature, C.inf, dev())
    t = 0
    while st.not_finished(t):
        st.step(logits_seq[t].to(dev()), t)
        t += 1
    got = st.result()
    assert t == want["steps"]
    np.testing.assert_array_equal(got["seq"].cpu().numpy(), want["seq"].numpy())
    np.testing.assert_allclose(got["score"].cpu().numpy(), want["score"].numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,K,V", [(3, 1, 64), (4, 3, 207), (2, 5, 1000), (64, 4, 32000), (2, 8, 60001), (2, 2, 50000)])
def test_beam_row_kernel_equals_sentence_kernel(B, K, V, monkeypatch):
    """The row-parallel step kernel (one CTA per (sentence, beam) row, last-arriver merge) against the
    one-CTA-per-sentence kernel on the same logits, step for step: sequences, parents and flags bit-exact, scores
    to fp32 round-off (the two kernels sum the soft-max normaliser in different orders).  Covers vocabularies that
    are / are not a multiple of 4 (vector vs scalar staging), beyond the shared-memory staging limit, beams 1..8."""
    from zero_b200.search import BeamState
    g = torch.Generator().manual_seed(V + K)
    S = 6
    src = torch.randint(3, 50, (B, S), generator=g)
    src[0, 3:] = 0
    states = []
    monkeypatch.setenv("ZB_BEAM_PARTS", "0")      # the row kernel itself (the cluster kernel has its own test below)
    for rows in ("0", "1"):
        monkeypatch.setenv("ZB_BEAM_ROWS", rows)
        states.append(BeamState(B, K, V, src.to(dev()), 4, 0.6, 1.0 if V != 1000 else 0.7, 1e8, dev()))
    assert states[0].row_ws is None and states[1].row_ws is not None
    import zero_b200.lib as L
    n_rows, n_sent = L.path_launch_count("beam_rows"), L.path_launch_count("beam_sentence")
    t = 0
    while True:
        nf = [st.not_finished(t) for st in states]
        assert nf[0] == nf[1]
        if not nf[0]:
            break
        lg = (torch.randn(B * K, V, generator=g) * 3)
        lg[:, 2] += 3.0 if t % 3 == 2 else -1.0
        lg = lg.to(dev())
        for st in states:
            st.step(lg, t)
        for name in ("alive_seq", "fin_seq", "fin_flag", "parent"):
            assert torch.equal(getattr(states[0], name), getattr(states[1], name)), (name, t)
        for name in ("alive_logp", "alive_score", "fin_score"):
            torch.testing.assert_close(getattr(states[0], name), getattr(states[1], name), rtol=1e-5, atol=1e-5)
        t += 1
    assert t >= 3
    assert L.path_launch_count("beam_rows") == n_rows + t and L.path_launch_count("beam_sentence") == n_sent + t
    # the arrival tickets (last word of each sentence's scratch) are back to zero after every step
    assert int(states[1].row_ws.view(torch.int32).view(B, -1)[:, -1].abs().sum()) == 0


# Kernels behind a switch (validated on the GPU in round 2, gpurun_out/r02a_summary.txt -> profiles/): their parity
# tests flip the switch inside the test.
def unvalidated(fn):
    return fn


@unvalidated
@pytest.mark.parametrize("B,K,V", [(3, 1, 64), (4, 3, 207), (2, 5, 1000), (64, 4, 32000), (2, 8, 60000), (3, 2, 7)])
def test_beam_part_kernel_equals_sentence_kernel(B, K, V, monkeypatch):
    """ZB_BEAM_PARTS=1 (a 4-CTA cluster per row, DSMEM exchange of the soft-max statistics, threshold pass before
    the sorted lists) against the one-CTA-per-sentence kernel, step for step: sequences / parents / flags bit-exact,
    scores to fp32 round-off.  V = 7 leaves parts empty; 207 / 1000 exercise the scalar staging path."""
    import zero_b200.lib as L
    from zero_b200.search import BeamState
    g = torch.Generator().manual_seed(V + K)
    src = torch.randint(3, 50, (B, 6), generator=g)
    src[0, 3:] = 0
    monkeypatch.setenv("ZB_BEAM_ROWS", "0")
    ref = BeamState(B, K, V, src.to(dev()), 4, 0.6, 1.0 if V != 1000 else 0.7, 1e8, dev())
    monkeypatch.setenv("ZB_BEAM_ROWS", "1")
    new = BeamState(B, K, V, src.to(dev()), 4, 0.6, 1.0 if V != 1000 else 0.7, 1e8, dev())
    before = L.path_launch_count("beam_parts")
    t = 0
    while True:
        nf = [ref.not_finished(t), new.not_finished(t)]
        assert nf[0] == nf[1]
        if not nf[0]:
            break
        lg = torch.randn(B * K, V, generator=g) * 3
        lg[:, 2] += 3.0 if t % 3 == 2 else -1.0
        lg = lg.to(dev())
        ref.step(lg, t)
        monkeypatch.setenv("ZB_BEAM_PARTS", "1")
        new.step(lg, t)
        monkeypatch.setenv("ZB_BEAM_PARTS", "0")
        for name in ("alive_seq", "fin_seq", "fin_flag", "parent"):
            assert torch.equal(getattr(ref, name), getattr(new, name)), (name, t)
        for name in ("alive_logp", "alive_score", "fin_score"):
            torch.testing.assert_close(getattr(ref, name), getattr(new, name), rtol=1e-5, atol=1e-5)
        t += 1
    assert t >= 3 and L.path_launch_count("beam_parts") == before + t


def _topk_problem(R, V, d, seed):
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(R, d, generator=g).to(bf16).to(dev())
    pitch = (d + 7) // 8 * 8
    table = (torch.randn(V, pitch, generator=g) * 0.2).to(bf16).to(dev())[:, :d]
    logits = torch.zeros(R, (V + 7) // 8 * 8, dtype=f32, device=dev())[:, :V]
    from zero_b200 import ops
    ops.gemm(feat, table, logits, 0, 0)
    return feat, table, logits.contiguous()


@pytest.mark.parametrize("R,V,d,skip,temp", [(256, 32000, 512, -1, 1.0), (256, 32000, 512, 2, 1.0), (20, 1000, 128, 2, 0.7),
                                             (7, 256, 64, 5, 1.0), (33, 208, 128, 2, 1.0), (130, 1003, 72, -1, 1.0), (300, 4099, 512, 2, 1.3)])
def test_vocab_topk_candidates_hold_every_rows_top8_and_its_logsumexp(R, V, d, skip, temp):
    """zb_vocab_topk (K8 fused, the tcgen05 GEMM's ce_mode 3 epilogue) against zb_gemm's fp32 logits: per row, the union
    of the parts' candidates contains the row's 8 largest logits / T (skip column left out) with their columns in
    tf.nn.top_k's order, every candidate is the logit of its column, and the parts' statistics fold into the row's
    log-sum-exp.  Ragged rows (7, 20, 130, 300), a vocabulary that leaves the last part partly / wholly empty, K = 72."""
    from zero_b200 import ops
    feat, table, logits = _topk_problem(R, V, d, 5 + V)
    cands = ops.vocab_topk(feat, table, lambda n: torch.full(((n + 3) // 4,), float("nan"), dtype=f32, device=dev()),
                           skip_col=skip, temperature=temp)
    torch.cuda.synchronize()
    stats, vals, cols = cands.unpack()
    assert cands.parts == 2 * ((V + 255) // 256) and tuple(vals.shape) == (cands.parts, R, 8)
    x = logits / temp
    # statistics
    m, s = stats[:, :, 0], stats[:, :, 1]
    lse = torch.logsumexp(torch.where(m > -float("inf"), m + torch.log(s), m), 0)
    torch.testing.assert_close(lse, torch.logsumexp(x, 1), rtol=1e-5, atol=2e-5)
    part_max = torch.stack([x[:, p * 128:(p + 1) * 128].max(1).values if p * 128 < V else
                            torch.full((R,), -float("inf"), device=dev()) for p in range(cands.parts)])
    torch.testing.assert_close(m, part_max, rtol=0, atol=1e-5)
    # candidates: valid slots are real (column, logit) pairs of their own part, sorted, never the skip column
    valid = vals > -float("inf")
    pv, pc = vals.permute(1, 0, 2).reshape(R, -1), cols.permute(1, 0, 2).reshape(R, -1).long()
    pvalid = valid.permute(1, 0, 2).reshape(R, -1)
    assert bool((pc[pvalid] >= 0).all()) and bool((pc[pvalid] < V).all()) and not bool((pc[pvalid] == skip).any())
    got_x = torch.gather(x, 1, pc.clamp(0, V - 1))
    torch.testing.assert_close(pv[pvalid], got_x[pvalid], rtol=1e-5, atol=1e-5)
    part_of = torch.arange(cands.parts, device=dev()).repeat_interleave(8)[None, :].expand(R, -1)
    assert bool(((pc // 128) == part_of)[pvalid].all())
    assert bool((vals[:, :, :-1] >= vals[:, :, 1:]).all())
    n_valid = valid.sum(2)
    for p in range(cands.parts):
        width = max(0, min(V, (p + 1) * 128) - p * 128) - (1 if p * 128 <= skip < (p + 1) * 128 else 0)
        assert bool((n_valid[p] == min(8, width)).all()), p
    # the row's top-8 from the candidates == top-8 of the logits (skip column removed)
    xs = x.clone()
    if skip >= 0:
        xs[:, skip] = -float("inf")
    want_v, want_c = torch.topk(xs, 8, dim=1)
    key = torch.where(pvalid, pv, torch.full_like(pv, -float("inf")))
    order = torch.argsort(key, dim=1, descending=True, stable=True)[:, :8]     # parts are in column order: stable = lower col
    got_v, got_c = torch.gather(key, 1, order), torch.gather(pc, 1, order)
    torch.testing.assert_close(got_v, want_v, rtol=1e-5, atol=1e-5)
    gap_ok = (want_v[:, :-1] - want_v[:, 1:]).min(1).values > 1e-4            # rows without a near-tie in the top 8
    assert int(gap_ok.sum()) >= R * 3 // 4
    assert torch.equal(got_c[gap_ok], want_c[gap_ok])


@pytest.mark.parametrize("B,K,V,d", [(64, 4, 32000, 512), (5, 4, 1000, 128), (3, 2, 300, 64), (4, 1, 4099, 256), (6, 4, 208, 128)])
def test_beam_step_from_candidates_equals_the_logits_step(B, K, V, d, monkeypatch):
    """zb_beam_step fed by zb_vocab_topk's candidates (no logits in memory) against the same step fed by the logits of
    the same features, step for step over a whole search: sequences / parents / flags bit-exact, scores to fp32
    round-off (the two log-sum-exps add the same terms in a different order).  Temperature 0.7 for V = 1000."""
    import zero_b200.lib as L
    from zero_b200 import ops
    from zero_b200.search import BeamState
    g = torch.Generator().manual_seed(V + K)
    src = torch.randint(3, 50, (B, 6), generator=g)
    src[0, 3:] = 0
    temp = 0.7 if V == 1000 else 1.0
    ref = BeamState(B, K, V, src.to(dev()), 4, 0.6, temp, 1e8, dev())
    new = BeamState(B, K, V, src.to(dev()), 4, 0.6, temp, 1e8, dev())
    before = L.path_launch_count("beam_cand")
    ws = {}
    t = 0
    while True:
        nf = [ref.not_finished(t), new.not_finished(t)]
        assert nf[0] == nf[1]
        if not nf[0]:
            break
        feat, table, logits = _topk_problem(B * K, V, d, 100 * t + V)
        if t % 3 == 2:
            table[2] = feat[0] * 0.5            # EOS scores high for some rows: beams finish
            logits = torch.zeros(B * K, (V + 7) // 8 * 8, dtype=f32, device=dev())[:, :V]
            ops.gemm(feat, table, logits, 0, 0)
            logits = logits.contiguous()
        ref.step(logits, t)
        cands = ops.vocab_topk(feat, table, lambda n: ws.setdefault(n, torch.zeros((n + 3) // 4, dtype=f32, device=dev())),
                               **new.candidate_request(t))
        new.step(cands, t)
        for name in ("alive_seq", "fin_seq", "fin_flag", "parent"):
            assert torch.equal(getattr(ref, name), getattr(new, name)), (name, t)
        for name in ("alive_logp", "alive_score", "fin_score"):
            torch.testing.assert_close(getattr(ref, name), getattr(new, name), rtol=1e-5, atol=1e-5)
        t += 1
    assert t >= 3 and L.path_launch_count("beam_cand") == before + t


@unvalidated
@pytest.mark.parametrize("m,n,k", [(256, 512, 512), (256, 2048, 512), (256, 512, 2048), (256, 1024, 1024),
                                   (1, 64, 64), (37, 136, 72), (100, 1000, 520), (300, 2048, 512), (511, 384, 128)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (0, 0), (1, 1), (1, 0)])
def test_gemm_bm64_tiles(m, n, k, a_mn, b_mn, monkeypatch):
    """ZB_GEMM_BM64=1: the single-CTA tcgen05 kernel with 64-row tiles (accumulator in the lower 16 lanes of each TMEM
    quadrant) for problems of <= 512 rows: every operand layout, ragged m / n / k, fp32 output; then bias + relu into
    a strided bf16 destination and the relu-masked dgrad epilogue."""
    from zero_b200 import ops
    import zero_b200.lib as L
    monkeypatch.setenv("ZB_GEMM_BM64", "1")
    before = L.path_launch_count("gemm_bm64")
    A, B = rnd(m, k, seed=21), rnd(n, k, scale=0.1, seed=22)

    def pad(t):
        c = (t.shape[1] + 7) // 8 * 8
        buf = torch.zeros(t.shape[0], c, dtype=t.dtype, device=t.device)
        buf[:, :t.shape[1]] = t
        return buf[:, :t.shape[1]]
    a_st, b_st = pad(A.t().contiguous() if a_mn else A), pad(B.t().contiguous() if b_mn else B)
    out = pad(torch.zeros(m, n, dtype=f32, device=dev()))
    ops.gemm(a_st, b_st, out, a_mn, b_mn, m=m, n=n, k=k)
    ref = A.float() @ B.float().t()
    torch.testing.assert_close(out, ref, atol=2e-3, rtol=2e-3)
    assert L.path_launch_count("gemm_bm64") == before + 1
    if (a_mn, b_mn) == (0, 1) and n % 8 == 0:
        bias = torch.randn(n, device=dev())
        buf = torch.full((m, 3, n + 64), 7.0, dtype=bf16, device=dev())
        ops.linear_fwd(pad(A), pad(B.t().contiguous()), bias, buf[:, 1, :n], relu=True)
        torch.testing.assert_close(buf[:, 1, :n].float(), torch.relu(ref + bias), atol=3e-2, rtol=2e-2)
        assert float((buf[:, 0] - 7).abs().max()) == 0 and float((buf[:, 2] - 7).abs().max()) == 0
        assert float((buf[:, 1, n:] - 7).abs().max()) == 0
    if (a_mn, b_mn) == (0, 0) and n % 8 == 0:
        mask = rnd(m, n, seed=23)
        dx = torch.empty(m, n, dtype=bf16, device=dev())
        ops.linear_dgrad(pad(A), pad(B), dx, relu_mask=mask)   # dx = (A @ B^T) * (mask > 0), B stored [n][k]
        torch.testing.assert_close(dx.float(), ref * (mask.float() > 0), atol=3e-2, rtol=2e-2)


@unvalidated
@pytest.mark.parametrize("rows,d", [(256, 512), (5, 64), (33, 1024)])
def test_fused_small_decode_kernels_equal_the_unfused_pairs(rows, d):
    """zb_aan_cat_step == zb_aan_step + two zb_add2d copies (bit-exact); zb_aan_gate_ln == zb_aan_gate_fwd +
    zb_add_ln_fwd (same rounding points and reduction order; compared to 1 bf16 ulp)."""
    from zero_b200 import ops
    x, z = rnd(rows, d, seed=31), rnd(rows, 2 * d, seed=32)
    sums_a = torch.randn(rows, d, device=dev())
    sums_b = sums_a.clone()
    t = 5
    # unfused
    xf = torch.empty_like(x)
    cat = torch.zeros(rows, 2 * d, dtype=bf16, device=dev())
    ops.aan_step(x, sums_a, xf, t)
    ops.add2d(x, None, cat[:, :d])
    ops.add2d(xf, None, cat[:, d:])
    # fused
    xf2 = torch.empty_like(x)
    cat2 = torch.zeros(rows, 2 * d, dtype=bf16, device=dev())
    ops.aan_cat_step(x, sums_b, cat2, xf2, t)
    assert torch.equal(xf, xf2) and torch.equal(cat, cat2) and torch.equal(sums_a, sums_b)
    scale, offset = torch.randn(d, device=dev()), torch.randn(d, device=dev())
    y = torch.empty_like(x)
    out = torch.empty_like(x)
    ops.aan_gate_fwd(x, xf, z, y)
    ops.add_ln_fwd(x, y, out, scale, offset, eps=1e-8)
    out2 = torch.empty_like(x)
    ops.aan_gate_ln(x, xf, z, out2, scale, offset, 1e-8)
    torch.testing.assert_close(out2.float(), out.float(), atol=2e-2, rtol=1e-2)
    assert float((out2.float() - out.float()).abs().max()) <= 0.0625   # at most an ulp of bf16 at |v| < 8


@pytest.mark.parametrize("B,h,Lq,Lk,causal,klen,fused", [
    (4, 8, 64, 64, False, True, True),        # the 64-token training batches: two heads stacked per 128-row block
    (4, 8, 64, 64, True, False, True),
    (64, 8, 64, 64, False, True, True),       # BASELINE configs[1]: 256 units over the persistent grid
    (3, 2, 64, 64, True, True, False),
    (3, 4, 40, 33, False, True, False),       # ragged pair blocks (zero-filled rows / keys)
    (3, 3, 40, 64, False, True, False),       # odd head count: one head per block
    (2, 8, 128, 128, True, False, True),      # BASELINE configs[3] lengths: one 128 x 128 block per (batch, head)
    (2, 4, 128, 128, False, True, True),
    (2, 4, 100, 100, True, False, False),
    (2, 2, 17, 200, False, True, False),      # cross attention over two key blocks: fp32 dq reduction
    (2, 2, 300, 300, True, True, True),       # three blocks each way, causal block skipping
    (1, 8, 64, 1024, False, True, False),     # BASELINE configs[4] decoder cross attention
    (1, 2, 1024, 1024, False, True, True),    # BASELINE configs[4] encoder self attention: online softmax over 8 blocks
])
def test_attention_tcgen05(B, h, Lq, Lk, causal, klen, fused, monkeypatch):
    """attention_tc.cu (the default for dh = 64): tcgen05 forward / backward against the torch restatement of
    func.dot_attention and against the mma.sync kernels it replaced (ZB_ATTN_TC=0), reading q / k / v in place from a
    fused [tokens, 3d] buffer (fused=True, self attention) or from separate tensors."""
    from zero_b200 import ops
    import zero_b200.lib as L
    D = h * 64
    if fused:
        assert Lq == Lk
        qkv = rnd(B, Lq, 3 * D, seed=41)
        q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    else:
        q, k, v = rnd(B, Lq, D, seed=42), rnd(B, Lk, D, seed=43), rnd(B, Lk, D, seed=44)
    key_len = None
    if klen:
        key_len = torch.randint(1, Lk + 1, (B,), dtype=torch.int32, device=dev())
        key_len[0] = Lk
    outs, lses = [], []
    for tc in ("0", "1"):
        monkeypatch.setenv("ZB_ATTN_TC", tc)
        before = L.path_launch_count("attn_tc")
        o = torch.empty(B, Lq, D, dtype=bf16, device=dev())
        lse = torch.empty(B, h, Lq, device=dev())
        ops.attention_fwd(ops.attention_args(q, k, v, o, h, key_len=key_len, causal=causal, lse=lse))
        assert L.path_launch_count("attn_tc") == before + int(tc)
        outs.append(o)
        lses.append(lse)
    leaves = [t.float().detach().clone().requires_grad_(True) for t in (q, k, v)]
    ref = _attn_ref(leaves[0], leaves[1], leaves[2], h, key_len, causal, 0, 1e8, None, None, 0, False)
    torch.testing.assert_close(outs[1].float(), ref.detach(), atol=3e-2, rtol=3e-2)
    torch.testing.assert_close(outs[1].float(), outs[0].float(), atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(lses[1], lses[0], atol=1e-3, rtol=1e-3)
    d_o = rnd(B, Lq, D, seed=45)
    ref.backward(d_o.float())
    grads = []
    for tc in ("0", "1"):
        monkeypatch.setenv("ZB_ATTN_TC", tc)
        before = L.path_launch_count("attn_tc")
        if fused:
            dqkv = torch.zeros(B, Lq, 3 * D, dtype=bf16, device=dev())
            dq, dk, dv = dqkv[:, :, :D], dqkv[:, :, D:2 * D], dqkv[:, :, 2 * D:]
        else:
            dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
        a = ops.attention_args(q, k, v, outs[0], h, key_len=key_len, causal=causal, lse=lses[0])
        delta = torch.empty(B, h, Lq, device=dev())
        ops.attention_bwd(a, d_o, dq, dk, dv, delta, None, None,
                          workspace=lambda nbytes: torch.empty(nbytes, dtype=torch.uint8, device=dev()))
        assert L.path_launch_count("attn_tc") == before + int(tc)
        grads.append((dq, dk, dv))
    for got, old_, want in zip(grads[1], grads[0], [t.grad for t in leaves]):
        torch.testing.assert_close(got.float(), want, atol=6e-2, rtol=5e-2)
        torch.testing.assert_close(got.float(), old_.float(), atol=4e-2, rtol=4e-2)
    # without a workspace a multi-block backward is declined by the tcgen05 path and served by the mma.sync kernels
    if Lk > 128:
        monkeypatch.setenv("ZB_ATTN_TC", "1")
        before = L.path_launch_count("attn_tc")
        a = ops.attention_args(q, k, v, outs[0], h, key_len=key_len, causal=causal, lse=lses[0])
        assert ops.attention_bwd_workspace_bytes(a) == B * Lq * D * 4
        dq2, dk2, dv2 = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
        ops.attention_bwd(a, d_o, dq2, dk2, dv2, torch.empty(B, h, Lq, device=dev()), None, None)
        assert L.path_launch_count("attn_tc") == before
        torch.testing.assert_close(dq2.float(), grads[0][0].float(), atol=1e-6, rtol=0)


@pytest.mark.parametrize("B,h,Lq,Lk,causal,klen,R,drop", [
    (2, 8, 128, 128, False, True, 16, 0.0),    # BASELINE configs[3]: encoder self attention
    (2, 8, 128, 128, True, False, 16, 0.0),    # ... decoder self attention
    (3, 2, 100, 128, False, True, 16, 0.0),    # ... cross attention with ragged lengths
    (2, 2, 40, 36, False, True, 6, 0.0),       # the len-40 golden model's shapes
    (2, 4, 20, 45, False, True, 16, 0.0),
    (2, 2, 128, 128, True, False, 1, 0.0),     # three buckets: almost everything is clipped
    (2, 4, 128, 128, True, True, 16, 0.25),    # with attention dropout (same keep mask as the generic kernels)
])
def test_attention_tcgen05_relative_positions(B, h, Lq, Lk, causal, klen, R, drop, monkeypatch):
    """modules/rpr.py:10-75 on the tensor cores (attention_tc.cu: bucket GEMMs Q E_k^T / W E_v and their gradients)
    against the torch restatement (einsum over the gathered [L, L, dh] tensor, as the reference does) and against
    the generic CUDA-core kernels (ZB_ATTN_TC=0), including d_rpr_keys / d_rpr_values."""
    from zero_b200 import ops
    import zero_b200.lib as L
    D = h * 64
    q, k, v = rnd(B, Lq, D, seed=61), rnd(B, Lk, D, seed=62), rnd(B, Lk, D, seed=63)
    ek, ev = rnd(2 * R + 1, 64, scale=0.5, seed=64), rnd(2 * R + 1, 64, scale=0.5, seed=65)
    d_o = rnd(B, Lq, D, seed=66)
    key_len = None
    if klen:
        key_len = torch.randint(1, Lk + 1, (B,), dtype=torch.int32, device=dev())
        key_len[0] = Lk
    seed = torch.tensor([4242], dtype=torch.int64, device=dev())
    res = []
    for tc in ("0", "1"):
        monkeypatch.setenv("ZB_ATTN_TC", tc)
        before = L.path_launch_count("attn_tc")
        o = torch.empty(B, Lq, D, dtype=bf16, device=dev())
        lse = torch.empty(B, h, Lq, device=dev())
        a = ops.attention_args(q, k, v, o, h, key_len=key_len, causal=causal, lse=lse, rpr_k=ek, rpr_v=ev, max_rel=R,
                               dropout=(drop, 9, seed) if drop else None)
        ops.attention_fwd(a)
        dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
        dek, dev_ = torch.zeros(2 * R + 1, 64, device=dev()), torch.zeros(2 * R + 1, 64, device=dev())
        ops.attention_bwd(a, d_o, dq, dk, dv, torch.empty(B, h, Lq, device=dev()), dek, dev_)
        assert L.path_launch_count("attn_tc") == before + 2 * int(tc)
        res.append((o, lse, dq, dk, dv, dek, dev_))
    for x0, x1 in zip(res[0][:5], res[1][:5]):
        torch.testing.assert_close(x1.float(), x0.float(), atol=4e-2, rtol=4e-2)
    for x0, x1 in zip(res[0][5:], res[1][5:]):
        torch.testing.assert_close(x1, x0, atol=0.15, rtol=5e-2)
    if drop:
        return
    leaves = [t.float().detach().clone().requires_grad_(True) for t in (q, k, v)]
    ekf, evf = ek.float().requires_grad_(True), ev.float().requires_grad_(True)
    ref = _attn_ref(leaves[0], leaves[1], leaves[2], h, key_len, causal, 0, 1e8, ekf, evf, R, False)
    o, lse, dq, dk, dv, dek, dev_ = res[1]
    torch.testing.assert_close(o.float(), ref.detach(), atol=3e-2, rtol=3e-2)
    ref.backward(d_o.float())
    for got, want in zip((dq, dk, dv), [t.grad for t in leaves]):
        torch.testing.assert_close(got.float(), want, atol=6e-2, rtol=5e-2)
    torch.testing.assert_close(dek, ekf.grad, atol=0.15, rtol=5e-2)
    torch.testing.assert_close(dev_, evf.grad, atol=0.15, rtol=5e-2)


@pytest.mark.parametrize("B,h,Lq,Lk,causal,klen,drop", [
    (4, 8, 64, 64, False, True, 0.0), (4, 8, 64, 64, True, False, 0.0), (2, 4, 128, 128, True, True, 0.0),
    (2, 2, 40, 36, False, True, 0.0), (1, 2, 300, 300, True, True, 0.0), (2, 4, 64, 200, False, True, 0.0),
    (2, 4, 64, 64, True, True, 0.2),
])
def test_attention_tcgen05_rela(B, h, Lq, Lk, causal, klen, drop, monkeypatch):
    """ReLA (modules/rela.py:52-75: relu(logits * keep) instead of softmax, no normaliser) on the tcgen05 kernels against
    the torch restatement and against the generic CUDA-core kernels (ZB_ATTN_TC=0)."""
    from zero_b200 import ops
    import zero_b200.lib as L
    D = h * 64
    q, k, v = rnd(B, Lq, D, seed=71), rnd(B, Lk, D, seed=72), rnd(B, Lk, D, seed=73)
    d_o = rnd(B, Lq, D, seed=74)
    key_len = None
    if klen:
        key_len = torch.randint(1, Lk + 1, (B,), dtype=torch.int32, device=dev())
        key_len[0] = Lk
    seed = torch.tensor([777], dtype=torch.int64, device=dev())
    res = []
    for tc in ("0", "1"):
        monkeypatch.setenv("ZB_ATTN_TC", tc)
        before = L.path_launch_count("attn_tc")
        o = torch.empty(B, Lq, D, dtype=bf16, device=dev())
        lse = torch.empty(B, h, Lq, device=dev())
        a = ops.attention_args(q, k, v, o, h, key_len=key_len, causal=causal, lse=lse, relu_attn=True,
                               dropout=(drop, 5, seed) if drop else None)
        ops.attention_fwd(a)
        dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
        ops.attention_bwd(a, d_o, dq, dk, dv, torch.empty(B, h, Lq, device=dev()), None, None,
                          workspace=lambda nbytes: torch.empty(nbytes, dtype=torch.uint8, device=dev()))
        assert L.path_launch_count("attn_tc") == before + 2 * int(tc)
        res.append((o, dq, dk, dv))
    # the unnormalised weights make large sums: compare relative to each tensor's scale
    for x0, x1 in zip(res[0], res[1]):
        scale = float(x0.float().abs().max()) + 1e-6
        assert float((x1.float() - x0.float()).abs().max()) <= 4e-2 * scale
    if drop:
        return
    leaves = [t.float().detach().clone().requires_grad_(True) for t in (q, k, v)]
    ref = _attn_ref(leaves[0], leaves[1], leaves[2], h, key_len, causal, 0, 1e8, None, None, 0, True)
    ref.backward(d_o.float())
    for got, want in zip(res[1], [ref.detach()] + [t.grad for t in leaves]):
        scale = float(want.abs().max()) + 1e-6
        assert float((got.float() - want).abs().max()) <= 5e-2 * scale


@pytest.mark.parametrize("B,h,Lq,Lk,causal", [(4, 8, 64, 64, True), (2, 4, 128, 128, False), (2, 2, 200, 200, True)])
def test_attention_tcgen05_dropout_matches_the_mma_kernels(B, h, Lq, Lk, causal, monkeypatch):
    """Attention dropout (func.py:245): the keep mask is a pure function of (seed, site, [b, h, i, j]) shared by every
    attention kernel, so the tcgen05 path must reproduce the mma.sync path's outputs and gradients to round-off."""
    from zero_b200 import ops
    D = h * 64
    q, k, v, d_o = rnd(B, Lq, D, seed=51), rnd(B, Lk, D, seed=52), rnd(B, Lk, D, seed=53), rnd(B, Lq, D, seed=54)
    seed = torch.tensor([12345], dtype=torch.int64, device=dev())
    res = []
    for tc in ("0", "1"):
        monkeypatch.setenv("ZB_ATTN_TC", tc)
        o = torch.empty(B, Lq, D, dtype=bf16, device=dev())
        lse = torch.empty(B, h, Lq, device=dev())
        a = ops.attention_args(q, k, v, o, h, causal=causal, lse=lse, dropout=(0.3, 77, seed))
        ops.attention_fwd(a)
        dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
        ops.attention_bwd(a, d_o, dq, dk, dv, torch.empty(B, h, Lq, device=dev()), None, None,
                          workspace=lambda nbytes: torch.empty(nbytes, dtype=torch.uint8, device=dev()))
        res.append((o, lse, dq, dk, dv))
    assert float((res[0][0].float() - rnd(B, Lq, D, seed=51).float()).abs().max()) > 0   # sanity: something was computed
    for x0, x1 in zip(res[0], res[1]):
        torch.testing.assert_close(x1.float(), x0.float(), atol=4e-2, rtol=4e-2)
    # ... and the mask really drops: the no-dropout output differs
    monkeypatch.setenv("ZB_ATTN_TC", "1")
    o2 = torch.empty(B, Lq, D, dtype=bf16, device=dev())
    ops.attention_fwd(ops.attention_args(q, k, v, o2, h, causal=causal, lse=torch.empty(B, h, Lq, device=dev())))
    assert float((o2.float() - res[1][0].float()).abs().max()) > 0.05


@unvalidated
def test_gumbel_add_noise_statistics_and_reproducibility():
    """zb_gumbel_add (util.gumbel_noise, utils/util.py:189-195): Gumbel(0, 1) has mean 0.5772 (Euler's constant) and
    variance pi^2 / 6; draws are a pure function of (seed, site, index)."""
    from zero_b200 import ops
    n = 1 << 20
    seed = torch.tensor([77], dtype=torch.int64, device=dev())
    x = torch.zeros(n + 1, device=dev())[:n]                      # odd tail handled too
    base = torch.full((n,), 3.0, device=dev())
    a = ops.gumbel_add(base.clone(), seed, 5) - 3.0
    assert abs(float(a.mean()) - 0.5772) < 8e-3 and abs(float(a.var()) - math.pi ** 2 / 6) < 2.5e-2
    assert float(a.max()) < 18.5 and float(a.min()) > -3.0        # -log(eps) = 18.4 caps the right tail
    b = ops.gumbel_add(base.clone(), seed, 5) - 3.0
    assert torch.equal(a, b)
    c = ops.gumbel_add(base.clone(), seed, 6) - 3.0
    seed.add_(1)
    d = ops.gumbel_add(base.clone(), seed, 5) - 3.0
    assert not torch.equal(a, c) and not torch.equal(a, d)
    assert abs(float((a * c).mean()) - 0.5772 ** 2) < 1.2e-2      # different sites are uncorrelated (6 sigma)
    odd = torch.zeros(1001, device=dev())
    ops.gumbel_add(odd, seed, 1)
    assert bool(torch.isfinite(odd).all()) and float(odd.abs().sum()) > 0
    del x
