"""Dropout on the CUDA path (tf.nn.dropout through util.valid_apply_dropout, utils/util.py:75-79): the kernels
regenerate a counter-based keep mask from (seed, site, flat index); the oracle is handed the SAME masks
(zero_oracle.dropout_keep restates the hash) so losses and gradients are compared like in the dropout-free tests."""
import numpy as np
import pytest
import torch

from tests.golden_util import load_golden

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def dev():
    return torch.device("cuda")


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(bf16).to(dev())


@pytest.mark.parametrize("n,rate", [(8 * 1000, 0.1), (12345, 0.3), (7, 0.5), (4096 * 512, 0.1)])
def test_dropout_kernel_matches_restated_mask(n, rate):
    from oracle import zero_oracle as zo
    from zero_b200 import ops
    from zero_b200.engine import dropout_site
    x, x2 = rnd(n, seed=1), rnd(n, seed=2)
    seed = torch.tensor([987654321012345], dtype=torch.int64, device=dev())
    out = torch.empty_like(x)
    ops.dropout(x, out, rate, seed, dropout_site("enc3.ffn.relu"))
    keep = torch.from_numpy(zo.dropout_keep(987654321012345, "enc3.ffn.relu", n, rate)).to(dev())
    want = (x.float() * keep / (1.0 - rate)).to(bf16)
    assert torch.equal(out, want)
    frac = float(keep.float().mean())
    assert abs(frac - (1.0 - rate)) < max(4.0 * (rate * (1 - rate) / n) ** 0.5, 1e-3) or n < 100
    # second addend + in place
    y = x.clone()
    ops.dropout(y, y, rate, seed, dropout_site("enc3.ffn.relu"), x2=x2)
    want2 = ((x.float() + x2.float()) * keep / (1.0 - rate)).to(bf16)
    assert torch.equal(y, want2)
    # another site / seed gives another mask
    other = torch.from_numpy(zo.dropout_keep(987654321012346, "enc3.ffn.relu", n, rate))
    if n > 100:
        assert (other != keep.cpu()).float().mean() > 0.5 * rate


def _attn_ref_drop(q, k, v, heads, key_len, causal, keep, rate, ek=None, ev=None, max_rel=0, relu=False):
    from oracle import zero_oracle as zo
    B, Lq, D = q.shape
    Lk = k.shape[1]
    dh = D // heads
    qh, kh, vh = [zo.heads_split(t, heads) for t in (q, k, v)]
    qh = qh * dh ** -0.5
    logits = qh @ kh.transpose(-1, -2)
    if ek is not None:
        idx = zo.rel_index(Lq, Lk, max_rel, 0).to(q.device)
        logits = logits + torch.einsum("bhid,ijd->bhij", qh, ek[idx])
    bias = torch.zeros(B, 1, Lq, Lk, device=q.device)
    if key_len is not None:
        pad = torch.arange(Lk, device=q.device)[None, :] >= key_len[:, None].long()
        bias = bias + pad[:, None, None, :].float() * -1e8
    if causal:
        i = torch.arange(Lq, device=q.device)[:, None]
        j = torch.arange(Lk, device=q.device)[None, :]
        bias = bias + (j > i).float()[None, None] * -1e8
    w = torch.relu(logits * (bias == 0).float()) if relu else torch.softmax(logits + bias, -1)
    w = w * keep / (1.0 - rate)
    o = w @ vh
    if ev is not None:
        o = o + torch.einsum("bhij,ijd->bhid", w, ev[idx])
    return zo.heads_merge(o)


@pytest.mark.parametrize("cfg", [
    dict(B=4, h=8, Lq=64, Lk=64, dh=64, causal=False, klen=True),      # single-tile tensor-core path
    dict(B=2, h=4, Lq=50, Lk=64, dh=64, causal=True, klen=False),      # single tile, ragged rows
    dict(B=2, h=4, Lq=100, Lk=130, dh=64, causal=False, klen=True),    # multi-tile tensor-core path
    dict(B=2, h=2, Lq=70, Lk=70, dh=64, causal=True, klen=False),
    dict(B=2, h=2, Lq=33, Lk=45, dh=32, causal=False, klen=True),      # CUDA-core path (dh 32)
    dict(B=2, h=2, Lq=20, Lk=45, dh=64, causal=False, klen=True, rpr=8),
    dict(B=2, h=2, Lq=17, Lk=17, dh=32, causal=True, klen=False, relu=True),
])
def test_attention_dropout_fwd_bwd(cfg):
    from oracle import zero_oracle as zo
    from zero_b200 import ops
    from zero_b200.engine import dropout_site
    B, h, Lq, Lk, dh = cfg["B"], cfg["h"], cfg["Lq"], cfg["Lk"], cfg["dh"]
    D, rate, rpr, relu = h * dh, 0.2, cfg.get("rpr", 0), cfg.get("relu", False)
    q, k, v = rnd(B, Lq, D, seed=1), rnd(B, Lk, D, seed=2), rnd(B, Lk, D, seed=3)
    key_len = None
    if cfg["klen"]:
        key_len = torch.randint(1, Lk + 1, (B,), dtype=torch.int32, device=dev())
        key_len[0] = Lk
    ek = ev = None
    if rpr:
        ek, ev = rnd(2 * rpr + 1, dh, scale=0.5, seed=4), rnd(2 * rpr + 1, dh, scale=0.5, seed=5)
    seed_val = 1234567890123
    seed = torch.tensor([seed_val], dtype=torch.int64, device=dev())
    o = torch.empty(B, Lq, D, dtype=bf16, device=dev())
    lse = torch.empty(B, h, Lq, device=dev())
    a = ops.attention_args(q, k, v, o, h, key_len=key_len, causal=cfg["causal"], lse=lse, rpr_k=ek, rpr_v=ev,
                           max_rel=rpr, relu_attn=relu, dropout=(rate, dropout_site("dec1.cross.att"), seed))
    ops.attention_fwd(a)
    keep = torch.from_numpy(zo.dropout_keep(seed_val, "dec1.cross.att", B * h * Lq * Lk, rate)).reshape(B, h, Lq, Lk)
    keep = keep.float().to(dev())
    leaves = [t.float().requires_grad_(True) for t in (q, k, v)]
    ekf = ek.float().requires_grad_(True) if rpr else None
    evf = ev.float().requires_grad_(True) if rpr else None
    ref = _attn_ref_drop(leaves[0], leaves[1], leaves[2], h, key_len, cfg["causal"], keep, rate, ekf, evf, rpr, relu)
    torch.testing.assert_close(o.float(), ref.detach(), atol=4e-2, rtol=3e-2)
    # without the mask the result must differ (the dropout path is really taken)
    assert float((o.float() - _attn_ref_drop(leaves[0], leaves[1], leaves[2], h, key_len, cfg["causal"],
                                             torch.ones_like(keep), 0.0, ekf, evf, rpr, relu).detach()).abs().max()) > 0.05
    d_o = rnd(B, Lq, D, seed=6)
    ref.backward(d_o.float())
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    delta = torch.empty(B, h, Lq, device=dev())
    dek = torch.zeros(2 * rpr + 1, dh, device=dev()) if rpr else None
    dev_ = torch.zeros(2 * rpr + 1, dh, device=dev()) if rpr else None
    ops.attention_bwd(a, d_o, dq, dk, dv, delta, dek, dev_)
    tol = dict(atol=8e-2, rtol=5e-2)
    torch.testing.assert_close(dq.float(), leaves[0].grad, **tol)
    torch.testing.assert_close(dk.float(), leaves[1].grad, **tol)
    torch.testing.assert_close(dv.float(), leaves[2].grad, **tol)
    if rpr:
        torch.testing.assert_close(dek, ekf.grad, atol=0.2, rtol=5e-2)
        torch.testing.assert_close(dev_, evf.grad, atol=0.2, rtol=5e-2)


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("name", ["transformer", "transformer_rpr", "transformer_rela", "transformer_aan_cumsum",
                                  "transformer_fuse"])
def test_train_step_with_all_dropouts_matches_oracle_with_same_masks(name):
    from oracle import zero_oracle as zo
    from zero_b200.engine import Engine
    z, hp, variables, grads, vs, vt = load_golden(name)
    rates = dict(dropout=0.1, attention_dropout=0.15, relu_dropout=0.2, residual_dropout=0.1)
    hp.override_from_dict(rates)
    eng = Engine(hp, vs, vt)
    eng.ps.load_state_dict(variables)
    seed_val = 4242424242
    eng.set_dropout_seed(seed_val)
    src, tgt = torch.from_numpy(z["source"]), torch.from_numpy(z["target"])
    loss = eng.forward_backward(src, tgt).clone()   # the loss lives in a workspace buffer that score() reuses
    got = eng.ps.grad_dict()
    torch.cuda.synchronize()
    # dropout is closed outside the training step (util.closing_dropout): score == golden score
    np.testing.assert_allclose(eng.score(src, tgt).cpu().numpy(), z["score"], atol=3e-2, rtol=1e-2)
    c = zo.Cfg(hp, vs, vt)
    P = {k: v.clone().requires_grad_(True) for k, v in variables.items()}
    zo.DROP = zo.make_drop(seed_val, {"emb": 0.1, "att": 0.15, "relu": 0.2, "res": 0.1})
    try:
        ref_loss, _, _, _ = zo.train_loss(c, P, src.long(), tgt.long())
        ref_loss.backward()
    finally:
        zo.DROP = None
    assert abs(float(ref_loss) - float(z["loss"])) > 1e-3          # the masks changed the loss ...
    assert abs(float(loss[0]) - float(ref_loss)) < 3e-2, (float(loss[0]), float(ref_loss))   # ... the same way
    worst = ("", 0.0)
    for k, p in P.items():
        g = p.grad
        if g is None or float(g.abs().max()) < 1e-6:
            continue
        r = _rel(got[k], g)
        worst = (k, r) if r > worst[1] else worst
        cos = torch.nn.functional.cosine_similarity(got[k].double().flatten(), g.double().flatten(), dim=0)
        assert cos > (0.93 if name == "transformer_rela" else 0.97), "%s: cosine %.4f rel %.4f" % (k, float(cos), r)
    assert worst[1] < (0.4 if name == "transformer_rela" else 0.15), "worst gradient %s rel err %.4f" % worst


def test_new_mask_every_step_and_graph_replay():
    """The seed is device-resident: CUDA-graph replays of the step must draw a new mask each time."""
    from zero_b200.engine import Engine
    from zero_b200.train import Trainer
    z, hp, variables, grads, vs, vt = load_golden("transformer")
    hp.override_from_dict(dict(dropout=0.1, attention_dropout=0.1, relu_dropout=0.1, residual_dropout=0.1,
                               lrate=1e-9, lrate_strategy="vanilla", beta1=0.9, beta2=0.98, epsilon=1e-8,
                               clip_grad_norm=0.0))
    eng = Engine(hp, vs, vt)
    eng.ps.load_state_dict(variables)
    tr = Trainer(eng, hp, world_size=1, use_graph=True)
    src, tgt = torch.from_numpy(z["source"]).cuda(), torch.from_numpy(z["target"]).cuda()
    losses = [float(tr.step(src, tgt)[0]) for _ in range(4)]
    assert len(set(round(l, 4) for l in losses)) == 4, losses      # lr ~ 0: only the masks differ
    assert all(abs(l - float(z["loss"])) < 0.5 for l in losses)
