"""Pin oracle/zero_oracle.py to vectors produced by the reference's own code (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import zero_oracle as zo
from tests.golden_util import MODELS, load_golden


@pytest.mark.parametrize("name", MODELS)
def test_train_loss_grads_logits(name):
    z, hp, variables, grads, vs, vt = load_golden(name)
    c = zo.Cfg(hp, vs, vt)
    assert set(zo.param_shapes(c).keys()) == set(variables.keys())
    for k, shp in zo.param_shapes(c).items():
        assert tuple(variables[k].shape) == tuple(shp), k
    P = {k: v.clone().requires_grad_(True) for k, v in variables.items()}
    src, tgt = torch.from_numpy(z["source"]), torch.from_numpy(z["target"])
    loss, logits, per_sample, enc = zo.train_loss(c, P, src, tgt)
    assert abs(float(loss.detach()) - float(z["loss"])) < 2e-5
    np.testing.assert_allclose(enc["encodes"].detach().numpy(), z["encodes"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(logits.detach().numpy(), z["logits"], atol=1e-4, rtol=1e-5)
    np.testing.assert_allclose(per_sample.detach().numpy(), z["per_sample_loss"], atol=2e-5, rtol=1e-5)
    names = sorted(P)
    gs = torch.autograd.grad(loss, [P[n] for n in names], allow_unused=True)
    for n, g in zip(names, gs):
        g = torch.zeros_like(P[n]) if g is None else g
        np.testing.assert_allclose(g.numpy(), grads[n].numpy(), atol=2e-6, rtol=2e-4, err_msg=n)


@pytest.mark.parametrize("name", MODELS)
def test_score_and_beam_search(name):
    z, hp, variables, grads, vs, vt = load_golden(name)
    c = zo.Cfg(hp, vs, vt)
    src, tgt = torch.from_numpy(z["source"]), torch.from_numpy(z["target"])
    with torch.no_grad():
        sc = zo.score(c, variables, src, tgt)
        np.testing.assert_allclose(sc.numpy(), z["score"], atol=2e-5, rtol=1e-5)
        enc_fn, dec_fn = zo.make_infer_fns(c, variables)
        steps = {}
        out = zo.beam_search(c, src, enc_fn, dec_fn, logits_hook=lambda t, lg: steps.setdefault(t, lg.clone()))
    # golden step_logits_0 is the cache_init dummy call; _1 is t = 0, _2 is t = 1 (search.py:56-77,141)
    for t in (0, 1, 2):
        np.testing.assert_allclose(steps[t].numpy(), z["step_logits_%d" % (t + 1)], atol=1e-4, rtol=1e-5)
    assert out["steps"] + 1 == int(z["n_decode_calls"])
    np.testing.assert_array_equal(out["seq"].numpy(), z["beam_seq"])  # bit-exact indices
    np.testing.assert_allclose(out["score"].numpy(), z["beam_score"], atol=1e-5, rtol=1e-5)


def test_adam_tf_matches_closed_form():
    p, m, v, g = [torch.tensor([x]) for x in (1.0, 0.0, 0.0, 0.5)]
    p1, m1, v1 = zo.adam_tf_step(p, m, v, g, 1, 0.1, 0.9, 0.98, 1e-8)
    # first step: m = 0.05, v = 0.005, lr_t = 0.1*sqrt(0.02)/0.1
    want = 1.0 - 0.1 * (0.02 ** 0.5) / 0.1 * 0.05 / (0.005 ** 0.5 + 1e-8)
    assert abs(float(p1) - want) < 1e-6


def test_dropout_hook_and_restated_keep_mask():
    """DROP = None is the dropout-free model; with masks the loss moves; the restated keep function is a pure
    function of (seed, site, index) with the requested rate."""
    import numpy as np
    keep = zo.dropout_keep(123, "enc0.ffn.relu", 200000, 0.25)
    assert abs(float(keep.mean()) - 0.75) < 5e-3
    assert (keep == zo.dropout_keep(123, "enc0.ffn.relu", 200000, 0.25)).all()
    assert (keep[:1000] == zo.dropout_keep(123, "enc0.ffn.relu", 1000, 0.25)).all()
    assert (keep != zo.dropout_keep(124, "enc0.ffn.relu", 200000, 0.25)).mean() > 0.2
    assert (keep != zo.dropout_keep(123, "enc1.ffn.relu", 200000, 0.25)).mean() > 0.2
    assert zo.dropout_keep(5, "x.res", 64, 0.0).all()
    z, hp, variables, grads, vs, vt = load_golden("transformer")
    c = zo.Cfg(hp, vs, vt)
    src, tgt = torch.from_numpy(z["source"]).long(), torch.from_numpy(z["target"]).long()
    base = float(zo.train_loss(c, variables, src, tgt)[0])
    zo.DROP = zo.make_drop(77, {"emb": 0.1, "att": 0.1, "relu": 0.1, "res": 0.1})
    try:
        dropped = float(zo.train_loss(c, variables, src, tgt)[0])
    finally:
        zo.DROP = None
    assert abs(base - float(z["loss"])) < 2e-5 and abs(dropped - base) > 1e-3


def _search_variants():
    import json
    import os
    from tests.golden_util import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "search_variants.npz"), allow_pickle=False)
    variants = json.loads(str(z["variants_json"]))
    return z, [(m, v, variants[v]) for m in json.loads(str(z["models_json"])) for v in sorted(variants)]


@pytest.mark.parametrize("model,variant,over", _search_variants()[1], ids=lambda x: x if isinstance(x, str) else "")
def test_beam_search_hyper_parameters_away_from_the_defaults(model, variant, over):
    """tests/golden/make_search_golden.py: the reference's search.py on the golden models' weights with greedy search,
    beam widths 2 / 3 / 5, decode_alpha 0 / 0.2 / 1, temperatures 0.7 / 1.5, decode_length 0 / 3 / 10 — the oracle's
    beam_search must return the same sequences (bit-exact), scores and number of decoder calls."""
    from zero_b200.params import HParams
    zs, _ = _search_variants()
    z, hp, variables, _, vs, vt = load_golden(model)
    hp = HParams(**dict(hp.values(), **over))
    c = zo.Cfg(hp, vs, vt)
    assert (c.beam, c.alpha, c.decode_length, c.temperature) == (
        over.get("beam_size", 4), over.get("decode_alpha", 0.6), over.get("decode_length", 6),
        over.get("beam_search_temperature", 1.0))
    with torch.no_grad():
        enc_fn, dec_fn = zo.make_infer_fns(c, variables)
        out = zo.beam_search(c, torch.from_numpy(z["source"]), enc_fn, dec_fn)
    key = "%s:%s" % (model, variant)
    np.testing.assert_array_equal(out["seq"].numpy(), zs[key + ":seq"])
    np.testing.assert_allclose(out["score"].numpy(), zs[key + ":score"], atol=2e-5, rtol=1e-5)
    assert out["steps"] + 1 == int(zs[key + ":calls"])       # + the cache_init dummy call (search.py:56-77)
