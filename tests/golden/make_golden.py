"""Generate golden vectors by EXECUTING the reference's own Python (read-only /root/reference) over the eager
TF1 shim in oracle/tf1_shim.  Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Writes tests/golden/<model>.npz: every variable the reference created (by its TF name), the inputs, and the
reference's outputs — train loss, d loss / d variable (torch autograd through the reference's forward graph),
score_fn per-sentence scores, teacher-forced logits, beam-search sequences/scores and the per-step logits of
the first decode steps.  Dropout is 0 everywhere (RNG streams are not comparable across frameworks).
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("ZERO_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf1_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import tensorflow as tf  # noqa: E402  (the shim)

from zero_b200.params import SimpleVocab, global_params  # noqa: E402

import search as ref_search  # noqa: E402
from models import model as ref_model  # noqa: E402
from modules import initializer as ref_init  # noqa: E402
from utils import dtype as ref_dtype  # noqa: E402
import models.transformer  # noqa: E402,F401
import models.transformer_aan  # noqa: E402,F401
import models.transformer_rpr  # noqa: E402,F401
import models.transformer_rela  # noqa: E402,F401
import models.transformer_fuse  # noqa: E402,F401


def make_params(model_name, **kw):
    p = global_params()
    p.override_from_dict(dict(
        hidden_size=128, embed_size=128, filter_size=256, num_heads=2, num_encoder_layer=2,
        num_decoder_layer=2, model_name=model_name, scope_name=model_name,
        initializer="uniform_unit_scaling", initializer_gain=1.0,
        dropout=0.0, relu_dropout=0.0, residual_dropout=0.0, attention_dropout=0.0, label_smooth=0.1,
        beam_size=4, decode_length=6, decode_alpha=0.6, max_relative_position=4))
    p.override_from_dict({k: v for k, v in kw.items() if not k.startswith("_")})   # _vs / _vt: vocabulary sizes
    p.add_hparam("src_vocab", SimpleVocab(kw.get("_vs", 200)))
    p.add_hparam("tgt_vocab", SimpleVocab(kw.get("_vt", 208)))
    return p


def synth_batch(rng, batch, smax, tmax, vs, vt):
    src = np.zeros((batch, smax), np.int64)
    tgt = np.zeros((batch, tmax), np.int64)
    for b in range(batch):
        ls = smax if b == 0 else int(rng.integers(3, smax + 1))
        lt = tmax if b == 1 else int(rng.integers(2, tmax + 1))
        src[b, :ls - 1] = rng.integers(3, vs, ls - 1)
        src[b, ls - 1] = 2
        tgt[b, :lt - 1] = rng.integers(3, vt, lt - 1)
        tgt[b, lt - 1] = 2
    return src, tgt


def run(model_name, out_name=None, seed=7, shape=(5, 11, 9), **kw):
    tf.reset_default_graph(seed=1000 + seed)
    ref_dtype.set_floatx("float32")
    p = make_params(model_name, **kw)
    rng = np.random.default_rng(seed)
    vs, vt = p.src_vocab.size(), p.tgt_vocab.size()
    src, tgt = kw["_batch"] if "_batch" in kw else synth_batch(rng, shape[0], shape[1], shape[2], vs, vt)
    feats = {"source": tf.constant(src), "target": tf.constant(tgt)}
    graph = ref_model.get_model(model_name)
    init = ref_init.get_initializer(p.initializer, p.initializer_gain)

    out = graph.train_fn(feats, p, initializer=init)
    loss = out["loss"]
    variables = tf.all_variables()
    # perturb biases / LN params away from their 0/1 init so that parity is not vacuous
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, v in variables.items():
            if name.endswith("/b_0") or name.endswith("/offset") or name.endswith("/bias") or name.endswith("/gate"):
                v.add_(torch.randn(v.shape, generator=g) * 0.05)
            if name.endswith("/scale"):
                v.add_(torch.randn(v.shape, generator=g) * 0.05)
    out = graph.train_fn(feats, p, initializer=init)
    loss = out["loss"]
    names = list(variables.keys())
    grads = torch.autograd.grad(loss, [variables[n] for n in names], allow_unused=True)
    res = {"source": src, "target": tgt, "loss": loss.detach().numpy()}
    for n, gr in zip(names, grads):
        res["var:" + n] = variables[n].detach().numpy()
        res["grad:" + n] = (torch.zeros_like(variables[n]) if gr is None else gr).detach().numpy()

    with torch.no_grad():
        sc = graph.score_fn(feats, p, initializer=init)
        res["score"] = sc["score"].detach().numpy()
        # teacher-forced logits: re-run decoder pieces through train graph internals
        mod = sys.modules["models." + model_name]
        with tf.variable_scope(p.scope_name, reuse=tf.AUTO_REUSE, dtype=tf.float32,
                               custom_getter=ref_dtype.float32_variable_storage_getter):
            state = mod.encoder(feats["source"], p)
            _, logits, _, per_sample = mod.decoder(feats["target"], state, p)
        res["encodes"] = state["encodes"].detach().numpy()
        res["logits"] = logits.detach().numpy()
        res["per_sample_loss"] = per_sample.detach().numpy()

        # beam search through the reference's search.py
        pp = copy.copy(p)
        enc_fn, dec_fn = graph.infer_fn(pp)
        step_logits = []

        def dec_fn_rec(target, state, time):
            lg, st = dec_fn(target, state, time)
            step_logits.append(lg.detach().numpy().copy())
            return lg, st

        bs = ref_search.beam_search({"source": feats["source"]}, enc_fn, dec_fn_rec, pp)
        res["beam_seq"] = bs["seq"].detach().numpy()
        res["beam_score"] = bs["score"].detach().numpy()
        # step_logits[0] is the cache_init dummy step (search.py:56-77); [1] is t=0, [2] is t=1
        for i, lg in enumerate(step_logits[:4]):
            res["step_logits_%d" % i] = lg
        res["n_decode_calls"] = np.asarray(len(step_logits))
    res["params_json"] = np.asarray(p.to_json())
    path = os.path.join(HERE, (out_name or model_name) + ".npz")
    np.savez_compressed(path, **res)
    print("%-22s loss %.6f  vars %d  beam_seq %s  decode calls %d -> %s" % (
        out_name or model_name, float(res["loss"]), len(names), res["beam_seq"].shape, len(step_logits),
        os.path.relpath(path, ROOT)))


def main_small():
    small = dict(hidden_size=64, embed_size=64, filter_size=128, num_heads=2)   # dh = 32
    run("transformer")                                                          # d = 128, dh = 64
    run("transformer", out_name="transformer_h4", seed=11, **dict(small, num_heads=4))  # dh = 16
    run("transformer_aan", **small)
    run("transformer_aan", out_name="transformer_aan_cumsum", aan_mask=False, use_ffn=True, seed=9, **small)
    run("transformer_rpr", **small)
    run("transformer_rela", **small)
    run("transformer_fuse", **small)


def main_long():
    """Sequences longer than 16 tokens with dh = 64: the tensor-core attention kernels of the training step (which
    the library only selects from 16 query rows on) run inside a reference-executed model test, masks included."""
    long = dict(hidden_size=128, embed_size=128, filter_size=256, num_heads=2, num_encoder_layer=1,
                num_decoder_layer=1, decode_length=8)
    for name, model, extra in (("transformer_len40", "transformer", {}),
                               ("transformer_rpr_len40", "transformer_rpr", dict(max_relative_position=6)),
                               ("transformer_rela_len40", "transformer_rela", {}),
                               ("transformer_fuse_len40", "transformer_fuse", {})):
        run(model, out_name=name, seed=21, shape=(4, 40, 36), **dict(long, **extra))


def main_embeddings():
    """The two embedding-sharing switches away from their defaults (models/transformer.py:21-22,96-97,186-189): one
    table for source, target and soft-max ("embedding"), and a soft-max table of its own ("softmax_embedding")."""
    small = dict(hidden_size=64, embed_size=64, filter_size=128, num_heads=2)   # dh = 32
    run("transformer", out_name="transformer_shared_emb", seed=31, shared_source_target_embedding=True,
        _vs=208, _vt=208, **small)
    run("transformer", out_name="transformer_softmax_emb", seed=32, shared_target_softmax_embedding=False, **small)
    run("transformer_aan", out_name="transformer_aan_shared_emb", seed=33, shared_source_target_embedding=True,
        _vs=208, _vt=208, **small)


def main_edge():
    """Edge cases of the data: sentences of ONE token (the end-of-sentence mark alone) on either side next to full-
    length ones, a batch of one sentence, and label_smooth = 0 (utils/util.py:88-103 takes its one-hot branch)."""
    small = dict(hidden_size=64, embed_size=64, filter_size=128, num_heads=2)   # dh = 32
    src = np.array([[5, 9, 7, 4, 11, 2], [2, 0, 0, 0, 0, 0], [8, 6, 2, 0, 0, 0], [12, 13, 14, 15, 2, 0]], np.int64)
    tgt = np.array([[2, 0, 0, 0, 0], [6, 7, 8, 9, 2], [10, 2, 0, 0, 0], [3, 3, 3, 2, 0]], np.int64)
    run("transformer", out_name="transformer_edge", seed=41, _batch=(src, tgt), **small)
    run("transformer_aan", out_name="transformer_aan_edge", seed=42, _batch=(src, tgt), **small)
    run("transformer", out_name="transformer_edge_b1_nosmooth", seed=43, label_smooth=0.0,
        _batch=(src[3:4, :5], tgt[1:2]), **small)


if __name__ == "__main__":
    if "--edge-only" in sys.argv:
        main_edge()
        sys.exit(0)
    if "--embeddings-only" in sys.argv:
        main_embeddings()
        sys.exit(0)
    if "--long-only" not in sys.argv:
        main_small()
        main_embeddings()
        main_edge()
    main_long()
