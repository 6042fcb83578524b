"""The Transformer-family plugins of the hot path, registered under the reference's names
(models/transformer.py:289, transformer_aan.py:331, transformer_rpr.py:297, transformer_rela.py:290,
transformer_fuse.py:269).

Callable contract, tensors being torch tensors instead of TF ones:
  train_fn(features, params, initializer=None) -> {"loss": fp32[1]}      (also leaves d loss / d params in the
        engine's gradient arena: the reference gets them from optimizer.compute_gradients, main.py:28)
  score_fn(features, params, initializer=None) -> {"score": fp32[B]}
  infer_fn(params) -> (encoding_fn(source) -> state, decoding_fn(target, state, time) -> (logits, state))
The TF variable store is replaced by one Engine per (scope_name, model_name), created on first use and kept in
`_engines` (the analogue of tf.AUTO_REUSE variable scopes).
"""
import copy

from . import model
from ..engine import Engine

_engines = {}


def get_engine(params, initializer=None):
    key = (params.scope_name or "model", str(params.model_name).lower())
    eng = _engines.get(key)
    if eng is None:
        eng = Engine(params)
        eng.ps.init_random(int(getattr(params, "random_seed", 1234)))
        _engines[key] = eng
    return eng


def reset_engines():
    _engines.clear()


def _closing_dropout(params):
    """utils/util.py:106-114."""
    for k in list(params.values()):
        if "dropout" in k or "label_smoothing" in k:
            setattr(params, k, 0.0)
    return params


def _make(name):
    def train_fn(features, params, initializer=None):
        eng = get_engine(params, initializer)
        eng.advance_dropout_seed()   # dropout rates come from params at engine creation (utils/util.py:75-79)
        loss = eng.forward_backward(features["source"], features["target"])
        return {"loss": loss}

    def score_fn(features, params, initializer=None):
        params = _closing_dropout(copy.copy(params))
        eng = get_engine(params, initializer)
        return {"score": eng.score(features["source"], features["target"])}

    def infer_fn(params):
        params = _closing_dropout(copy.copy(params))
        eng = get_engine(params)
        eng.decode_length = int(params.decode_length)
        if getattr(params, "search_mode", "cache") != "cache":
            # 'dev' (search.py:129-140, models/transformer.py:276-281): the decoder re-run on the whole prefix every
            # step, no caches.  (The reference's own branch calls encoder(state, params) with the state dict where the
            # source ids are expected and cannot run as written; the documented intent is what is built.)
            return eng.encoding_fn, eng.decoding_fn_dev
        return eng.encoding_fn, eng.decoding_fn

    model.model_register(name, train_fn, score_fn, infer_fn)


for _name in ("transformer", "transformer_aan", "transformer_rpr", "transformer_rela", "transformer_fuse"):
    _make(_name)
