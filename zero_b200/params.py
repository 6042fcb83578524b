"""Config surface of Zero's run.py, kept name-for-name (reference run.py:24-239, 241-246, 250-272, 367-376).

`HParams` re-does the slice of tf.contrib.training.HParams the reference relies on:
attribute access, `values()`, `parse("k=v,k2=v2")`, `override_from_dict`, `to_json` / `parse_json`,
`add_hparam`, shallow `copy.copy`.  `global_params()` returns the reference's defaults; only the keys the
Transformer hot path consumes matter to the kernels, the rest are carried so that a
reference `param.json` round-trips.
"""
from __future__ import annotations

import ast
import copy
import json


class HParams(object):
    def __init__(self, **kwargs):
        object.__setattr__(self, "_hp", {})
        for k, v in kwargs.items():
            self.add_hparam(k, v)

    # -- tf.contrib.training.HParams API subset --------------------------------------------------
    def add_hparam(self, name, value):
        if name in self._hp:
            raise ValueError("Hyperparameter name is reserved: %s" % name)
        self._hp[name] = value

    def values(self):
        return dict(self._hp)

    def __getattr__(self, name):
        hp = object.__getattribute__(self, "_hp")
        if name in hp:
            return hp[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self._hp[name] = value

    def __contains__(self, name):
        return name in self._hp

    def __copy__(self):
        new = HParams()
        object.__setattr__(new, "_hp", dict(self._hp))
        return new

    def __deepcopy__(self, memo):
        new = HParams()
        object.__setattr__(new, "_hp", copy.deepcopy(self._hp, memo))
        return new

    def set_hparam(self, name, value):
        if name not in self._hp:
            raise KeyError(name)
        self._hp[name] = self._coerce(name, value)

    def _coerce(self, name, value):
        old = self._hp.get(name)
        if old is None or isinstance(value, type(old)):
            return value
        if isinstance(old, bool):
            if isinstance(value, str):
                if value.lower() in ("true", "1"):
                    return True
                if value.lower() in ("false", "0"):
                    return False
                raise ValueError("Could not parse bool hparam %s=%r" % (name, value))
            return bool(value)
        if isinstance(old, int) and not isinstance(old, bool):
            return int(value)
        if isinstance(old, float):
            return float(value)
        if isinstance(old, str):
            return str(value)
        if isinstance(old, list):
            if isinstance(value, str):
                value = ast.literal_eval(value)
            if not isinstance(value, (list, tuple)):
                value = [value]
            if old:
                return [type(old[0])(v) for v in value]
            return list(value)
        return value

    def override_from_dict(self, values_dict):
        for k, v in values_dict.items():
            self.set_hparam(k, v)
        return self

    def parse(self, values):
        """`name=value,name2=value2`; list values as `name=[a,b]` (run.py:375)."""
        if not values:
            return self
        depth, cur, parts = 0, "", []
        for ch in values:
            if ch == "[":
                depth += 1
            elif ch == "]":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += ch
        if cur:
            parts.append(cur)
        for part in parts:
            if not part.strip():
                continue
            if "=" not in part:
                raise ValueError("Could not parse hparam assignment %r" % part)
            k, v = part.split("=", 1)
            k, v = k.strip(), v.strip()
            if k not in self._hp:
                raise ValueError("Unknown hyperparameter: %s" % k)
            self.set_hparam(k, v)
        return self

    def to_json(self, **kw):
        return json.dumps({k: v for k, v in self._hp.items() if _jsonable(v)}, sort_keys=True, **kw)

    def parse_json(self, values_json):
        return self.override_from_dict({k: v for k, v in json.loads(values_json).items() if k in self._hp})


def _jsonable(v):
    try:
        json.dumps(v)
        return True
    except TypeError:
        return False


def global_params() -> HParams:
    """Defaults, name for name, from reference run.py:24-239."""
    return HParams(
        shared_source_target_embedding=False, shared_target_softmax_embedding=True,
        decode_length=50, beam_size=4, decode_alpha=0.6, enable_noise_beam_search=False,
        beam_search_temperature=1.0, top_beams=1, search_mode="cache",
        max_relative_position=16,
        nstable=4, lrdecay_start=600000, lrdecay_end=1200000, warmup_steps=400, lrate_strategy="gnmt+",
        lrate_decay=0.5, lrate_patience=1, cosine_period=5000, cosine_factor=1,
        estop_patience=100,
        initializer="uniform", initializer_gain=0.08,
        hidden_size=1000, embed_size=620, dropout=0.1, relu_dropout=0.1, residual_dropout=0.1,
        label_smooth=0.1, model_name="rnnsearch", scope_name="rnnsearch", cell="atr", caencoder=True,
        layer_norm=False, use_deep_att=False, swap_memory=True,
        filter_size=2048, attention_dropout=0.1, num_encoder_layer=6, num_decoder_layer=6, num_heads=8,
        aan_mask=True, use_ffn=False,
        max_len=100, eval_max_len=1000000, batch_size=80, token_size=3000, batch_or_token="token",
        eval_batch_size=32, shuffle_batch=True,
        strategies=["aan"],
        process_num=1, buffer_size=100, input_queue_size=100, output_queue_size=100,
        src_vocab_file="", tgt_vocab_file="", src_train_file="", tgt_train_file="", src_dev_file="",
        tgt_dev_file="", src_test_file="", tgt_test_file="", output_dir="", test_output="",
        pretrained_model="",
        beta1=0.9, beta2=0.999, epsilon=1e-9, clip_grad_norm=5.0, gnorm_upper_bound=1e20,
        lrate=1e-5, min_lrate=0.0, max_lrate=1.0,
        epoches=10, update_cycle=1, gpus=[0],
        safe_nan=False, dl4mt_redict=True, ema_decay=-1.0, data_leak_ratio=0.5,
        deep_transformer_init=False,
        disp_freq=100, eval_freq=10000, save_freq=5000, sample_freq=1000, checkpoints=5,
        best_checkpoints=1, max_training_steps=1000,
        nthreads=6, random_seed=1234, train_continue=True,
        default_dtype="float32", dtype_epsilon=1e-8, dtype_inf=1e8, loss_scale=1.0,
        l0_norm_reg_scalar=1.0, l0_norm_start_reg_ramp_up=0, l0_norm_end_reg_ramp_up=10000,
        l0_norm_warm_up=True,
    )


class SimpleVocab(object):
    """The slice of reference vocab.Vocab the hot path uses: size() and the fixed special ids (vocab.py:20-22)."""

    def __init__(self, size):
        self._size = int(size)

    def size(self):
        return self._size

    @staticmethod
    def pad():
        return 0

    @staticmethod
    def unk():
        return 1

    @staticmethod
    def eos():
        return 2


def merge_params(defaults: HParams, saved_json: str | None = None, config: dict | None = None,
                 cmdline: str = "") -> HParams:
    """Precedence of reference run.py:367-376: defaults < saved param.json < --config dict < --parameters."""
    p = copy.copy(defaults)
    if saved_json:
        p.parse_json(saved_json)
    if config:
        p.override_from_dict({k: v for k, v in config.items() if k in p})
    if cmdline:
        p.parse(cmdline)
    return p


def transformer_base(**overrides) -> HParams:
    """Transformer-base recipe of docs/l0drop/README.md:81-106 (d=512, f=2048, h=8, 6+6)."""
    p = global_params()
    p.override_from_dict(dict(
        hidden_size=512, embed_size=512, filter_size=2048, num_heads=8, num_encoder_layer=6,
        num_decoder_layer=6, model_name="transformer", scope_name="transformer",
        initializer="uniform_unit_scaling", initializer_gain=1.0, dropout=0.0, relu_dropout=0.0,
        residual_dropout=0.0, attention_dropout=0.0, label_smooth=0.1, lrate_strategy="noam",
        warmup_steps=4000, beta1=0.9, beta2=0.98, epsilon=1e-8, lrate=1.0, clip_grad_norm=0.0))
    p.override_from_dict(overrides)
    return p
