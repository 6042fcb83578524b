"""Golden vectors for the SEARCH hyper-parameters away from the values the model goldens use (beam 4, alpha 0.6,
temperature 1, decode_length 6): the reference's own search.py (search.py:19-275) run over the TF1 shim on the weights
of existing model goldens, for greedy search (beam 1), no length penalty (alpha 0), a strong one (alpha 1), beam
widths 2 / 3 / 5, temperatures below and above 1 and decode lengths 0 / 3 / 10.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_search_golden.py

Writes tests/golden/search_variants.npz: per (model, variant) the reference's `seq` / `score` and the number of decoder
calls, plus the variant table itself as json.  tests/test_oracle_golden.py pins oracle.beam_search to them.
"""
import copy
import json
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

from zero_b200.params import HParams, SimpleVocab  # noqa: E402

import search as ref_search  # noqa: E402
from models import model as ref_model  # noqa: E402
from utils import dtype as ref_dtype  # noqa: E402
import models.transformer  # noqa: E402,F401
import models.transformer_aan  # noqa: E402,F401
import models.transformer_rpr  # noqa: E402,F401

VARIANTS = {
    "greedy": dict(beam_size=1),
    "beam3_no_length_penalty": dict(beam_size=3, decode_alpha=0.0),
    "beam4_alpha1_len3": dict(beam_size=4, decode_alpha=1.0, decode_length=3),
    "beam2_sharp": dict(beam_size=2, beam_search_temperature=0.7),
    "beam5_len0": dict(beam_size=5, decode_length=0),
    "beam4_flat_len10": dict(beam_size=4, beam_search_temperature=1.5, decode_alpha=0.2, decode_length=10),
}
MODELS = ["transformer_h4", "transformer_aan", "transformer_rpr"]


def main():
    res = {"variants_json": np.asarray(json.dumps(VARIANTS, sort_keys=True)), "models_json": np.asarray(json.dumps(MODELS))}
    for name in MODELS:
        z = np.load(os.path.join(HERE, name + ".npz"), allow_pickle=False)
        base = json.loads(str(z["params_json"]))
        variables = {k[4:]: z[k] for k in z.files if k.startswith("var:")}
        vs = [v for k, v in variables.items() if k.endswith("src_embedding") or k.endswith("/embedding")][0].shape[0]
        vt = [v for k, v in variables.items() if k.endswith("tgt_embedding") or k.endswith("/embedding")][0].shape[0]
        for vname, over in VARIANTS.items():
            tf.reset_default_graph(seed=1)
            ref_dtype.set_floatx("float32")
            store = tf.all_variables()
            for k, v in variables.items():            # the golden's weights, found by name by tf.get_variable
                store[k] = tf.convert_to_tensor(torch.from_numpy(v.copy()))
            p = HParams(**dict(base, **over))
            p.add_hparam("src_vocab", SimpleVocab(vs))
            p.add_hparam("tgt_vocab", SimpleVocab(vt))
            graph = ref_model.get_model(p.model_name)
            calls = {"n": 0}
            with torch.no_grad():
                enc_fn, dec_fn = graph.infer_fn(copy.copy(p))

                def counted(target, state, time):
                    calls["n"] += 1
                    return dec_fn(target, state, time)

                out = ref_search.beam_search({"source": tf.constant(z["source"])}, enc_fn, counted, p)
            assert set(store.keys()) == set(variables.keys()), "the search created variables the golden does not hold"
            key = "%s:%s" % (name, vname)
            res[key + ":seq"] = out["seq"].detach().numpy()
            res[key + ":score"] = out["score"].detach().numpy()
            res[key + ":calls"] = np.asarray(calls["n"])
            print("%-16s %-26s seq %-12s calls %2d  best %s" % (name, vname, res[key + ":seq"].shape, calls["n"],
                                                                 np.round(res[key + ":score"][:, 0], 3)))
    np.savez_compressed(os.path.join(HERE, "search_variants.npz"), **res)


if __name__ == "__main__":
    main()
