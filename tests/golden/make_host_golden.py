"""Golden vectors for the host-side pieces, produced by EXECUTING the reference's own modules
(utils/metric.py, utils/util.py indexers, data.py Dataset, vocab.py, lrs/*) from the read-only checkout.
utils/util.py and data.py import TensorFlow at module level, so they run over oracle/tf1_shim like
make_golden.py does.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_host_golden.py   ->  tests/golden/host_golden.json
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("ZERO_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf1_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import tensorflow as tf  # noqa: E402,F401  (the shim)
import data as ref_data  # noqa: E402
import lrs as ref_lrs  # noqa: E402
import vocab as ref_vocab  # noqa: E402
from utils import metric as ref_metric  # noqa: E402
from utils import util as ref_util  # noqa: E402

from zero_b200.data import synthetic_corpus  # noqa: E402


class P(object):
    pass


def lr_params(strategy, **kw):
    p = P()
    p.lrate_strategy = strategy
    p.lrate, p.min_lrate, p.max_lrate = 1.0, 0.0, 10.0
    p.warmup_steps, p.hidden_size = 400, 128
    p.nstable, p.lrdecay_start, p.lrdecay_end = 4, 600, 1200
    p.lrate_decay, p.lrate_patience = 0.5, 1
    p.cosine_factor, p.cosine_period = 1, 500
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def main():
    rng = np.random.RandomState(7)
    out = {}
    # ---- BLEU: random candidate / multi-reference corpora over a small alphabet (so n-grams do match)
    cases = []
    for case in range(6):
        nsent = int(rng.randint(1, 12))
        nref = int(rng.randint(1, 4))
        cand, refs = [], []
        for _ in range(nsent):
            base = [str(x) for x in rng.randint(0, 6, size=rng.randint(1, 14))]
            c = [t for t in base if rng.rand() > 0.15] or base[:1]
            if rng.rand() < 0.3:
                c = c + [str(x) for x in rng.randint(0, 6, size=rng.randint(1, 4))]
            rs = []
            for _ in range(nref):
                r = [t if rng.rand() > 0.2 else str(int(rng.randint(0, 6))) for t in base]
                if rng.rand() < 0.3:
                    r = r[:-1] or r
                rs.append(r)
            cand.append(c)
            refs.append(rs)
        smooth = bool(case % 2)
        bp = "closest" if case % 3 else "shortest"
        cases.append({"cand": cand, "refs": refs, "smooth": smooth, "bp": bp,
                      "bleu": ref_metric.bleu(cand, refs, bp=bp, smooth=smooth)})
    cases.append({"cand": [["a", "b", "c", "d", "e"]], "refs": [[["a", "b", "c", "d", "e"]]], "smooth": False,
                  "bp": "closest", "bleu": ref_metric.bleu([["a", "b", "c", "d", "e"]], [[["a", "b", "c", "d", "e"]]])})
    out["bleu"] = cases
    # ---- indexers
    idx = []
    for _ in range(8):
        n = int(rng.randint(1, 60))
        lens = [[int(rng.randint(1, 30)), int(rng.randint(1, 30))] for _ in range(n)]
        tok = int(rng.randint(20, 200))
        idx.append({"lens": lens, "token_size": tok, "token_indexer": ref_util.token_indexer(lens, tok),
                    "batch_size": int(tok // 10 + 1), "batch_indexer": ref_util.batch_indexer(n, int(tok // 10 + 1))})
    out["indexers"] = idx
    # ---- Dataset.batcher on the C1 synthetic corpus (files written to a temp dir; shuffle off and on)
    corp = synthetic_corpus(n_train=97, n_heldout=8, seed=5)
    tmp = tempfile.mkdtemp()
    vf = os.path.join(tmp, "vocab.txt")
    with open(vf, "w") as f:
        for s in corp["symbols"]:
            f.write(s + "\n")
    sf, tf_ = os.path.join(tmp, "src.txt"), os.path.join(tmp, "tgt.txt")
    with open(sf, "w") as f:
        for r in corp["train_src"]:
            f.write(" ".join(r) + "\n")
    with open(tf_, "w") as f:
        for r in corp["train_tgt"]:
            f.write(" ".join(r) + "\n")
    v = ref_vocab.Vocab(vf)
    runs = []
    for mode, size, shuffle, train, buf in (("batch", 16, False, True, 40), ("token", 150, False, True, 50),
                                            ("token", 150, True, True, 50), ("batch", 16, False, False, 1000)):
        ds = ref_data.Dataset(sf, tf_, v, v, max_len=20, batch_or_token=mode, data_leak_ratio=0.5)
        np.random.seed(11)
        epochs = []
        for _ in range(2):  # two passes: the leak buffer carries over
            epochs.append([{"index": [int(i) for i in b["index"]], "src": b["src"].tolist(), "tgt": b["tgt"].tolist()}
                           for b in ds.batcher(size, buffer_size=buf, shuffle=shuffle, train=train)])
        runs.append({"mode": mode, "size": size, "shuffle": shuffle, "train": train, "buffer_size": buf,
                     "epochs": epochs})
    out["batcher"] = {"corpus_seed": 5, "n_train": 97, "max_len": 20, "runs": runs,
                      "vocab_size": v.size(), "ids_w3": v.to_id(["w3", "nope", "w10"])}
    # ---- learning-rate schedules
    lr = {}
    for name, kw in (("noam", {}), ("gnmt+", {}), ("vanilla", {}), ("cosine", {}), ("cosine", {"cosine_factor": 2}),
                     ("epoch", {}), ("score", {})):
        p = lr_params(name, **kw)
        if name == "score":
            class R(object):
                valid_script_scores = []
            p.recorder = R()
        sched = ref_lrs.get_lr(p)
        vals = []
        if name == "epoch":
            for e in range(1, 5):
                sched.after_epoch(eidx=e)
                vals.append(sched.get_lr())
        elif name == "score":
            for s in (0.1, 0.2, 0.15, 0.18, 0.3, 0.3):
                sched.after_eval(s)
                vals.append(sched.get_lr())
        else:
            for t in (0, 1, 10, 399, 400, 401, 650, 900, 1199, 1500, 2500):
                sched.step(t)
                vals.append(float(sched.get_lr()))
        lr[name + ("_tmult2" if kw else "")] = vals
    out["lrs"] = lr
    with open(os.path.join(HERE, "host_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote host_golden.json:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
