"""Golden vector for the vocabulary builder: runs the reference's own vocab.py (py3-importable) over a small corpus
and records the file it writes.  Run in the build container (needs /root/reference); the test reads only the
committed tests/golden/vocab_golden.json.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_vocab_golden.py
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import vocab as ref_vocab  # noqa: E402


def corpus(seed=3, lines=60):
    rng = np.random.RandomState(seed)
    words = ["tok%d" % i for i in range(40)]
    p = 1.0 / np.arange(1, 41)
    p /= p.sum()
    return [" ".join(words[int(i)] for i in rng.choice(40, size=int(rng.randint(1, 12)), p=p)) for _ in range(lines)]


def main():
    text = corpus()
    out = {"corpus_seed": 3, "lines": text, "runs": []}
    for size in (10 ** 6, 12):
        v = ref_vocab.Vocab()
        for line in text:
            for token in line.strip().split():
                v.insert(token)
        v.sort_vocab()
        path = os.path.join(tempfile.mkdtemp(), "vocab.txt")
        v.save_vocab(path, size)
        out["runs"].append({"size": size, "file": open(path).read().splitlines(), "vocab_size": v.size()})
    with open(os.path.join(HERE, "vocab_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote vocab_golden.json:", [len(r["file"]) for r in out["runs"]])


if __name__ == "__main__":
    main()
