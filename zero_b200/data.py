"""Host-side data path feeding the hot loop (reference data.py:11-117, utils/util.py:17-65, main.py:242-273).

* `batch_indexer` / `token_indexer` — sentence-count and token-budget batch boundaries (utils/util.py:17-65).
* `Dataset` — line-parallel corpus reader, length-sorted bucketed batcher with the "leak buffer" that carries
  undersized tail batches into the next buffer (data.py:67-117), `to_matrix` zero-padded int32 id matrices
  (data.py:47-65).  Sources may be files (like the reference) or in-memory token lists (synthetic corpora).
* `shard_for_rank` — the reference feeds `len(gpus)` consecutive batches to the towers of one step
  (main.py:268-273); with one process per GPU, rank r of N takes batch N * step + r.
* `synthetic_corpus` — BASELINE.json configs[0] / SURVEY.md 8(d) C1: a learnable token transduction over a 1k
  vocabulary (256 training pairs + 64 held-out, none of which is a training sentence).

Batches are numpy int32 matrices; `pin()` turns one into pinned torch tensors for the H2D copy of a step.
"""
from __future__ import annotations

import numpy as np


def batch_indexer(datasize, batch_size):
    """Consecutive index groups of `batch_size`; the remainder forms a last, smaller group (utils/util.py:17-27)."""
    idx = list(range(int(datasize)))
    return [idx[s:s + batch_size] for s in range(0, len(idx), batch_size)]


def token_indexer(dataset, token_size):
    """Token-budget batches over per-sample length tuples (utils/util.py:30-65): a group is closed, WITHOUT the
    sample that made it overflow, as soon as `group_size * max_length_so_far >= token_size` on any length column;
    a sample that overflows on its own becomes a batch of one."""
    n = len(dataset)
    if n == 0:
        return []
    cols = len(dataset[0])
    groups, start, i = [], 0, 0
    width = [0] * cols
    while i < n:
        width = [max(w, l) for w, l in zip(width, dataset[i])]
        count = i - start + 1
        if any(count * w >= token_size for w in width):
            if count > 1:
                groups.append(list(range(start, i)))
                start = i           # sample i opens the next group (it is looked at again)
            else:
                groups.append([i])
                start = i = i + 1
            width = [0] * cols
        else:
            i += 1
    if start < n:
        groups.append(list(range(start, n)))
    return groups


class Dataset(object):
    def __init__(self, src_file, tgt_file, src_vocab, tgt_vocab, max_len=100, batch_or_token="batch",
                 data_leak_ratio=0.5):
        self.source, self.target = src_file, tgt_file
        self.src_vocab, self.tgt_vocab = src_vocab, tgt_vocab
        self.max_len = max_len
        self.batch_or_token = batch_or_token
        self.data_leak_ratio = data_leak_ratio
        self.leak_buffer = []

    # -- reading -----------------------------------------------------------------------------------
    @staticmethod
    def _lines(src):
        if isinstance(src, str):
            with open(src, "r") as f:
                for line in f:
                    yield line
        else:  # in-memory corpus: a list of token lists (or strings)
            for item in src:
                yield item if isinstance(item, str) else " ".join(item) + "\n"

    def load_data(self):
        """(src_ids, tgt_ids) pairs; stops at the shorter file, skips pairs with an empty side (data.py:26-45)."""
        for s, t in zip(self._lines(self.source), self._lines(self.target)):
            s, t = s.strip(), t.strip()
            if not s or not t:
                continue
            yield (self.src_vocab.to_id(s.split()[:self.max_len]), self.tgt_vocab.to_id(t.split()[:self.max_len]))

    # -- batching ----------------------------------------------------------------------------------
    def to_matrix(self, batch):
        """Zero-padded int32 matrices, width = min(max_len, longest) (data.py:47-65)."""
        ws = min(self.max_len, max(len(b[1]) for b in batch))
        wt = min(self.max_len, max(len(b[2]) for b in batch))
        s = np.zeros((len(batch), ws), dtype=np.int32)
        t = np.zeros((len(batch), wt), dtype=np.int32)
        for r, (_, src_ids, tgt_ids) in enumerate(batch):
            s[r, :min(ws, len(src_ids))] = src_ids[:ws]
            t[r, :min(wt, len(tgt_ids))] = tgt_ids[:wt]
        return [b[0] for b in batch], s, t

    def _batches_of(self, buf, size, shuffle):
        buf = sorted(buf, key=lambda b: max(len(b[1]), len(b[2])))
        if self.batch_or_token == "batch":
            groups = batch_indexer(len(buf), size)
        else:
            groups = token_indexer([(len(b[1]), len(b[2])) for b in buf], size)
        order = list(range(len(groups)))
        if shuffle:
            np.random.shuffle(order)   # the reference shuffles whole batches with the global numpy RNG
        for g in order:
            rows = [buf[i] for i in groups[g]]
            index, s, t = self.to_matrix(rows)
            yield {"src": s, "tgt": t, "index": index, "raw": rows}

    def _weight(self, data):
        if self.batch_or_token == "batch":
            return len(data["raw"])
        return max(int(np.sum(data["tgt"] > 0)), int(np.sum(data["src"] > 0)))

    def batcher(self, size, buffer_size=1000, shuffle=True, train=True):
        """data.py:67-117.  Batches lighter than size * data_leak_ratio are not yielded but carried over (their
        samples re-enter the next buffer); at the end of the data they are yielded only when `train` is False."""
        buf, self.leak_buffer = self.leak_buffer, []
        for i, (s, t) in enumerate(self.load_data()):
            buf.append((i, s, t))
            if len(buf) >= buffer_size:
                for data in self._batches_of(buf, size, shuffle):
                    if self._weight(data) < size * self.data_leak_ratio:
                        self.leak_buffer += data["raw"]
                    else:
                        yield data
                buf, self.leak_buffer = self.leak_buffer, []
        if buf:
            for data in self._batches_of(buf, size, shuffle):
                if train and self._weight(data) < size * self.data_leak_ratio:
                    self.leak_buffer += data["raw"]
                else:
                    yield data


def shard_for_rank(batches, world, rank):
    """Step k consumes batches [k*world, (k+1)*world); rank r gets the r-th of them (main.py:268-273).  A trailing
    incomplete group is dropped, like the reference's `continue` until every tower has data."""
    group = []
    for b in batches:
        group.append(b)
        if len(group) == world:
            yield group[rank]
            group = []


def pin(data):
    """numpy id matrices of one batch -> pinned int32 torch tensors (source, target)."""
    import torch
    s, t = torch.from_numpy(np.ascontiguousarray(data["src"])), torch.from_numpy(np.ascontiguousarray(data["tgt"]))
    if torch.cuda.is_available():
        s, t = s.pin_memory(), t.pin_memory()
    return s, t


def synthetic_corpus(n_train=256, n_heldout=64, n_symbols=997, min_len=12, max_len=24, zipf=0.6, seed=1234,
                     active_symbols=16, swap_pairs=False):
    """C1 of SURVEY.md 8(d): a 1k vocabulary (3 specials + symbols w0..w{n-1}); source length ~U[min_len, max_len],
    tokens ~Zipf(zipf) over the `active_symbols` most frequent symbols (so 256 sentences cover every symbol many
    times and the held-out set measures generalisation, not vocabulary coverage); the target is a fixed token
    substitution (a permutation of the whole symbol set) of the source, optionally with every adjacent pair swapped
    (local reordering) — a lexical mapping plus positional alignment that a 2-layer model learns, and generalises,
    within about a thousand updates."""
    rng = np.random.RandomState(seed)
    symbols = ["w%d" % i for i in range(n_symbols)]
    perm = rng.permutation(n_symbols)
    k = min(active_symbols, n_symbols)
    p = 1.0 / np.arange(1, k + 1) ** zipf
    p /= p.sum()

    def transduce(ids):
        out = list(ids)
        if swap_pairs:
            for i in range(0, len(out) - 1, 2):
                out[i], out[i + 1] = out[i + 1], out[i]
        return [int(perm[i]) for i in out]

    def draw():
        n = int(rng.randint(min_len, max_len + 1))
        return [int(i) for i in rng.choice(k, size=n, p=p)]

    train = [draw() for _ in range(n_train)]
    seen = set(tuple(s) for s in train)
    held = []
    while len(held) < n_heldout:
        s = draw()
        if tuple(s) not in seen:
            held.append(s)

    def text(rows):
        return [[symbols[i] for i in r] for r in rows]

    return {"symbols": symbols,
            "train_src": text(train), "train_tgt": text([transduce(s) for s in train]),
            "dev_src": text(held), "dev_tgt": text([transduce(s) for s in held])}
