"""Decode / score drivers and corpus BLEU around the hot path (reference evalu.py:14-280, utils/metric.py:243-297).

`decoding` is where the reference measures decode throughput (evalu.py:106-120: wall time of beam search per
batch); `bleu` restates utils/metric.py:bleu (multi-reference clipped n-gram precision, 'closest' reference
length, brevity penalty, geometric mean over n = 1..4) and is pinned against values produced by the reference's own
metric.py (tests/golden/host_golden.json).
"""
from __future__ import annotations

import math
import time
from collections import Counter

import numpy as np
import torch

from . import search
from .data import shard_for_rank  # noqa: F401  (re-exported for drivers)


# ---------------------------------------------------------------------------------------------- hypotheses
def decode_target_token(id_seq, vocab):
    """ids -> tokens, cut at the first <eos> / <pad> (evalu.py:14-22)."""
    out = []
    for tok in id_seq:
        tok = int(tok)
        if tok == vocab.eos() or tok == vocab.pad():
            break
        out.append(tok)
    return vocab.to_tokens(out)


def decode_hypothesis(seqs, scores, params, mask=None):
    """Top-1 beam of every sentence of every tower output (evalu.py:25-46).
    seqs: list (towers) of [B, beam, L]; scores: list of [B, beam]."""
    hypos, marks = [], []
    if mask is None:
        mask = [1.0] * len(seqs)
    for _seqs, _scores, m in zip(seqs, scores, mask):
        if m < 1.0:
            continue
        for seq, score in zip(_seqs, _scores):
            hypos.append(decode_target_token(seq[0], params.tgt_vocab))
            marks.append(float(score[0]))
    return hypos, marks


# ---------------------------------------------------------------------------------------------- drivers
def gather_decoded(local, world_size):
    """Every rank's (translations, scores, indices, tokens, seconds) -> the whole corpus on every rank (so that all
    ranks take the same BLEU-driven decisions): lists concatenated in rank order, tokens summed, seconds = the
    slowest rank's.  The reference's towers decode consecutive batches in one session.run (evalu.py:66-104)."""
    if world_size <= 1:
        return local
    import torch.distributed as dist
    parts = [None] * world_size
    dist.all_gather_object(parts, local)
    translations, scores, indices, tokens, seconds = [], [], [], 0, 0.0
    for tr, sc, ix, tk, sec in parts:
        translations.extend(tr)
        scores.extend(sc)
        indices.extend(ix)
        tokens += tk
        seconds = max(seconds, sec)
    return translations, scores, indices, tokens, seconds


def decoding(infer_fns, dataset, params, log=None, world_size=1, rank=0):
    """Beam-search the dev/test set batch by batch (evalu.py:49-139).  Returns translations, scores, sample
    indices and a timing record {sentences, tokens, seconds}: tokens = top-1 hypothesis lengths + 1 (<eos>),
    seconds = device time of beam_search only (the reference's per-batch wall clock around session.run).
    With world_size > 1 rank r decodes every world_size-th batch (what tower r is fed, evalu.py:66-92) and the
    results are exchanged at the end; the order of the returned lists is then by rank, `indices` restores the corpus."""
    encoding_fn, decoding_fn = infer_fns
    translations, scores, indices = [], [], []
    tokens, seconds = 0, 0.0
    for bidx, data in enumerate(dataset.batcher(params.eval_batch_size, buffer_size=params.buffer_size,
                                                shuffle=False, train=False)):
        if bidx % world_size != rank:
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = search.beam_search({"source": torch.from_numpy(data["src"])}, encoding_fn, decoding_fn, params)
        e1.record()
        seq, sc = out["seq"].cpu().numpy(), out["score"].cpu().numpy()
        dt = e0.elapsed_time(e1) * 1e-3
        hyp, mk = decode_hypothesis([seq], [sc], params)
        translations.extend(hyp)
        scores.extend(mk)
        indices.extend(data["index"])
        tokens += sum(len(h) + 1 for h in hyp)
        seconds += dt
        if log:
            log("Decoding Batch %d using %.3f s, translating %d sentences using %.3f s in total" % (
                bidx, dt, len(translations), seconds))
    translations, scores, indices, tokens, seconds = gather_decoded(
        (translations, scores, [int(i) for i in indices], tokens, seconds), world_size)
    return translations, scores, indices, {"sentences": len(translations), "tokens": tokens, "seconds": seconds}


def scoring(score_fn, dataset, params):
    """Teacher-forced per-sentence scores and corpus perplexity (evalu.py:142-246)."""
    scores, indices = [], []
    total_entropy, total_tokens = 0.0, 0.0
    for data in dataset.batcher(params.eval_batch_size, buffer_size=params.buffer_size, shuffle=False, train=False):
        s = score_fn({"source": torch.from_numpy(data["src"]), "target": torch.from_numpy(data["tgt"])}, params)
        s = s["score"].float().cpu().numpy()
        ntok = (data["tgt"] > 0).sum(1).astype(np.float64)
        scores.extend(float(x) for x in s)
        indices.extend(data["index"])
        total_entropy += float((s * ntok).sum())
        total_tokens += float(ntok.sum())
    order = np.argsort(np.asarray(indices), kind="stable")
    return [scores[i] for i in order], float(np.exp(total_entropy / max(total_tokens, 1.0)))


def in_corpus_order(items, indices):
    """Batches are length-sorted; results go back to corpus order by sample index (evalu.py:256-257)."""
    return [x for _, x in sorted(zip(indices, items), key=lambda p: p[0])]


def eval_metric(trans, references, indices=None):
    """BLEU of `trans` against `references` (a list of reference corpora, each a list of token lists)
    (evalu.py:249-266; the reference reads them from tgt_dev_file / tgt_dev_file0..N)."""
    if not references:
        return 0.0
    if indices is not None:
        trans = in_corpus_order(trans, indices)
    return bleu(trans, list(zip(*references)))


# ---------------------------------------------------------------------------------------------- BLEU
def _ngrams(sentence, n):
    c = Counter()
    for k in range(1, n + 1):
        for i in range(len(sentence) - k + 1):
            c[tuple(sentence[i:i + k])] += 1
    return c


def _closest_length(ref_lengths, cand_length):
    """'closest' brevity-penalty reference length; ties go to the shorter reference (utils/metric.py:67-87)."""
    best, best_d = 9999, 9999
    for r in ref_lengths:
        d = abs(r - cand_length)
        if d < best_d or (d == best_d and r < best):
            best, best_d = r, d
    return best


def bleu(cand, refs, bp="closest", smooth=False, n=4, weights=None):
    """Corpus BLEU-n in [0, 1] (utils/metric.py:243-297).  cand: list of token lists; refs: per sentence a
    tuple/list of reference token lists."""
    len_c = len_ref = 0
    total = Counter()   # candidate n-grams per order
    match = Counter()   # clipped matches per order
    for candidate, references in zip(cand, refs):
        len_c += len(candidate)
        rl = [len(r) for r in references]
        len_ref += _closest_length(rl, len(candidate)) if bp == "closest" else min(rl)
        cg = _ngrams(candidate, n)
        rgs = [_ngrams(r, n) for r in references]
        for g, cnt in cg.items():
            total[len(g)] += cnt
            match[len(g)] += max(min(cnt, rg.get(g, 0)) for rg in rgs)
    if len_ref == 0:
        return 0.0
    precisions = []
    for k in range(1, n + 1):
        if total.get(k, 0) == 0:
            precisions.append(0.0)   # no candidate n-gram of this order: the reference's defaultdict(int) gives 0
            continue
        m, t = match[k], total[k]
        if smooth and k > 1:
            m, t = m + 1, t + 1
        precisions.append(m / t)
    lp = 1.0
    if len_c <= len_ref:
        lp = math.exp(1.0 - len_ref / len_c) if len_c > 0 else 0.0
    if weights is None:
        weights = [1.0 / n] * n
    assert len(weights) == n
    logsum = 0.0
    for p, w in zip(precisions, weights):
        logsum += (math.log(p) if p > 0 else -9999999999.0) * w   # utils/metric.py _safe_log
    return lp * math.exp(logsum)


def dump_tanslation(tranes, output, indices=None):
    """evalu.dump_tanslation (evalu.py:269-280, the reference's spelling): one hypothesis (token list) or score per
    line, restored to corpus order when `indices` are given."""
    if indices is not None:
        tranes = [d[1] for d in sorted(zip(indices, tranes), key=lambda x: x[0])]
    with open(output, "w") as writer:
        for hypo in tranes:
            writer.write((" ".join(hypo) if isinstance(hypo, list) else str(hypo)) + "\n")


def timer():
    return time.time()
