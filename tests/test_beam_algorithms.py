"""CPU checks of the selection algorithms the beam-step kernels rely on (csrc/beam.cu), restated in numpy:
the row / part decompositions and the threshold pass must return exactly tf.nn.top_k's answer over the flat
beam * V candidates (search.py:175-176: descending, ties -> lower flat index), including on tie-heavy inputs.
These are statements about the algorithm, not about the CUDA code (that is tests/test_kernels_gpu.py)."""
import numpy as np
import pytest


def exact_topk(score, n):
    """(values, flat indices) of the n best entries: value descending, ties -> lower index (tf.nn.top_k)."""
    idx = np.lexsort((np.arange(score.size), -score.astype(np.float64)))[:n]
    return score[idx], idx


def row_lists_then_merge(score, beam, vocab, n2):
    """beam_row_kernel: every (sentence, beam) row keeps its own exact top-2k; the last arriver merges the lists."""
    cand_s, cand_i = [], []
    for k in range(beam):
        s, i = exact_topk(score[k * vocab:(k + 1) * vocab], min(n2, vocab))
        cand_s.append(s)
        cand_i.append(i + k * vocab)
    cs, ci = np.concatenate(cand_s), np.concatenate(cand_i)
    order = np.lexsort((ci, -cs.astype(np.float64)))[:n2]
    return cs[order], ci[order]


def part_lists_with_threshold(score, beam, vocab, n2, parts=4, threads=256):
    """beam_part_kernel: a row is cut into `parts` ranges; inside a part each of `threads` threads owns elements
    tid, tid + threads, ...; T = the n2-th largest of the thread maxima; only elements with score >= T enter the
    sorted lists; part lists -> row list -> sentence list."""
    vp = (((vocab + parts - 1) // parts) + 3) & ~3
    cand_s, cand_i = [], []
    passed = 0
    for k in range(beam):
        row = score[k * vocab:(k + 1) * vocab]
        for p in range(parts):
            lo, hi = p * vp, min(vocab, (p + 1) * vp)
            if hi <= lo:
                continue
            part = row[lo:hi]
            tmax = np.full(threads, -np.inf, dtype=np.float32)
            for tid in range(min(threads, part.size)):
                tmax[tid] = part[tid::threads].max()
            thr = np.sort(tmax)[::-1][n2 - 1] if n2 <= threads else -np.inf
            keep = np.nonzero(part >= thr)[0]
            passed += keep.size
            s, i = exact_topk(part[keep], min(n2, keep.size))
            cand_s.append(s)
            cand_i.append(keep[i] + lo + k * vocab)
    cs, ci = np.concatenate(cand_s), np.concatenate(cand_i)
    order = np.lexsort((ci, -cs.astype(np.float64)))[:n2]
    return cs[order], ci[order], passed


@pytest.mark.parametrize("beam,vocab", [(1, 64), (3, 207), (4, 32000), (5, 1000), (8, 4099), (2, 7)])
@pytest.mark.parametrize("ties", [False, True])
def test_row_and_part_decompositions_equal_flat_topk(beam, vocab, ties):
    rng = np.random.RandomState(beam * 131 + vocab)
    score = (rng.randn(beam * vocab) * 3).astype(np.float32)
    if ties:
        score = np.round(score)          # a few dozen distinct values: ties everywhere, also across rows and parts
    n2 = 2 * beam
    want_s, want_i = exact_topk(score, n2)
    got_s, got_i = row_lists_then_merge(score, beam, vocab, n2)
    np.testing.assert_array_equal(got_i, want_i)
    np.testing.assert_array_equal(got_s, want_s)
    got_s, got_i, passed = part_lists_with_threshold(score, beam, vocab, n2)
    np.testing.assert_array_equal(got_i, want_i)
    np.testing.assert_array_equal(got_s, want_s)
    if not ties and vocab >= 4096:
        # the point of the threshold: a small multiple of 2k elements per part reach the sorted lists,
        # not vocab / parts of them
        assert passed <= beam * 4 * 16 * n2


def test_threshold_keeps_every_tied_candidate():
    """All-equal scores: T equals that value, every element passes `>= T`, and the index order decides."""
    score = np.zeros(4 * 1000, dtype=np.float32)
    s, i, passed = part_lists_with_threshold(score, 4, 1000, 8)
    np.testing.assert_array_equal(i, np.arange(8))
    assert passed == 4000


def test_gumbel_generator_statistics():
    """The counter-based generator behind zb_gumbel_add (csrc/elementwise.cu gumbel_add_kernel; the splitmix64 hash of
    the dropout masks, 24 bits per element), restated in numpy integer arithmetic: Gumbel(0, 1) has mean 0.5772 and
    variance pi^2 / 6 (util.gumbel_noise, utils/util.py:189-195); sites, seeds and neighbours are uncorrelated."""
    import math

    def hash4(seed, site, idx):
        with np.errstate(over="ignore"):
            z = np.uint64(seed) + np.uint64(site) * np.uint64(0x9E3779B97F4A7C15) + idx * np.uint64(0xD1B54A32D192ED03)
            z ^= z >> np.uint64(30)
            z *= np.uint64(0xBF58476D1CE4E5B9)
            z ^= z >> np.uint64(27)
            z *= np.uint64(0x94D049BB133111EB)
            z ^= z >> np.uint64(31)
        return z

    def gumbel(n, seed, site, eps=1e-8):
        h = hash4(seed, site ^ 0x6A09E667, np.arange((n + 1) // 2, dtype=np.uint64))
        u = np.empty(2 * len(h), np.float32)
        u[0::2] = ((h & np.uint64(0xFFFFFFFF)) >> np.uint64(8)).astype(np.float32) * np.float32(1 / 16777216)
        u[1::2] = (h >> np.uint64(40)).astype(np.float32) * np.float32(1 / 16777216)
        u = u[:n]
        assert float(u.min()) >= 0.0 and float(u.max()) < 1.0
        return -np.log(-np.log(u + np.float32(eps)) + np.float32(eps))

    n = 1 << 20
    a, c, d = gumbel(n, 77, 5), gumbel(n, 77, 6), gumbel(n, 78, 5)
    assert abs(float(a.mean()) - 0.5772) < 8e-3 and abs(float(a.var()) - math.pi ** 2 / 6) < 2.5e-2
    assert -3.0 < float(a.min()) and float(a.max()) < 18.5
    assert abs(float((a * c).mean()) - 0.5772 ** 2) < 1.2e-2 and abs(float((a * d).mean()) - 0.5772 ** 2) < 1.2e-2
    assert abs(float(np.corrcoef(a[:-1], a[1:])[0, 1])) < 5e-3
