"""L2 -> SM traffic model of one training step's GEMMs at BASELINE configs[1] (no GPU needed).

A persistent tile kernel streams (BM + BN) * K * 2 bytes from L2 into shared memory per output tile.  The chip's L2
slices deliver ~6300 B/clk in total (B300_MICROARCH.md "LTS throughput cap"; 42.6 B/clk per SM when all 148 pull,
which is the 0.27 us per 64-wide k-block of a 24 KB CTA stage measured in profiles/r01_gemm2_trace_v1_vs_v5.log), so a
tile shape fixes an upper bound on the GEMM rate independent of the tensor pipe:  2 * BM * BN / ((BM + BN) * 2) FLOP/B.

    python tools/gemm_l2_model.py            # the step's 201 GEMM problems under the current tile choice
    python tools/gemm_l2_model.py 256        # ... with 256-wide tiles forced (ZB_GEMM2_BN=256)
    python tools/gemm_l2_model.py l2         # ... with the byte-weighted wave rule (ZB_GEMM2_TILE_MODEL=l2)
"""
import sys

PAIRS, L2_BPS, TENSOR = 74, 6300 * 1.9e9, 1386.6e12   # pairs of SMs, L2 bytes/s at 1.9 GHz, sustained bf16 FLOP/s


def tiles(m, n, bn):
    return -(-m // 256) * -(-n // bn)


def choose_bn(m, n, force=None, accum=False):
    """gemm2_launch's wave-quantisation rule (csrc/gemm2_tcgen05.cu); force = 128 / 256 (ZB_GEMM2_BN) or "l2"
    (ZB_GEMM2_TILE_MODEL=l2: a wave costs its operand bytes, 256 + bn, instead of its flops)."""
    if n <= 128:
        return 128
    if force in (128, 256):
        return force
    if accum:          # weight gradients: split-K fills the machine, the widest tile is kept
        return 256
    c256, c128 = (512, 384) if force == "l2" else (256, 128)
    w256 = -(-tiles(m, n, 256) // PAIRS) * c256
    w128 = -(-tiles(m, n, 128) // PAIRS) * c128
    return 128 if w128 < w256 else 256


def step_problems(T=4096, d=512, f=2048, V=32000):
    enc = [(T, 3 * d, d), (T, d, d), (T, f, d), (T, d, f)]
    dec = [(T, 3 * d, d), (T, d, d), (T, d, d), (T, 2 * d, d), (T, d, d), (T, f, d), (T, d, f)]
    out = []
    for _ in range(6):
        for fw in (enc, dec):
            out += [p + (False,) for p in fw] + [(m, k, n, False) for m, n, k in fw]      # forward, dgrad
            out += [(k, n, m, True) for m, n, k in fw]                                   # wgrad (accumulating)
    return out + [(T, V, d, False), (T, d, V, True), (V, d, T, True)]


def main():
    force = (sys.argv[1] if sys.argv[1] == "l2" else int(sys.argv[1])) if len(sys.argv) > 1 else None
    rows = {}
    for m, n, k, accum in step_problems():
        bn = choose_bn(m, n, force, accum)
        traffic = tiles(m, n, bn) * (256 + bn) * k * 2
        r = rows.setdefault((m, n, k, bn), [0, 0.0, 0.0])
        r[0] += 1
        r[1] += traffic
        r[2] += 2.0 * m * n * k
    tot_b = sum(r[1] for r in rows.values())
    tot_f = sum(r[2] for r in rows.values())
    print("%6s %6s %6s %4s %4s %10s %9s %9s" % ("m", "n", "k", "bn", "x", "L2 MB", "L2 us", "tensor us"))
    for (m, n, k, bn), (cnt, b, fl) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
        print("%6d %6d %6d %4d %4d %10.1f %9.1f %9.1f" % (m, n, k, bn, cnt, b / 1e6, b / L2_BPS * 1e6, fl / TENSOR * 1e6))
    print("step: %d problems, %.2f GB L2 -> SM, %.3f ms at the L2 cap, %.3f ms at sustained tensor peak (%.3f TFLOP)" % (
        sum(r[0] for r in rows.values()), tot_b / 1e9, tot_b / L2_BPS * 1e3, tot_f / TENSOR * 1e3, tot_f / 1e12))


if __name__ == "__main__":
    main()
