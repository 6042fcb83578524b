"""North-star kernel shape: one fused Transformer encoder layer forward + backward at [4096, 64, 512] bf16
(262 144 tokens; SURVEY.md 8(d): 5.051 TFLOP algorithmic, recompute not counted).

    python tools/bench_layer.py [--batch 4096] [--iters 5]

Prints one JSON line: ms per fwd+bwd, achieved TFLOP/s and the fraction of the measured / nominal bf16 peaks."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from zero_b200.engine import Engine  # noqa: E402
from zero_b200.params import transformer_base  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--seq", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--no-side", action="store_true")
    args = ap.parse_args()
    hp = transformer_base(num_encoder_layer=1, num_decoder_layer=1)
    eng = Engine(hp, 1024, 1024)
    eng.ps.init_random(1)
    if not args.no_side:
        eng.enable_side_stream(True)
    c = eng.cfg
    B, S = args.batch, args.seq
    N = B * S
    x = (torch.randn(N, c.d, device="cuda") * 1.0).to(torch.bfloat16)
    d_out = (torch.randn(N, c.d, device="cuda") * 0.01).to(torch.bfloat16)
    src_len = torch.full((B,), S, dtype=torch.int32, device="cuda")

    def step():
        sv = {"att": {}, "ln1": {}, "ffn": {}, "ln2": {}}
        y = eng._self_attn_fwd("enc0.self", x, B, S, src_len, False, sv["att"], "L.att")
        x1 = eng._ln_fwd("enc0.self.ln", x, y, N, sv["ln1"], "L.ln1")
        y2 = eng._ffn_fwd("enc0.ffn", x1, N, sv["ffn"], "L.ffn")
        eng._ln_fwd("enc0.ffn.ln", x1, y2, N, sv["ln2"], "L.ln2")
        ds2 = eng._ln_bwd("enc0.ffn.ln", d_out, None, N, sv["ln2"], "L.bw.ln2", eng.ps.g("enc0.ffn.w2.b"))
        dx1 = eng._ffn_bwd("enc0.ffn", x1, ds2, N, sv["ffn"], "L.bw.ffn")
        ds1 = eng._ln_bwd("enc0.self.ln", ds2, dx1, N, sv["ln1"], "L.bw.ln1", eng.ps.g("enc0.self.o.b"))
        eng._self_attn_bwd("enc0.self", x, ds1, B, S, sv["att"], "L.bw.att")
        eng._side_join()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    flops = 19267584.0 * (S / 64.0 * 0 + 1) * N if S == 64 else (3 * (8 * c.d ** 2 + 4 * c.d * c.f + 4 * S * c.d)) * N
    tf = flops / (ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    print(json.dumps({"shape": [B, S, c.d], "tokens": N, "ms_fwd_bwd": ms, "algorithmic_tflop": flops / 1e12,
                      "achieved_tflops": tf, "frac_of_nominal_2250": tf / 2250.0,
                      "frac_of_measured_burst": tf / peaks.get("bf16_tflops", 1590.0),
                      "frac_of_measured_sustained": tf / peaks.get("bf16_tflops_sustained", 1400.0)}))


if __name__ == "__main__":
    main()
