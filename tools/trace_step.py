"""In-situ kernel timeline of the training step (CUPTI through torch.profiler; works inside CUDA-graph replays).

ncu serialises kernels and flushes caches, so its per-kernel times over-state the small kernels of the step.
This tool records the real start / end of every kernel while the step runs exactly as bench.py runs it, and prints
  * per-kernel totals (launches per step, mean duration, share of the step)
  * how much of the step the GPU has >= 1 kernel running, and the idle gaps
Usage: python tools/trace_step.py [--steps 3] [--csv gpurun_out/step_timeline.csv]
"""
import argparse
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--csv", default=None)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--decode", action="store_true",
                    help="trace the beam-4 cached decode of BASELINE configs[2] (one batch) instead of the training step")
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile

    import bench
    from zero_b200.engine import Engine
    from zero_b200.params import transformer_base
    from zero_b200.train import Trainer

    if args.decode:
        from zero_b200 import search
        from zero_b200.params import SimpleVocab
        hp = transformer_base(model_name="transformer_aan", scope_name="transformer_aan", use_ffn=False, aan_mask=True,
                              beam_size=4, decode_length=0, decode_alpha=0.6)
        hp.add_hparam("decode_graph", not args.no_graph)
        hp.add_hparam("src_vocab", SimpleVocab(bench.VOCAB))
        hp.add_hparam("tgt_vocab", SimpleVocab(bench.VOCAB))
        eng = Engine(hp, bench.VOCAB, bench.VOCAB, device="cuda:0")
        eng.ps.init_random(7)
        eng.decode_length = 0
        for i in range(3):
            search.beam_search({"source": bench.make_batch(500 + i, 64)[0]}, eng.encoding_fn, eng.decoding_fn, hp)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            out = search.beam_search({"source": bench.make_batch(503, 64)[0]}, eng.encoding_fn, eng.decoding_fn, hp)
            torch.cuda.synchronize()
        args.steps = int(out["seq"].shape[-1])     # per-"step" figures below are per decode step
    else:
        hp = transformer_base()
        eng = Engine(hp, bench.VOCAB, bench.VOCAB, device="cuda:0")
        eng.ps.init_random(1234)
        trainer = Trainer(eng, hp, world_size=1, use_graph=not args.no_graph)
        batches = [tuple(t.cuda() for t in bench.make_batch(i, bench.B_PER_GPU)) for i in range(4)]
        for i in range(6):
            trainer.step(*batches[i % 4])
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(args.steps):
                trainer.step(*batches[i % 4])
            torch.cuda.synchronize()
    evs = []
    for e in prof.events():
        if str(e.device_type).endswith("CUDA"):
            tr = e.time_range
            if tr.end > tr.start and not e.name.lower().startswith("memcpy"):
                evs.append((float(tr.start), float(tr.end), e.name))
    evs.sort()
    if not evs:
        print("no CUDA events recorded")
        return
    t0, t1 = evs[0][0], max(e[1] for e in evs)
    span = t1 - t0
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s, e, n in evs:
        n = re.sub(r"\(.*", "", n).replace("void ", "")[:60]
        agg[n][0] += 1
        agg[n][1] += e - s
    busy = 0.0
    cur_s, cur_e = evs[0][0], evs[0][1]
    gaps = []
    for s, e, _ in evs[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append(s - cur_e)
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    print("steps %d  span %.1f us (%.1f us / step)  GPU busy %.1f%%  sum of kernel durations %.1f us / step" % (
        args.steps, span, span / args.steps, 100.0 * busy / span, sum(v[1] for v in agg.values()) / args.steps))
    gaps.sort(reverse=True)
    print("idle gaps: n=%d total %.1f us / step, largest %s" % (
        len(gaps), sum(gaps) / args.steps, ["%.1f" % g for g in gaps[:8]]))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
        print("%-60s n/step=%6.1f avg %7.2f us  total/step %8.1f us  %5.1f%%" % (
            n, c / args.steps, t / c, t / args.steps, 100.0 * t / span))
    if args.csv:
        os.makedirs(os.path.dirname(args.csv) or ".", exist_ok=True)
        with open(args.csv, "w") as f:
            f.write("start_us,dur_us,name\n")
            for s, e, n in evs:
                f.write("%.3f,%.3f,%s\n" % (s - t0, e - s, re.sub(r"\(.*", "", n).replace("void ", "").replace(",", ";")[:80]))


if __name__ == "__main__":
    main()
