#!/usr/bin/env python
"""bench.py — train tokens/sec of Zero's Transformer-base (BASELINE.json configs[1]) on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                      (the reference's CPU path: oracle port, host cores)

One "step" = one full optimizer step of the reference's hot loop (main.py:312): forward + backward of the 6+6
d=512 model on this rank's batch of 64 x 64 source / 64 x 64 target tokens (4096 target tokens), one NCCL
all-reduce of the flat fp32 gradient arena, global norms, TF-semantics Adam, bf16 weight refresh.
Metric = non-pad target tokens per second (main.py:297 / :335-346), whole job.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, SRC_LEN, TGT_LEN, VOCAB = 64, 64, 64, 32000
METRIC = "train_tokens_per_sec"
UNIT = "target tokens/s"


def workload_config(n):
    return {"workload": "Transformer-base 6+6 d_model=512 h=8 f=2048 vocab=32k tied softmax, label_smooth=0.1, "
                        "dropout=0, batch 64x(src 64, tgt 64)=4096 target tokens per GPU per step, "
                        "fwd+bwd+allreduce+Adam (BASELINE.json configs[1])",
            "global_batch_tokens": 4096 * n, "src_len": SRC_LEN, "tgt_len": TGT_LEN,
            "parallelism": "dp%d" % n, "l2": "working set (1.2 GB of weights / optimizer state + 0.26 GB of "
                                             "bf16 d_logits + 0.4 GB of activations per step) exceeds the 126 MB L2; "
                                             "no explicit flush"}


def make_batch(seed, batch):
    import torch
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(3, VOCAB, (batch, SRC_LEN), generator=g, dtype=torch.int32)
    tgt = torch.randint(3, VOCAB, (batch, TGT_LEN), generator=g, dtype=torch.int32)
    src[:, -1] = 2
    tgt[:, -1] = 2
    return src, tgt


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def _nvml(self):
        """NVML in-process (sub-millisecond per sample); None when the binding or the driver call is unavailable."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = [(getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "Active"),
                    (getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "Active"),
                    (getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "Active"),
                    (getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4), "Active")]
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons

            def sample():
                r = int(get_reasons(h))
                return [str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx),
                        "%.1f" % (nv.nvmlDeviceGetPowerUsage(h) / 1000.0)] + \
                       ["Active" if r & b else "Not Active" for b, _ in bits]
            sample()
            return sample
        except Exception:
            return None

    def run(self):
        sample = self._nvml()
        self.source = "nvml" if sample else "nvidia-smi"
        while not self._halt.is_set():
            try:
                if sample:
                    self.rows.append(sample())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                    self.rows.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.01 if sample else 0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nm in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows), "source": getattr(self, "source", None)}


# ---------------------------------------------------------------------------------------------- CPU oracle leg
def cpu_reference_step_fn(sample_batch):
    """The reference's CPU path = oracle port (TF1.x cannot run here): fwd + bwd (autograd) + TF-Adam, fp32."""
    import torch
    from oracle import zero_oracle as zo
    from zero_b200.params import transformer_base
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = transformer_base()
    c = zo.Cfg(hp, VOCAB, VOCAB)
    P = {k: v.requires_grad_(True) for k, v in zo.init_params(c, seed=1).items()}
    M = {k: torch.zeros_like(v) for k, v in P.items()}
    Vv = {k: torch.zeros_like(v) for k, v in P.items()}
    state = {"t": 0}

    def step(seed):
        src, tgt = make_batch(seed, sample_batch)
        loss, _, _, _ = zo.train_loss(c, P, src.long(), tgt.long())
        grads = torch.autograd.grad(loss, list(P.values()))
        state["t"] += 1
        with torch.no_grad():
            for (k, p), g in zip(P.items(), grads):
                newp, M[k], Vv[k] = zo.adam_tf_step(p, M[k], Vv[k], g, state["t"], 1e-4, 0.9, 0.98, 1e-8)
                p.copy_(newp)
        return float(loss.detach()), int((tgt != 0).sum())

    return step, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the benchmarked batch itself (64 sentences = 4096 target tokens per optimizer step): the Adam / embedding-gradient
    # cost of the 77 M parameters is amortised over the same token count as on the GPU arm.  A step takes seconds on
    # the host cores, so the step count is capped to keep the run within a few minutes.
    sample_batch = B_PER_GPU
    args.steps = max(1, min(args.steps, 8))
    args.warmup = max(1, min(args.warmup, 2))
    step, cores = cpu_reference_step_fn(sample_batch)
    for i in range(args.warmup):
        step(1000 + i)
    t0 = time.perf_counter()
    toks = 0
    for i in range(args.steps):
        toks += step(i)[1]
    dt = time.perf_counter() - t0
    val = toks / dt
    sample = "%d optimizer steps of the full per-GPU batch: %d sentences x (src 64, tgt 64) = %d target tokens per " \
             "step, fp32, torch CPU, %d threads" % (args.steps, sample_batch, sample_batch * TGT_LEN, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def decode_bench(device, batches=3):
    """Secondary metric of BASELINE.json: beam-4 decode tok/s on configs[2] (transformer_aan 6+6, B=64, len 64,
    cached decode).  Tokens = top-1 hypothesis lengths incl. EOS (evalu.py:25-46) / time of beam_search per batch
    (evalu.py:106-120); also reported as decoder row-steps per second (rows = batch * beam)."""
    import torch
    from zero_b200 import search
    from zero_b200.engine import Engine
    from zero_b200.params import SimpleVocab, transformer_base
    hp = transformer_base(model_name="transformer_aan", scope_name="transformer_aan", use_ffn=False, aan_mask=True,
                          beam_size=4, decode_length=0, decode_alpha=0.6)
    hp.add_hparam("src_vocab", SimpleVocab(VOCAB))
    hp.add_hparam("tgt_vocab", SimpleVocab(VOCAB))
    hp.add_hparam("decode_graph", os.environ.get("ZB_DECODE_GRAPH", "1") != "0")   # 0: eager steps (for ncu)
    eng = Engine(hp, VOCAB, VOCAB, device=device)
    eng.ps.init_random(7)
    eng.decode_length = 0
    tot_tok, tot_steps, tot_ms = 0, 0, 0.0
    for i in range(batches + 2):
        src, _ = make_batch(500 + i, 64)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        # the engine's own decoding_fn: beam_search replays one CUDA graph per step index
        out = search.beam_search({"source": src}, eng.encoding_fn, eng.decoding_fn, hp)
        steps = {"n": int(out["seq"].shape[-1])}
        e1.record()
        torch.cuda.synchronize()
        if i < 2:
            continue  # batch 0 allocates the workspace, batch 1 captures the per-step CUDA graphs
        top1 = out["seq"][:, 0, :].cpu()
        lens = []
        for row in top1.tolist():
            n = 0
            for tok in row:
                n += 1
                if tok == 2 or tok == 0:
                    break
            lens.append(n)
        tot_tok += sum(lens)
        tot_steps += steps["n"]
        tot_ms += e0.elapsed_time(e1)
    # HBM roofline of one decode step: decoder weights 6 x (gate 1.05 M + cross q/o 0.52 M + FFN 2.10 M)
    # params + the tied vocabulary table 16.4 M, bf16, + the projected memories [64, 64, 1024] x 6 layers read once per
    # sentence.  The step logits are not algorithmic traffic: K8 fused reduces them to beam candidates in the GEMM
    # epilogue; only with ZB_BEAM_FUSED=0 are the fp32 [256, 32000] logits written and read back by the beam step.
    logits_path = os.environ.get("ZB_BEAM_FUSED", "1") == "0"
    step_bytes = 2.0 * (6 * (1024 * 1024 + 2 * 512 * 512 + 2 * 512 * 2048) + VOCAB * 512) \
        + 2.0 * 64 * 64 * 1024 * 6 + (2 * 4.0 * 256 * VOCAB if logits_path else 0.0)
    ms_step = tot_ms / max(tot_steps, 1)
    peak = None
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    roof = {"bound": "hbm", "algorithmic_bytes_per_step": step_bytes, "achieved": step_bytes / (ms_step * 1e-3) / 1e9,
            "peak": peak, "unit": "GB/s", "frac": (step_bytes / (ms_step * 1e-3) / 1e9 / peak) if peak else None,
            "note": "the step is ~75 dependent launches of 256-row problems: latency-bound, not bandwidth-bound",
            "vocabulary_projection": "logits + beam kernels" if logits_path else "beam candidates from the GEMM epilogue"}
    return {"metric": "beam4_decode_tokens_per_sec", "value": tot_tok / (tot_ms * 1e-3), "unit": "top-1 tokens/s",
            "roofline": roof,
            "row_steps_per_sec": tot_steps * 256 / (tot_ms * 1e-3), "steps": tot_steps, "ms_per_step": tot_ms / max(tot_steps, 1),
            "config": "transformer_aan 6+6 d=512, beam 4, batch 64, src len 64, max target len 64 (BASELINE configs[2])"}


# ---------------------------------------------------------------------------------------------- secondary legs
def model_flops_per_step(hp, B, S, T, V, rpr_k=0):
    """Algorithmic fwd + bwd FLOPs of one step (multiply-add = 2, backward = 2 x forward, attention
    recompute not counted).  Encoder layer per source token 8d^2 + 4df + 4Sd; decoder layer per target token
    8d^2 + 4Td (self) + 4d^2 + 4Sd (cross q / o + logits / context) + 4df, plus the memory projection 4d^2 per SOURCE
    token; vocabulary projection 2dV per target token; relative positions as bucket GEMMs 2 (2k + 1) d per token and
    attention, keys and values."""
    d, f = int(hp.hidden_size), int(hp.filter_size)
    ne, nd = int(hp.num_encoder_layer), int(hp.num_decoder_layer)
    rp = 2 * 2 * (2 * rpr_k + 1) * d if rpr_k else 0
    enc = B * S * ne * (8 * d * d + 4 * d * f + 4 * S * d + rp)
    dec = B * T * nd * (8 * d * d + 4 * T * d + 4 * d * d + 4 * S * d + 4 * d * f + 2 * rp) + B * S * nd * 4 * d * d
    return 3.0 * (enc + dec + B * T * 2 * d * V)


def config_leg(name, hp, B, S, T, steps, peaks, rpr_k=0):
    """Training throughput of another BASELINE config on this GPU: same Trainer / CUDA-graph path as the headline leg,
    device-resident batches, CUDA-event timing.  Reports tokens/s, ms/step and the step-level fraction of the measured
    sustained bf16 peak (whole-model FLOPs / time: everything in the step counts as time, only the model's algorithmic
    FLOPs count as work)."""
    import torch
    from zero_b200.engine import Engine
    from zero_b200.train import Trainer
    eng = Engine(hp, VOCAB, VOCAB, device="cuda")
    eng.ps.init_random(1234)
    trainer = Trainer(eng, hp, world_size=1, use_graph=True)
    g = torch.Generator().manual_seed(99)
    batches = []
    for i in range(4):
        src = torch.randint(3, VOCAB, (B, S), generator=g, dtype=torch.int32)
        tgt = torch.randint(3, VOCAB, (B, T), generator=g, dtype=torch.int32)
        src[:, -1] = 2
        tgt[:, -1] = 2
        batches.append((src.cuda(), tgt.cuda()))
    for i in range(3):
        trainer.step(*batches[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = trainer.step(*batches[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flops = model_flops_per_step(hp, B, S, T, VOCAB, rpr_k)
    tf = flops / (ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    out = {"config": name, "metric": METRIC, "value": B * T / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "batch": B, "src_len": S, "tgt_len": T, "final_loss": float(loss.item()),
           "model_tflop_per_step": flops / 1e12,
           "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                        "note": "step-level: whole-model algorithmic FLOPs over the whole step time"}}
    del trainer, eng
    torch.cuda.empty_cache()
    return out


def layer_peak_leg(peaks, iters=5):
    """The north-star kernel shape: ONE Transformer encoder layer (models/transformer.py:45-69: self-attention ->
    residual + LN -> FFN -> residual + LN) forward + backward on [4096, 64, 512] bf16 = 262 144 tokens, 5.051 TFLOP
    algorithmic, timed with CUDA events; target >= 40 % of the bf16 tensor peak."""
    import torch
    from zero_b200.engine import Engine
    from zero_b200.params import transformer_base
    hp = transformer_base(num_encoder_layer=1, num_decoder_layer=1)
    eng = Engine(hp, 1024, 1024)
    eng.ps.init_random(1)
    eng.enable_side_stream(True)
    c = eng.cfg
    B, S = 4096, 64
    N = B * S
    x = torch.randn(N, c.d, device="cuda").to(torch.bfloat16)
    d_out = (torch.randn(N, c.d, device="cuda") * 0.01).to(torch.bfloat16)
    src_len = torch.full((B,), S, dtype=torch.int32, device="cuda")

    def step():
        sv = {"att": {}, "ln1": {}, "ffn": {}, "ln2": {}}
        y = eng._self_attn_fwd("enc0.self", x, B, S, src_len, False, sv["att"], "L.att")
        x1 = eng._ln_fwd("enc0.self.ln", x, y, N, sv["ln1"], "L.ln1")
        y2 = eng._ffn_fwd("enc0.ffn", x1, N, sv["ffn"], "L.ffn")
        eng._ln_fwd("enc0.ffn.ln", x1, y2, N, sv["ln2"], "L.ln2")
        ds2, dy2 = eng._ln_bwd("enc0.ffn.ln", d_out, None, N, sv["ln2"], "L.bw.ln2", eng.ps.g("enc0.ffn.w2.b"))
        dx1 = eng._ffn_bwd("enc0.ffn", x1, dy2, N, sv["ffn"], "L.bw.ffn")
        ds1, dy1 = eng._ln_bwd("enc0.self.ln", ds2, dx1, N, sv["ln1"], "L.bw.ln1", eng.ps.g("enc0.self.o.b"))
        eng._self_attn_bwd("enc0.self", x, dy1, B, S, sv["att"], "L.bw.att")
        eng._side_join()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 19267584.0 * N
    tf = flops / (ms * 1e-3) / 1e12
    burst = float(peaks.get("bf16_tflops", 1590.0))
    out = {"shape": [B, S, c.d], "tokens": N, "ms_fwd_bwd": ms, "algorithmic_tflop": flops / 1e12,
           "roofline": {"bound": "tensor", "achieved": tf, "peak": burst, "unit": "TFLOP/s", "frac": tf / burst,
                        "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if "bf16_tflops" in peaks else "fallback"},
           "frac_of_nominal_2250": tf / 2250.0,
           "frac_of_measured_sustained": tf / float(peaks.get("bf16_tflops_sustained", 1400.0)),
           "target": "north_star: >= 0.40 of the sm_100a tensor-pipe peak"}
    del eng
    torch.cuda.empty_cache()
    return out


def variable_shape_leg(steps=60):
    """The reference's REAL usage (main.py:242-312): length-sorted, token-budget batches from the data pipeline — a new
    (B, S, T) almost every step — through the same public step with HOST (pinned) batches.  CUDA graphs are replayed
    per shape class (S / T zero-padded to multiples of 8, ZB_GRAPH_BUCKET) over the shared workspace; reports tokens/s
    once every class of the corpus has been captured (second epoch), the number of classes and the eager rate (no
    graphs) for comparison."""
    import numpy as np
    import torch
    from zero_b200.data import pin
    from zero_b200.engine import Engine
    from zero_b200.params import transformer_base
    from zero_b200.train import Trainer
    rng = np.random.default_rng(7)
    src, tgt = [], []
    for _ in range(6000):             # only the lengths matter here; ids are drawn when the matrices are built
        src.append([0] * int(rng.integers(8, 64)))
        tgt.append([0] * int(rng.integers(8, 64)))
    out = {}
    for mode in ("graphs", "eager"):
        os.environ["ZB_GRAPH_BUCKET"] = "8" if mode == "graphs" else "0"
        hp = transformer_base()
        eng = Engine(hp, VOCAB, VOCAB, device="cuda")
        eng.ps.init_random(1234)
        trainer = Trainer(eng, hp, world_size=1, use_graph=(mode == "graphs"))
        g = np.random.default_rng(3)
        batches = []
        # token-budget batches of ~4096 target tokens from length-sorted buckets (utils/util.py:30-65)
        order = np.argsort([max(len(a), len(b)) for a, b in zip(src, tgt)], kind="stable")
        cur, width = [], 0
        for i in order:
            w = max(len(src[i]), len(tgt[i])) + 1
            if cur and (len(cur) + 1) * max(width, w) > 4096:
                batches.append(cur)
                cur, width = [], 0
            cur.append(i)
            width = max(width, w)
        if cur:
            batches.append(cur)
        g.shuffle(batches)

        def matrix(idx, corpus):
            width = max(len(corpus[i]) for i in idx) + 1
            m = np.zeros((len(idx), width), dtype=np.int32)
            for r, i in enumerate(idx):
                ids = g.integers(3, VOCAB, len(corpus[i]))
                m[r, :len(ids)] = ids
                m[r, len(ids)] = 2
            return m
        host = [pin({"src": matrix(b, src), "tgt": matrix(b, tgt)}) for b in batches]
        for s, t in host:                      # epoch 1: allocate the workspace, capture every shape class
            trainer.step(s, t)
        torch.cuda.synchronize()
        if mode == "graphs":
            for s, t in host:
                trainer.step(s, t)
            torch.cuda.synchronize()
        n = min(steps, len(host))
        toks = sum(int((t != 0).sum()) for _, t in host[:n])
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s, t in host[:n]:
            loss = trainer.step(s, t)
        _ = loss.item()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[mode] = {"value": toks / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / n, "steps": n,
                     "host_wall_ms_per_step": 1000 * (time.perf_counter() - t0) / n,
                     "mean_target_tokens_per_step": toks / n, "distinct_batches": len(host),
                     "captured_shape_classes": len(trainer._graphs) if mode == "graphs" else 0}
        del trainer, eng
        torch.cuda.empty_cache()
    os.environ.pop("ZB_GRAPH_BUCKET", None)
    out["workload"] = ("6000 synthetic pairs, lengths ~U[8,64), length-sorted token-budget batches of <= 4096 padded "
                       "target tokens (main.py:242-312), pinned host ids in, configs[1] model")
    return out


def extra_legs(peaks):
    """BASELINE configs[3] and [4] as training-throughput legs + the encoder-layer peak shape (N = 1 only)."""
    from zero_b200.params import transformer_base
    legs = {}
    try:
        legs["layer_peak"] = layer_peak_leg(peaks)
    except Exception as e:      # a secondary leg must not take the headline line down with it
        legs["layer_peak"] = {"error": repr(e)[:300]}
    try:
        hp = transformer_base(model_name="transformer_rpr", scope_name="transformer_rpr", max_relative_position=16)
        legs["configs3_rpr_len128"] = config_leg(
            "transformer_rpr 6+6 d=512, src / tgt len 128, max_relative_position 16, 32 sentences = 4096 target tokens "
            "per GPU per step (BASELINE configs[3])", hp, 32, 128, 128, 20, peaks, rpr_k=16)
    except Exception as e:
        legs["configs3_rpr_len128"] = {"error": repr(e)[:300]}
    try:
        legs["variable_shape_token_batches"] = variable_shape_leg()
    except Exception as e:
        legs["variable_shape_token_batches"] = {"error": repr(e)[:300]}
    try:
        hp = transformer_base(num_encoder_layer=24, num_decoder_layer=6, deep_transformer_init=True,
                              initializer="uniform_unit_scaling", initializer_gain=1.0)
        legs["configs4_deep_len1024"] = config_leg(
            "24-layer DS-Init encoder + 6-layer decoder d=512, src len 1024 (token ids in place of the speech "
            "front-end, which is not in the reference checkout), tgt len 64, 8 sentences per GPU per step "
            "(BASELINE configs[4])", hp, 8, 1024, 64, 10, peaks)
    except Exception as e:
        legs["configs4_deep_len1024"] = {"error": repr(e)[:300]}
    return legs


# ---------------------------------------------------------------------------------------------- GPU arm
def gemm_roofline(eng, src, tgt, peaks):
    """Dominant kernel = zb_gemm (tcgen05, 201 launches per step).  The GEMM launches of one real step are recorded
    (same operands, same order), re-issued back to back from a CUDA graph — so no host launch gaps pollute the
    number — and timed with CUDA events on the launching stream: achieved = algorithmic GEMM FLOPs / that time."""
    import torch
    from zero_b200 import ops
    calls, flops, nsingle = [], [0.0], [0]
    real = ops.gemm

    def recording(a, b, out, a_layout=0, b_layout=1, **kw):
        M = kw.get("m") or (a.shape[0] if a_layout == 0 else a.shape[1])
        K = kw.get("k") or (a.shape[1] if a_layout == 0 else a.shape[0])
        N = kw.get("n") or (b.shape[0] if b_layout == 0 else b.shape[1])
        flops[0] += 2.0 * M * N * K
        nsingle[0] += 1
        calls.append(lambda: real(a, b, out, a_layout, b_layout, **kw))
        return real(a, b, out, a_layout, b_layout, **kw)

    real_grouped = ops.gemm_grouped

    def recording_grouped(problems):
        for g in problems:
            flops[0] += 2.0 * g.m * g.n * g.k
        calls.append(lambda: real_grouped(problems))
        nprob[0] += len(problems)
        return real_grouped(problems)

    nprob = [0]
    real_vce = ops.vocab_ce

    def recording_vce(feat, table, labels, nll, smooth, workspace, **kw):
        # K6: one algorithmic logits GEMM (the recomputation inside the fused op is its own cost, not extra work)
        flops[0] += 2.0 * feat.shape[0] * table.shape[0] * feat.shape[1]
        nsingle[0] += 1
        calls.append(lambda: real_vce(feat, table, labels, nll, smooth, workspace, **kw))
        return real_vce(feat, table, labels, nll, smooth, workspace, **kw)

    from zero_b200 import lib as L
    side = getattr(eng, "side", None)
    eng.side = None  # record the launches in program order on one stream
    ops.gemm = recording
    ops.gemm_grouped = recording_grouped
    ops.vocab_ce = recording_vce
    c0 = L.launch_count()
    try:
        eng.forward_backward(src, tgt, compact=False)
        torch.cuda.synchronize()
    finally:
        step_launches = L.launch_count() - c0  # every kernel of one eager fwd+bwd (graph replays bypass the counter)
        ops.gemm = real
        ops.gemm_grouped = real_grouped
        ops.vocab_ce = real_vce
        eng.side = side
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for c in calls:
            c()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    achieved = flops[0] / (ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    traffic, traffic_src = None, None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "gemm_dram_traffic.json")))
        traffic, traffic_src = float(t["dram_bytes_per_launch"]), t["source"]
    except Exception:
        pass
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src,
            "kernel": "gemm2_bf16_tcgen05", "launches_per_step": len(calls),
            "gemm_problems_per_step": nsingle[0] + nprob[0],
            "step_launches": step_launches,
            "gemm_ms_per_step": ms, "avg_launch_us": 1000.0 * ms / max(len(calls), 1),
            "gemm_flops_per_step": flops[0],
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
            if "bf16_tflops_sustained" in peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from zero_b200 import lib as L
    from zero_b200.engine import Engine
    from zero_b200.params import transformer_base
    from zero_b200.train import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hp = transformer_base()
    eng = Engine(hp, VOCAB, VOCAB, device="cuda:%d" % local)
    eng.ps.init_random(1234)   # identical replicas on every rank (run.py:379-381 seeds)
    trainer = Trainer(eng, hp, world_size=world, use_graph=not args.no_graph)

    # host-resident (pinned) batches for the e2e leg; device-resident copies for the kernel-side `value`
    n_batches = max(args.steps, 1)
    host = [tuple(t.pin_memory() for t in make_batch(rank * 100003 + i, B_PER_GPU)) for i in range(n_batches)]
    devb = [(s.cuda(non_blocking=True), t.cuda(non_blocking=True)) for s, t in host]
    tokens_per_step = int(sum(int((t != 0).sum()) for _, t in host) / n_batches)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        trainer.step(*devb[i % n_batches])
    barrier()

    # ---- leg 1: inputs resident in HBM
    launches0 = L.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        loss = trainer.step(*devb[i % n_batches])
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    eager_launches = L.launch_count() - launches0
    final_loss = float(loss.item())

    # ---- leg 2: end to end through the public step with HOST buffers (pinned H2D of ids + D2H of the loss)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        loss = trainer.step(*host[i % n_batches])
        _ = loss.item()
    t1.record()
    barrier()
    ms2 = torch.tensor([t0.elapsed_time(t1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms2 = float(ms2.item())
    clocks = sampler.stop() if sampler else None   # sampled across both timed regions

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # kernel launches per step: graph replays do not pass through the ABI counter, so count one eager step
    roof = gemm_roofline(eng, *devb[0], peaks)
    per_step_kernels = roof.pop("step_launches") + 1  # + the fused Adam/norm kernel outside forward_backward
    value = tokens_per_step * world * args.steps / (ms * 1e-3)
    e2e = tokens_per_step * world * args.steps / (ms2 * 1e-3)
    model_flops = 369623040.0 * tokens_per_step  # fwd+bwd FLOPs per (src,tgt) token pair
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        sample_batch = B_PER_GPU     # the benchmarked batch (same_config): ~2-4 s per step on the host cores
        stepf, cores = cpu_reference_step_fn(sample_batch)
        stepf(999)
        t_0 = time.perf_counter()
        toks = 0
        nrep = 0
        while time.perf_counter() - t_0 < 15.0 and nrep < 5:
            toks += stepf(nrep)[1]
            nrep += 1
        dt = time.perf_counter() - t_0
        cpu_baseline = {"value": toks / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "%d optimizer steps of %d sentences x (src 64, tgt 64) through the oracle port "
                                  "(torch CPU fp32, fwd+bwd+Adam), same model config" % (nrep, sample_batch)}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": dict(workload_config(world), gradient_step=(
            "NCCL all-reduce of the fp32 gradient arena (two buckets) + replicated Adam" if trainer.shard is None else
            "fused reduce-scatter + sharded Adam + bf16 all-gather in one kernel per rank over %s (ZB_SHARD_OPT)" % (
                "NVSwitch multicast" if trainer.shard.grad_mc else "NVLink peer memory")) if world > 1 else "Adam"),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 2 * B_PER_GPU * 64 * 4,
                "d2h_bytes_per_step": 4, "ms_per_step": ms2 / args.steps},
        "gpu_launches": int(per_step_kernels * args.steps),
        "kernels_per_step": int(per_step_kernels),
        "abi_calls_in_timed_region": int(eager_launches),
        "cuda_graph": not args.no_graph,
        "roofline": roof,
        "model_tflops_per_gpu": model_flops / (ms / args.steps * 1e-3) / 1e12,
        "cpu_baseline": cpu_baseline,
        "decode": None,
        "legs": None,
        "clocks": clocks,
        "final_loss": final_loss,
    }
    if world == 1 and not args.no_decode:
        try:
            out["decode"] = decode_bench("cuda:%d" % local)
        except Exception as e:      # a secondary leg must not take the headline line down with it
            out["decode"] = {"error": repr(e)[:300]}
    if world == 1 and not args.no_extra:
        del trainer, eng
        torch.cuda.empty_cache()
        out["legs"] = extra_legs(peaks)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[3] / configs[4] / layer-peak legs")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else min(args.steps, 50)
        args.warmup = 1 if args.warmup is None else min(args.warmup, 3)
        run_reference(args)
    else:
        args.steps = 100 if args.steps is None else args.steps
        args.warmup = 10 if args.warmup is None else max(args.warmup, 3)   # timing rule: never fewer than 3 warm-up steps
        run_ours(args)


if __name__ == "__main__":
    main()
