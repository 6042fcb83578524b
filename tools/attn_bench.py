"""Attention kernels alone: tcgen05 (attention_tc.cu) against mma.sync (ZB_ATTN_TC=0) at the BASELINE shapes, timed
with CUDA events over back-to-back launches (a CUDA graph of 20 launches, so host launch cost is out of the number).
ZB_ATTN_TRACE=1 additionally prints the in-kernel timeline of CTA 0 (first call of every shape).
Usage: python tools/attn_bench.py [--trace]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from zero_b200 import ops  # noqa: E402

dev = torch.device("cuda")
bf16 = torch.bfloat16


def run(B, h, Lq, Lk, causal, fused, reps=20):
    D = h * 64
    g = torch.Generator(device="cuda").manual_seed(1)
    if fused:
        qkv = torch.randn(B, Lq, 3 * D, generator=g, device=dev).to(bf16)
        q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
        dqkv = torch.zeros_like(qkv)
        dq, dk, dv = dqkv[:, :, :D], dqkv[:, :, D:2 * D], dqkv[:, :, 2 * D:]
    else:
        q = torch.randn(B, Lq, D, generator=g, device=dev).to(bf16)
        kv = torch.randn(B, Lk, 2 * D, generator=g, device=dev).to(bf16)
        k, v = kv[:, :, :D], kv[:, :, D:]
        dq = torch.zeros_like(q)
        dkv = torch.zeros_like(kv)
        dk, dv = dkv[:, :, :D], dkv[:, :, D:]
    d_o = torch.randn(B, Lq, D, generator=g, device=dev).to(bf16)
    o = torch.empty(B, Lq, D, dtype=bf16, device=dev)
    lse = torch.empty(B, h, Lq, device=dev)
    delta = torch.empty(B, h, Lq, device=dev)
    key_len = torch.full((B,), Lk, dtype=torch.int32, device=dev)
    ws = torch.empty(B * Lq * D * 4, dtype=torch.uint8, device=dev)
    out = {"shape": dict(B=B, h=h, Lq=Lq, Lk=Lk, causal=causal, fused=fused)}
    # algorithmic bytes: q, k, v read + o written (fwd); q, k, v, o, dO read + dq, dk, dv written (bwd)
    fb = 2.0 * D * B * (2 * Lq + 2 * Lk)
    bb = 2.0 * D * B * (4 * Lq + 4 * Lk)
    for tc in ("0", "1"):
        os.environ["ZB_ATTN_TC"] = tc
        a = ops.attention_args(q, k, v, o, h, key_len=None if causal else key_len, causal=causal, lse=lse)
        for name, fn, nbytes in (("fwd", lambda: ops.attention_fwd(a), fb),
                                 ("bwd", lambda: ops.attention_bwd(a, d_o, dq, dk, dv, delta, None, None, workspace=ws), bb)):
            fn()
            torch.cuda.synchronize()
            if os.environ.get("ZB_ATTN_TRACE"):
                continue
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(reps):
                    fn()
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1000.0 / (5 * reps)
            out["%s_%s_us" % (name, "tc" if tc == "1" else "mma")] = round(us, 2)
            out["%s_%s_GBs" % (name, "tc" if tc == "1" else "mma")] = round(nbytes / us / 1e3, 1)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if "--trace" in sys.argv:
        os.environ["ZB_ATTN_TRACE"] = "1"
    run(64, 8, 64, 64, False, True)       # configs[1] encoder self attention
    run(64, 8, 64, 64, True, True)        # configs[1] decoder self attention
    run(64, 8, 64, 64, False, False)      # configs[1] cross attention
    run(32, 8, 128, 128, True, True)      # configs[3] lengths
    run(8, 8, 1024, 1024, False, True)    # configs[4] encoder self attention
    run(8, 8, 64, 1024, False, False)     # configs[4] cross attention
