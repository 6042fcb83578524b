"""Where is the floor of a few-hundred-row projection?  Times y = x W (+b) for the decode-step shapes through
zb_gemm with the tcgen05 kernel, the skinny kernel and the skinny kernel with the cluster k split, with the weights
hot in L2 (same W every launch) and cold (launches rotate through > 126 MB of weights), 30 launches per CUDA graph
replay like one decode step.  Prints us per launch."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from zero_b200 import ops  # noqa: E402


def time_graph(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    dev = torch.device("cuda")
    pool = torch.randn(96 * 1024 * 1024, device=dev).to(torch.bfloat16)   # 192 MB of weights
    out = []
    for (m, n, k) in [(256, 512, 512), (256, 2048, 512), (256, 512, 2048), (256, 1024, 1024), (64, 512, 512)]:
        x = torch.randn(m, k, device=dev).to(torch.bfloat16)
        y = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
        bias = torch.zeros(n, device=dev)
        nw = pool.numel() // (n * k)
        ws = [pool[i * n * k:(i + 1) * n * k].view(k, n) for i in range(nw)]
        for name, env in [("tcgen05", {"ZB_SKINNY_GEMM": "0"}), ("skinny", {"ZB_SKINNY_GEMM": "1", "ZB_SKINNY_SPLIT": "0"}),
                          ("skinny+ksplit", {"ZB_SKINNY_GEMM": "1", "ZB_SKINNY_SPLIT": "1"})]:
            os.environ.update(env)
            state = {"i": 0}

            def hot():
                for _ in range(30):
                    ops.linear_fwd(x, ws[0], bias, y)

            def cold():
                for _ in range(30):
                    state["i"] = (state["i"] + 7) % nw
                    ops.linear_fwd(x, ws[state["i"]], bias, y)

            t_hot = time_graph(hot) / 30
            # cold: a fresh graph per measurement would pin the same 30 weights; use 8 graphs over disjoint weights
            graphs = []
            for _ in range(8):
                g = torch.cuda.CUDAGraph()
                cold()
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    cold()
                graphs.append(g)
            for g in graphs:
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                for g in graphs:
                    g.replay()
            e1.record()
            torch.cuda.synchronize()
            t_cold = e0.elapsed_time(e1) * 1e3 / (3 * 8 * 30)
            out.append({"m": m, "n": n, "k": k, "kernel": name, "us_hot": round(t_hot, 2), "us_cold": round(t_cold, 2)})
            print(json.dumps(out[-1]), flush=True)
    # floor of a trivial kernel in the same harness
    a = torch.zeros(256, 512, dtype=torch.bfloat16, device=dev)
    b = torch.zeros_like(a)

    def triv():
        for _ in range(30):
            ops.add2d(a, None, b)
    try:
        print(json.dumps({"kernel": "add2d 256x512 (copy)", "us": round(time_graph(triv) / 30, 2)}))
    except Exception as e:  # noqa: BLE001
        print("add2d probe failed:", e)


if __name__ == "__main__":
    main()
