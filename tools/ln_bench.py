"""add+LN kernels alone at the training-step shape (4096 x 512): forward and backward, CUDA graph of 30 launches
(the step has 30 of each), data L2-hot as in the step.  ZB_LN1P_WARPS selects the backward variant."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from zero_b200 import ops  # noqa: E402

dev = torch.device("cuda")
N, D = 4096, 512
g = torch.Generator(device="cuda").manual_seed(0)
x, y, d1, d2 = [torch.randn(N, D, generator=g, device=dev).to(torch.bfloat16) for _ in range(4)]
out, ds = torch.empty_like(x), torch.empty_like(x)
mean, rstd = torch.empty(N, device=dev), torch.empty(N, device=dev)
scale, offset = torch.ones(D, device=dev), torch.zeros(D, device=dev)
dscale, doffset, dbias = [torch.zeros(D, device=dev) for _ in range(3)]


def timeit(fn, reps=30):
    fn()
    torch.cuda.synchronize()
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
    return e0.elapsed_time(e1) * 1000.0 / (5 * reps)


fwd = timeit(lambda: ops.add_ln_fwd(x, y, out, scale, offset, mean, rstd, 1e-8))
bwd = timeit(lambda: ops.add_ln_bwd(x, y, d1, d2, mean, rstd, scale, ds, dscale, doffset, dbias))
print(json.dumps({"ZB_LN1P_WARPS": os.environ.get("ZB_LN1P_WARPS", "32"), "fwd_us": round(fwd, 2), "bwd_us": round(bwd, 2),
                  "fwd_GBs": round(N * D * 2 * 3 / fwd / 1e3, 1), "bwd_GBs": round(N * D * 2 * 5 / bwd / 1e3, 1)}))
