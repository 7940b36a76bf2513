"""Graph-timed LayerNorm forward / backward at the default step's shape (R = 51200 rows, H = 256)."""
import sys
import torch
sys.path.insert(0, '.')
from multi_modal_foundation_model_b200 import ops

R, H, NC = 51200, 256, 4
dev = "cuda"
xs = [torch.randn(R, H, device=dev) for _ in range(NC)]
ys = [torch.empty(R, H, device=dev, dtype=torch.bfloat16) for _ in range(NC)]
g, b = torch.randn(H, device=dev), torch.randn(H, device=dev)
mean, rstd = torch.empty(R, device=dev), torch.empty(R, device=dev)
dys = [torch.randn(R, H, device=dev).to(torch.bfloat16) for _ in range(NC)]
dres = [torch.randn(R, H, device=dev) for _ in range(NC)]
dxb = [torch.empty(R, H, device=dev, dtype=torch.bfloat16) for _ in range(NC)]
dg, db = torch.zeros(H, device=dev), torch.zeros(H, device=dev)


def bench(name, fn, nbytes, iters=40):
    for i in range(3):
        fn(i % NC)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(iters):
            fn(i % NC)
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{name:28s} {us:7.1f} us   {nbytes / us / 1e3:7.0f} GB/s  ({nbytes / us / 1e3 / 6544 * 100:4.1f} % of the measured copy peak)")


bench("layernorm_fwd", lambda i: ops.layernorm_fwd(xs[i], g, b, ys[i], mean, rstd, R=R, H=H), 6.0 * R * H)
bench("layernorm_bwd (dres,dx,dxb)", lambda i: ops.layernorm_bwd(dys[i], xs[i], mean, rstd, g, dres[i], dres[i], dxb[i], ops.NO_DROP, dg, db, R=R, H=H), 16.0 * R * H)
bench("torch copy bf16 (reference)", lambda i: ys[i].copy_(dxb[(i + 1) % NC]), 4.0 * R * H)
