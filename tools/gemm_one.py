"""One GEMM flavour in a short loop (ncu target): gemm_one.py {qkv|out|up|down|mulaux} [iters]"""
import sys
import torch
sys.path.insert(0, '.')
from multi_modal_foundation_model_b200 import ops
from multi_modal_foundation_model_b200._lib import ACT_GELU_DG, ACT_MULAUX

which = sys.argv[1] if len(sys.argv) > 1 else "qkv"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
R, H, I = 51200, 256, 512
dev, bf = "cuda", torch.bfloat16
NC = 3
mk = lambda r, c, dt=bf: [torch.randn(r, c, device=dev).to(dt) for _ in range(NC)]
seed = torch.tensor([1], dtype=torch.int64, device=dev)
if which == "qkv":
    x, w, out, b = mk(R, H), mk(3 * H, H)[0], mk(R, 3 * H), torch.randn(3 * H, device=dev)
    fn = lambda i: ops.gemm_tn(x[i], w, out[i], bias=b)
elif which == "out":
    x, w, out, res, b = mk(R, H), mk(H, H)[0], mk(R, H, torch.float32), mk(R, H, torch.float32), torch.randn(H, device=dev)
    fn = lambda i: ops.gemm_tn(x[i], w, out[i], bias=b, res=res[i])
elif which == "down":
    x, w, out, res, b = mk(R, I), mk(H, I)[0], mk(R, H, torch.float32), mk(R, H, torch.float32), torch.randn(H, device=dev)
    fn = lambda i: ops.gemm_tn(x[i], w, out[i], bias=b, res=res[i], drop=ops.DropSpec(seed, 3, 0.4))
elif which == "up":
    x, w, g, dg, b = mk(R, H), mk(I, H)[0], mk(R, I), mk(R, I), torch.randn(I, device=dev)
    fn = lambda i: ops.gemm_tn(x[i], w, g[i], bias=b, act=ACT_GELU_DG, D2=dg[i])
else:
    x, w, g, dg = mk(R, H), mk(I, H)[0], mk(R, I), mk(R, I)
    fn = lambda i: ops.gemm_tn(x[i], w, g[i], act=ACT_MULAUX, aux=dg[i])
for i in range(3):
    fn(i % NC)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters):
    fn(i % NC)
e1.record()
torch.cuda.synchronize()
print(which, f"{e0.elapsed_time(e1) / iters * 1e3:.1f} us")
