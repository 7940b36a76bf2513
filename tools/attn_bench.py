"""Micro-benchmark of the attention kernels: attn_bench.py [B] [dropout 0/1] [heads] [d_head] [S] [mask mode]
(default-model shape B=256, 8 heads, d=32, S=200; scaled config: 16 1 16 64 1000; mask mode 1 = encoder mask
eye | key_valid (default), 2 = causal: FLOPs are still counted for the full S x S product)."""
import sys
import torch
sys.path.insert(0, '.')
from multi_modal_foundation_model_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
drop_on = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
nh = int(sys.argv[3]) if len(sys.argv) > 3 else 8
d = int(sys.argv[4]) if len(sys.argv) > 4 else 32
S = int(sys.argv[5]) if len(sys.argv) > 5 else 200
mask_mode = int(sys.argv[6]) if len(sys.argv) > 6 else 1
H = nh * d
dev = "cuda"
qkv = torch.randn(B * S, 3 * H, device=dev).to(torch.bfloat16)
q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
o = torch.zeros(B * S, H, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(B, nh, S, device=dev)
kv = torch.ones(B, S, dtype=torch.uint8, device=dev)
nkb = (S + 63) // 64
p_keep = torch.zeros(B * nh * S * nkb * 4, dtype=torch.int16, device=dev)
seed = torch.tensor([12345], dtype=torch.int64, device=dev)
dp = ops.DropSpec(seed, 1, 0.4) if drop_on else ops.NO_DROP
do = ops.DropSpec(seed, 2, 0.4) if drop_on else ops.NO_DROP
d_o = torch.randn(B * S, H, device=dev).to(torch.bfloat16)
dqkv = torch.zeros(B * S, 3 * H, device=dev, dtype=torch.bfloat16)
delta = torch.zeros(B, nh, S, device=dev)
kw = dict(B=B, n_heads=nh, Sq=S, Sk=S, d_head=d, mask_mode=mask_mode, drop_p=dp, drop_o=do, p_keep=p_keep)


def timeit(name, fn, flops, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{name:28s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s")


fl = 4.0 * B * nh * S * S * d
timeit("attention fwd", lambda: ops.attention_fwd(q, k, v, o, lse, kv, **kw), fl)
timeit("attention bwd (prep+dq+dkv)", lambda: ops.attention_bwd(q, k, v, o, lse, kv, d_o=d_o, delta=delta, dq=dqkv[:, :H], dk=dqkv[:, H:2 * H], dv=dqkv[:, 2 * H:], **kw), 2 * fl)
