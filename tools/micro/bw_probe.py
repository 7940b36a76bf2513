"""HBM bandwidth probes (torch kernels, CUDA events): pure write, pure read, copy."""
import torch
def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
N = 1 << 30
a = torch.empty(N, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
print("write (fill_)  %.0f GB/s" % (2 * N / t(lambda: a.fill_(1.0)) / 1e9))
print("read  (sum)    %.0f GB/s" % (2 * N / t(lambda: a.view(torch.int16).sum()) / 1e9))
print("copy           %.0f GB/s (read+write)" % (4 * N / t(lambda: b.copy_(a)) / 1e9))
for mb in (26, 79, 160):
    n = mb * 1000 * 1000 // 2
    bufs = [torch.empty(n, dtype=torch.bfloat16, device="cuda") for _ in range(6)]
    i = [0]
    def f():
        bufs[i[0] % 6].fill_(1.0); i[0] += 1
    s = t(f, 30)
    print("write %3d MB buffers (6 rotating): %.1f us  %.0f GB/s" % (mb, s * 1e6, 2 * n / s / 1e9))
