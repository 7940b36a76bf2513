"""Per-phase clock64 timestamps of one CTA of attn_bwd_fused2_tc_kernel (library built with MMFM_NVCC_EXTRA=-DMMFM_DBG_TIMING)."""
import ctypes, sys, torch
sys.path.insert(0, '.')
from multi_modal_foundation_model_b200 import ops, _lib
exec(open('tools/attn_bench.py').read().split("def timeit")[0])
for _ in range(3):
    ops.attention_bwd(q, k, v, o, lse, kv, d_o=d_o, delta=delta, dq=dqkv[:, :H], dk=dqkv[:, H:2 * H], dv=dqkv[:, 2 * H:], **kw)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
_lib.lib().mmfm_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
print("rc", _lib.lib().mmfm_debug_read(buf, 64))
t = list(buf)
names = {0: "entry", 1: "setup issued", 2: "after setup sync", 40: "before done wait", 41: "after done wait", 42: "after readout sync", 43: "end"}
for qt in range(2):
    for kh in range(2):
        b = 16 * qt + 4 * kh
        names[4 + b] = f"A({qt},{kh}) wait S"; names[5 + b] = f"A({qt},{kh}) start"; names[6 + b] = f"A({qt},{kh}) done"; names[7 + b] = f"A({qt},{kh}) synced"
        names[12 + b] = f"B({qt},{kh}) wait dP"; names[13 + b] = f"B({qt},{kh}) start"; names[14 + b] = f"B({qt},{kh}) done"; names[15 + b] = f"B({qt},{kh}) synced"
ev = sorted((t[i], names[i]) for i in names if t[i] > 0)
t0 = ev[0][0]
prev = t0
for ts, nm in ev:
    print(f"{ts - t0:8d}  (+{ts - prev:6d})  {nm}")
    prev = ts
