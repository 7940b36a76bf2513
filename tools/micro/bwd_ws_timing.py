"""globaltimer trace of one item of one CTA of attn_bwd_ws_kernel (build with MMFM_NVCC_EXTRA=-DMMFM_DBG_TIMING):
the passes of one math warp (warp 1: quadrant 1, column group 0).
usage: bwd_ws_timing.py [dropout 0/1]"""
import ctypes, sys, torch
sys.path.insert(0, '.')
from multi_modal_foundation_model_b200 import ops, _lib
drop = sys.argv[1] if len(sys.argv) > 1 else "1"
sys.argv = [sys.argv[0], "256", drop]
exec(open('tools/attn_bench.py').read().split("def timeit")[0])
for _ in range(3):
    ops.attention_bwd(q, k, v, o, lse, kv, d_o=d_o, delta=delta, dq=dqkv[:, :H], dk=dqkv[:, H:2 * H], dv=dqkv[:, 2 * H:], **kw)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
L = _lib.lib()
L.mmfm_debug_read_bwd_ws.argtypes = [ctypes.c_void_p]
L.mmfm_debug_read_bwd_ws(buf)
t = list(buf)
names = {0: "item start", 1: "side data landed", 40: "before read-out of the previous item", 42: "after read-out", 41: "q1: after pass A", 43: "q1: before pass B"}
for qt in range(2):
    for kh in range(2):
        b = 16 * qt + 4 * kh
        names[4 + b] = f"A(q{qt},k{kh}) wait S"; names[5 + b] = f"A(q{qt},k{kh}) S ready"; names[6 + b] = f"A(q{qt},k{kh}) math done"
        names[12 + b] = f"B(q{qt},k{kh}) wait dP"; names[13 + b] = f"B(q{qt},k{kh}) dP ready"; names[14 + b] = f"B(q{qt},k{kh}) math done"
ev = sorted((t[i], names[i]) for i in names if t[i] > 0)
t0 = ev[0][0]
prev = t0
for ts, nm in ev:
    print(f"{(ts - t0) / 1e3:8.2f} us  (+{(ts - prev) / 1e3:5.2f})  {nm}")
    prev = ts
