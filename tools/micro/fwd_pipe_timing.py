"""Per-phase clock64 timestamps of the softmax warps of one item of attn_fwd_pipe_kernel (library built with
MMFM_NVCC_EXTRA=-DMMFM_DBG_TIMING).  Rows: (tile X, column group g) warp sets."""
import ctypes, sys, torch
sys.path.insert(0, '.')
from multi_modal_foundation_model_b200 import ops, _lib
exec(open('tools/attn_bench.py').read().split("def timeit")[0])
for _ in range(3):
    ops.attention_fwd(q, k, v, o, lse, kv, **kw)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
_lib.lib().mmfm_debug_read_pipe.argtypes = [ctypes.c_void_p, ctypes.c_int]
print("rc", _lib.lib().mmfm_debug_read_pipe(buf, 512))
names = ["item start", "q_full ok", "s_full ok", "pass1 done", "max exchanged", "pass2 done", "sum exchanged", "pv_done ok", "epilogue done"]
t0 = min(buf[w * 64] for w in range(4) if buf[w * 64] > 0)
for w in range(4):
    ts = [buf[w * 64 + i] for i in range(9)]
    print(f"tile {w >> 1} group {w & 1}: " + "  ".join(f"{names[i]} {ts[i] - t0}" for i in range(9) if ts[i] > 0))
