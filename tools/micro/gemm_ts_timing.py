"""clock64 trace of one CTA of the TMA-store GEMM (build with MMFM_NVCC_EXTRA=-DMMFM_DBG_TIMING).
usage: gemm_ts_timing.py {qkv|out|up|down|mulaux}"""
import ctypes as C
import subprocess
import sys
sys.path.insert(0, '.')
import torch
from multi_modal_foundation_model_b200 import _lib

which = sys.argv[1] if len(sys.argv) > 1 else "out"
sys.argv = [sys.argv[0], which, "4"]
exec(open("tools/gemm_one.py").read())
L = _lib.lib()
buf = (C.c_longlong * (8 * 64))()
L.mmfm_debug_read_gemm_ts.argtypes = [C.c_void_p]
L.mmfm_debug_read_gemm_ts(buf)
v = [[buf[r * 64 + i] for i in range(64)] for r in range(8)]
names = ["loads issued (tile)", "first kb landed+issued (tile)", "last MMA issued (tile)", "store: stg_full seen (unit)",
         "store: smem read done (unit)", "in: issue (unit)", "epi: start math (unit)", "epi: arrive (unit)"]
t0 = v[7][63]
print(f"kernel entry 0.00, exit {(v[7][62] - t0) / 1e3:.2f} us (globaltimer)")
for r in range(8):
    row = [x - t0 for x in v[r][:60] if x > 0][:14]
    print(f"{names[r]:34s}", " ".join(f"{x / 1e3:6.2f}" for x in row), "(us)")

cb = (C.c_longlong * 320)()
L.mmfm_debug_read_gemm_cta.argtypes = [C.c_void_p]
L.mmfm_debug_read_gemm_cta(cb)
ent = [cb[2 * i] for i in range(148)]
ext = [cb[2 * i + 1] for i in range(148)]
e0 = min(ent)
print("CTA entry  (us after the first entry): min %.2f  median %.2f  max %.2f" % (0.0, sorted(ent)[74] / 1e3 - e0 / 1e3, (max(ent) - e0) / 1e3))
print("CTA exit   (us after the first entry): min %.2f  median %.2f  max %.2f" % ((min(ext) - e0) / 1e3, (sorted(ext)[74] - e0) / 1e3, (max(ext) - e0) / 1e3))
life = sorted((b - a) / 1e3 for a, b in zip(ent, ext))
print("CTA lifetime (us): min %.2f median %.2f max %.2f" % (life[0], life[74], life[-1]))
