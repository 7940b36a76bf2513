"""Per-shape timing of one recorded step (CUDA events around every C-ABI call): which GEMM shapes sit furthest from
their HBM floor.  Usage: python tools/step_profile.py [batch] [neurons]"""
import sys, collections, torch
sys.path.insert(0, '.')
from multi_modal_foundation_model_b200 import ops
from multi_modal_foundation_model_b200.config import default_model_config
from multi_modal_foundation_model_b200.model import build_model
from multi_modal_foundation_model_b200.synthetic import make_batch, make_mod_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 668
torch.manual_seed(0)
model = build_model(N, 2, default_model_config()).cuda().train()
md = make_mod_dict(make_batch(B, N, 2, 100), ["ap", "behavior"], "encoding", device="cuda")
for _ in range(3):
    out = model(dict((k, dict(v)) for k, v in md.items())); out.loss.backward(); model.zero_grad(set_to_none=True)
torch.cuda.synchronize()
pl = model.engine().last_plan
agg = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
reps = 5
for _ in range(reps):
    for name, meta, t in ops.run_recorded_timed(pl.fwd_calls) + ops.run_recorded_timed(pl.bwd_calls):
        key = meta.get("tag", name)
        a = agg[key]; a[0] += t; a[1] += 1; a[2] += meta.get("bytes", 0.0); a[3] += meta.get("flops", 0.0)
tot = sum(a[0] for a in agg.values())
print(f"{'call':40s} {'n/step':>6s} {'us/call':>8s} {'ms/step':>8s} {'share':>6s} {'GB/s':>7s} {'TF/s':>7s} {'floor us':>8s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    n = a[1] / reps
    us = a[0] / a[1] * 1e3
    gbs = a[2] / (a[0] * 1e-3) / 1e9 if a[2] else 0
    tfs = a[3] / (a[0] * 1e-3) / 1e12 if a[3] else 0
    floor = max(a[2] / a[1] / 6.544e3, a[3] / a[1] / 1.37e9) * 1e-3 if (a[2] or a[3]) else 0   # us
    print(f"{k:40s} {n:6.0f} {us:8.1f} {a[0] / reps:8.3f} {100 * a[0] / tot:5.1f}% {gbs:7.0f} {tfs:7.0f} {floor:8.1f}")
