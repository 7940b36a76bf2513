"""Where does the end-to-end step lose time?  Resident vs H2D-fed steps, fp32 vs uint8 spikes, dense vs compact masks;
host time per step (launch-path cost) next to device time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_foundation_model_b200.config import default_model_config
from multi_modal_foundation_model_b200.model import build_model
from multi_modal_foundation_model_b200.synthetic import DevicePrefetcher, make_batch, make_mod_dict

MODES = ("encoding", "decoding", "token_masking")
dev = torch.device("cuda", 0)
B, N = 256, 668
torch.manual_seed(42)
model = build_model(N, 2, default_model_config()).to(dev).train()
hb32 = [make_batch(B, N, 2, 100, step=i, pin=True) for i in range(3)]
hb8 = [dict(h, spikes_data=h["spikes_data"].to(torch.uint8).pin_memory()) for h in hb32]
pf = DevicePrefetcher(dev)


def run(name, host_batches, compact, resident, steps=30):
    devb = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in h.items()} for h in host_batches]
    pf._next = None

    def one(i):
        if resident:
            db = devb[i % 3]
        else:
            if pf._next is None:
                pf.put(host_batches[i % 3])
            db = pf.get()
            pf.put(host_batches[(i + 1) % 3])
        md = make_mod_dict(db, ["ap", "behavior"], MODES[i % 3], device=dev, compact_masks=compact)
        out = model(md)
        out.loss.backward()
        model.zero_grad(set_to_none=True)
    for i in range(6):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        one(i)
    e1.record()
    host = (time.perf_counter() - t0) / steps * 1e3
    torch.cuda.synchronize()
    print(f"{name:42s} device {e0.elapsed_time(e1) / steps:7.3f} ms/step   host enqueue {host:7.3f} ms/step", flush=True)


run("resident fp32 dense masks", hb32, False, True)
run("resident fp32 compact masks", hb32, True, True)
run("resident uint8 dense masks", hb8, False, True)
run("resident uint8 compact masks", hb8, True, True)
run("e2e fp32 dense masks", hb32, False, False)
run("e2e fp32 compact masks", hb32, True, False)
run("e2e uint8 dense masks", hb8, False, False)
run("e2e uint8 compact masks", hb8, True, False)
