import sys, time, torch
sys.path.insert(0, '.')
from multi_modal_foundation_model_b200.config import default_model_config
from multi_modal_foundation_model_b200.model import build_model
from multi_modal_foundation_model_b200.synthetic import make_batch, make_mod_dict
for B, N in ((16, 668), (256, 668)):
    torch.manual_seed(0)
    model = build_model(N, 2, default_model_config()).cuda().train()
    batch = make_batch(B, N, 2, 100)
    md = make_mod_dict(batch, ["ap", "behavior"], "encoding", device="cuda")
    for it in range(3):
        out = model(dict((k, dict(v)) for k, v in md.items())); out.loss.backward(); model.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    pl = model.engine().last_plan
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    n = 10
    tf = tb = 0
    for it in range(n):
        e0.record(); pl.run_forward(); e1.record(); pl.run_backward(); e2.record(); torch.cuda.synchronize()
        tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
    t0 = time.time()
    for it in range(n):
        out = model(dict((k, dict(v)) for k, v in md.items())); out.loss.backward(); model.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    wall = (time.time() - t0) / n * 1e3
    print(f"B={B} N={N}: fwd {tf/n:.3f} ms bwd {tb/n:.3f} ms (launches {pl.n_fwd}+{pl.n_bwd}); full step wall {wall:.3f} ms -> {B/wall*1e3:.0f} trials/s; loss {out.loss.item():.4f}")
