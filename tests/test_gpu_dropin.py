"""Zero-edit drop-in: ``install()`` rebinds ``multi_modal.mm.MultiModal.forward`` at class level;
the UNMODIFIED reference model object, trainer (``trainer/base.py:182-270,302-308``) and pickling then run on the B200
kernels.  The reference sources come from the vendored, git-ignored ``baseline/_ref`` (the GPU box has no
``/root/reference``); the tests skip when it is absent.

Tolerances as everywhere (bf16 tensor-core operands vs the reference's fp32): loss rel 2e-3, preds abs 3e-2, gradients
rel-L2 2e-2 and cosine >= 0.999; masks / counts bit-exact.
"""
import os
import random
import subprocess
import sys

import numpy as np
import pytest
import torch

from _util import cosine, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_loader  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref not vendored")]


@pytest.fixture
def dropin():
    import multi_modal_foundation_model_b200 as mmfm
    ref_loader.activate(trainer=True)
    yield mmfm
    mmfm.uninstall()


def _batches(n, B, N, pad=0):
    from multi_modal_foundation_model_b200.synthetic import make_batch
    out = []
    for i in range(n):
        b = make_batch(B, N, 2, 100, step=i, pad_bins=pad)
        b.pop("_regions_T")
        out.append(b)
    return out


@pytest.mark.parametrize("mode,over", [
    ("token_masking", {}),
    ("encoding", {}),
    ("decoding", {"decoder.decoder_causal_mask": True}),
    ("token_masking", {"decoder.decoder_sep_mask": True}),
])
def test_same_model_object_reference_path_vs_b200_path(dropin, mode, over):
    """ONE reference model object: its own PyTorch forward/backward (fp32, TF32 off) and, after install(), the B200
    kernels -- same weights, same batch, same masks."""
    from multi_modal_foundation_model_b200.synthetic import make_mod_dict
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = ref_loader.load_config(over)
    torch.manual_seed(5)
    N, B = 96, 6
    model = ref_loader.build_reference_model(cfg, N, 2).cuda().eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    batch = _batches(1, B, N, pad=10)[0]
    md = make_mod_dict(batch, ["ap", "behavior"], mode, device="cuda")
    torch.manual_seed(99)                       # the reference Masker draws from the CPU stream
    ref = model(md)
    ref.loss.backward()
    g_ref = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    masks = {m: md[m]["inputs_mask"].clone() for m in md}
    model.zero_grad(set_to_none=True)

    dropin.install(mask_stream="reference")
    assert type(model).__module__ == "multi_modal.mm"
    md2 = make_mod_dict(batch, ["ap", "behavior"], mode, device="cuda")
    torch.manual_seed(99)
    out = model(md2)
    out.loss.backward()
    torch.cuda.synchronize()
    assert type(out).__module__ == "multi_modal.mm"                       # the reference's own MultiModalOutput
    for m in md:
        assert torch.equal(md2[m]["inputs_mask"], masks[m]), m           # bit-exact masks (mm.py:270,272)
        assert torch.equal(md2[m]["targets_mask"], masks[m])
        assert int(out.mod_n_examples[m]) == int(ref.mod_n_examples[m])
        assert (out.mod_preds[m] - ref.mod_preds[m]).abs().max().item() < 3e-2
        assert md2[m]["preds"] is out.mod_preds[m] and md2[m]["gt"] is md2[m]["targets"]
    assert abs(out.loss.item() - ref.loss.item()) <= 2e-3 * abs(ref.loss.item())
    for n, p in model.named_parameters():
        if n not in g_ref or g_ref[n].norm() < 1e-6:
            continue
        r, c = rel_l2(p.grad, g_ref[n]), cosine(p.grad, g_ref[n])
        assert r < 2e-2 and c > 0.999, f"{n}: rel-L2 {r:.4g} cosine {c:.6f}"


def test_unmodified_trainer_epoch_eval_save_load(dropin, tmp_path):
    """train_multi_modal.py:160-231 on synthetic batches: the reference's own MultiModalTrainer.train_epoch / eval_epoch /
    save_model with torch.optim.AdamW + OneCycleLR, then torch.load of the pickled module -- in this process and in a
    process that imports ONLY the reference."""
    dropin.install()                                                     # device-side mask sampler (default)
    cfg = ref_loader.load_config(trainer_overrides={"wandb.use": False, "training.num_epochs": 1})
    torch.manual_seed(42)
    random.seed(42)
    N, B = 80, 8
    model = ref_loader.build_reference_model(cfg, N, 2).cuda()
    opt = torch.optim.AdamW(model.parameters(), lr=cfg.optimizer.lr, weight_decay=cfg.optimizer.wd, eps=cfg.optimizer.eps)
    train, evalb = _batches(6, B, N), _batches(2, B, N)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, total_steps=len(train), max_lr=cfg.optimizer.lr,
                                                pct_start=cfg.optimizer.warmup_pct, div_factor=cfg.optimizer.div_factor)
    tr = ref_loader.make_trainer(model, train, evalb, opt, sched, cfg, "cuda", str(tmp_path), [N], mixed_training=True)
    w0 = {n: p.detach().clone() for n, p in model.named_parameters()}
    res = tr.train_epoch(0)
    assert np.isfinite(res["train_loss"]) and res["train_loss"] > 0
    moved = sum(float((p.detach() - w0[n]).abs().max()) > 0 for n, p in model.named_parameters())
    assert moved > 0.9 * len(w0), f"only {moved} of {len(w0)} parameters were updated by the optimizer"
    ev = tr.eval_epoch()
    assert np.isfinite(ev["eval_loss"]) and "eval_trial_avg_r2" in ev
    assert ev["eval_preds"][0]["ap"].shape == (2 * B, 100, N)
    tr.save_model(name="last", epoch=0)
    path = os.path.join(str(tmp_path), "model_last.pt")
    loaded = torch.load(path, weights_only=False)["model"]
    assert type(loaded).__module__ == "multi_modal.mm" and "_b200_engine" not in loaded.__dict__
    for (n, a), (_, b) in zip(model.state_dict().items(), loaded.state_dict().items()):
        assert torch.equal(a, b), n
    # the loaded copy runs on the B200 path again (fresh engine), eval-mode, same answer as the live model
    from multi_modal_foundation_model_b200.synthetic import make_mod_dict
    model.eval(), loaded.eval()
    md = lambda: make_mod_dict(evalb[0], ["ap", "behavior"], "encoding", device="cuda")
    with torch.no_grad():
        a, b = model(md()), loaded(md())
    assert abs(a.loss.item() - b.loss.item()) < 1e-6 * abs(a.loss.item())
    # a process that has never imported this package unpickles the checkpoint with the reference alone
    code = ("import sys, torch; sys.path.insert(0, %r); m = torch.load(%r, weights_only=False, map_location='cpu')['model'];"
            "assert not any(k.startswith('multi_modal_foundation_model_b200') for k in sys.modules);"
            "print(type(m).__module__, type(m).__name__, sum(p.numel() for p in m.parameters()))"
            % (os.path.join(ref_loader.ref_root(), "src"), path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ref_loader.ref_root())
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.split()[:2] == ["multi_modal.mm", "MultiModal"]


def test_device_mask_sampler_matches_restatement(dropin):
    """stream='device': the (B,T) Bernoulli masks drawn inside mmfm_mask_prep == oracle/philox_ref.mask_bernoulli for
    the step's seed; different every step; right rate."""
    from multi_modal_foundation_model_b200.synthetic import make_mod_dict
    from oracle import philox_ref as px
    dropin.install(mask_stream="device")
    cfg = ref_loader.load_config()
    torch.manual_seed(1)
    N, B = 64, 32
    model = ref_loader.build_reference_model(cfg, N, 2).cuda().train()
    batch = _batches(1, B, N, pad=5)[0]
    prev = None
    for step in range(3):
        md = make_mod_dict(batch, ["ap", "behavior"], "token_masking", device="cuda")
        out = model(md)
        eng = model.b200_engine()
        seed = int(eng.seed.item()) & 0xFFFFFFFFFFFFFFFF
        for k, m in enumerate(("ap", "behavior")):
            want = torch.from_numpy(px.mask_bernoulli(seed, k, B, 100, cfg.model.masker.ratio)) & batch["time_attn_mask"]
            assert torch.equal(md[m]["inputs_mask"].cpu(), want), (step, m)
            assert int(out.mod_n_examples[m]) == int(want.sum()) * (N if m == "ap" else 2)
        cur = md["ap"]["inputs_mask"].clone()
        assert prev is None or not torch.equal(prev, cur)
        prev = cur
        rate = cur[:, :95].float().mean().item()
        assert abs(rate - cfg.model.masker.ratio) < 0.05
