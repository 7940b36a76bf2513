"""Generate the golden fixtures of tests/golden/ by running the UNMODIFIED reference classes
(/root/reference/src, imported read-only) on CPU in fp32.  Run in the build container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Fixtures (small on purpose; committed):
  mm_small.npz      reduced model (H=128, 4 heads, 1+1 layers, N=40, T=100, B=4, right-padded trials): weights,
                    inputs, the reference's masks, and for the three training modes of trainer/base.py:84-99 the
                    loss / per-modality sums / counts / predictions; full gradients for `token_masking`, gradient
                    norms for the other two.
  masker.npz        (B,T) masks of the reference Masker (temporal mode) for several seeds / shapes / ratios.
  init_default.npz  checksums of the reference's initial weights at seed 42 (default mm.yaml, N=64) -- pins that our
                    module classes draw the same initial parameters.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _reference as ref  # noqa: E402
from multi_modal_foundation_model_b200.synthetic import make_batch, make_mod_dict  # noqa: E402

SMALL = {
    "encoder.transformer.n_layers": 1, "decoder.transformer.n_layers": 1,
    "encoder.transformer.hidden_size": 128, "decoder.transformer.hidden_size": 128,
    "encoder.transformer.n_heads": 4, "decoder.transformer.n_heads": 4,
    "encoder.transformer.inter_size": 256, "decoder.transformer.inter_size": 256,
}


def main():
    torch.set_num_threads(4)
    assert ref.available()
    # ---------------- model fixture ----------------
    cfg = ref.load_config(SMALL)
    torch.manual_seed(7)
    N, NB, B, T = 40, 2, 4, 100
    model = ref.build_reference_model(cfg, N, NB)
    # perturb LayerNorm affines / biases so they are exercised
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    model.eval()   # dropout off; the Masker stays active (force_active)
    out = {f"w/{k}": v.detach().numpy() for k, v in model.state_dict().items()}
    batch = make_batch(B, N, NB, T, step=0, pad_bins=20)
    out["in/spikes"] = batch["spikes_data"].numpy()
    out["in/target"] = batch["target"].numpy()
    out["in/attn"] = batch["time_attn_mask"].numpy()
    out["in/ts"] = batch["spikes_timestamps"].numpy()
    for mode in ("token_masking", "encoding", "decoding"):
        md = make_mod_dict(batch, ["ap", "behavior"], mode)
        torch.manual_seed(123)
        model.zero_grad()
        o = model(md)
        o.loss.backward()
        out[f"{mode}/loss"] = o.loss.detach().numpy()
        for m in ("ap", "behavior"):
            out[f"{mode}/mask/{m}"] = md[m]["inputs_mask"].numpy()
            out[f"{mode}/mod_loss/{m}"] = o.mod_loss[m].detach().numpy()
            out[f"{mode}/n/{m}"] = o.mod_n_examples[m].numpy()
            out[f"{mode}/preds/{m}"] = o.mod_preds[m].detach().numpy()
        for n, p in model.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            if mode == "token_masking":
                out[f"{mode}/grad/{n}"] = g.numpy().copy()
            else:
                out[f"{mode}/gnorm/{n}"] = np.float64(g.double().norm().item())
    np.savez_compressed(os.path.join(HERE, "mm_small.npz"), **out)

    # ---------------- masker fixture ----------------
    from models.masker import Masker
    mk = {}
    for i, (seed, shape, ratio) in enumerate([(42, (16, 100, 32), 0.3), (1, (4, 100, 7), 0.1), (5, (8, 50, 2), 0.5),
                                               (9, (3, 100, 16), 0.3)]):
        c = ref.load_config()
        c["model"]["masker"]["ratio"] = ratio
        m = Masker(c.model.masker)
        torch.manual_seed(seed)
        regions = np.array([["CA1"] * shape[2]] * shape[0])
        calls = []
        for _ in range(3):   # three consecutive calls: pins the generator consumption between calls
            x = torch.poisson(torch.full(shape, 0.3))
            # the device-side torch.rand of masker.py:161 draws from the CPU generator when the input is a CPU
            # tensor; a CUDA input leaves the CPU stream untouched.  Mimic the CUDA case by restoring the state
            # after the call except for what the bernoulli draws consumed -- done by re-running with rand stubbed.
            real_rand = torch.rand
            torch.rand = lambda *a, **k: torch.zeros(a[0]) if a else real_rand(*a, **k)
            try:
                _, msk = m(x, regions)
            finally:
                torch.rand = real_rand
            calls.append(msk[:, :, 0].numpy().astype(np.int8))
        mk[f"case{i}/meta"] = np.array([seed, *shape, int(ratio * 1000)])
        mk[f"case{i}/masks"] = np.stack(calls)
    np.savez_compressed(os.path.join(HERE, "masker.npz"), **mk)

    # ---------------- init fixture ----------------
    sys.path.insert(0, os.path.join(ref.REF_ROOT, "src"))
    cfg = ref.load_config()
    torch.manual_seed(42)
    model = ref.build_reference_model(cfg, 64, 2)
    init = {}
    for k, v in model.state_dict().items():
        init[k] = np.array([v.double().sum().item(), v.double().abs().sum().item(), float(v.flatten()[0])])
    np.savez_compressed(os.path.join(HERE, "init_default.npz"), **init)
    print("golden fixtures written:", [f for f in os.listdir(HERE) if f.endswith(".npz")])


if __name__ == "__main__":
    main()
