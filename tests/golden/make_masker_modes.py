"""Golden (B,T) column-0 masks of the UNMODIFIED reference Masker (models/masker.py) in every masking mode, three
consecutive calls each (pins the generator consumption between calls, torch CPU stream + python `random`).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_masker_modes.py     (build container; writes masker_modes.npz)
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from baseline import ref_loader  # noqa: E402

CASES = [
    # (mode, seed, (B,T,C), overrides)
    ("temporal", 11, (6, 100, 9), dict(ratio=0.3)),
    ("temporal", 12, (5, 60, 4), dict(ratio=0.4, expand_prob=1.0, max_timespan=4)),
    ("random_token", 13, (4, 100, 3), dict(ratio=0.2)),
    ("causal", 14, (4, 50, 5), dict(max_timespan=3)),
    ("causal", 15, (4, 50, 5), dict(max_timespan=2, causal_zero=False)),
    ("neuron", 16, (16, 20, 7), dict(ratio=0.5)),
    ("random", 17, (5, 30, 6), dict(ratio=0.3)),
    ("co-smooth", 18, (3, 20, 8), dict(channels=[0, 3])),
    ("co-smooth", 19, (3, 20, 8), dict(channels=[2, 5])),
    ("forward-pred", 20, (3, 40, 4), dict(timesteps=[30, 31, 32, 39])),
    ("inter-region", 21, (6, 20, 10), dict(n_mask_regions=2)),
    ("intra-region", 22, (6, 20, 10), dict(ratio=0.5, n_mask_regions=2)),
]
REGIONS = ["CA1", "DG", "LP", "PO"]


def regions_for(B, C):
    return np.array([[REGIONS[(c + b) % len(REGIONS)] for c in range(C)] for b in range(B)])


def reference_masks(mode, seed, shape, over, n_calls=3):
    ref_loader.activate()
    from models.masker import Masker
    cfg = ref_loader.load_config()
    mk_cfg = cfg.model.masker
    mk_cfg["mode"] = mode
    for k, v in over.items():
        mk_cfg[k] = v
    m = Masker(mk_cfg)
    torch.manual_seed(seed)
    random.seed(seed)
    regions = regions_for(shape[0], shape[2])
    out = []
    real_rand = torch.rand
    for _ in range(n_calls):
        x = torch.ones(shape)
        # masker.py:161 draws torch.rand on the INPUT's device: the CUDA generator in training, never the CPU stream
        torch.rand = lambda *a, **k: torch.zeros(a[0])
        try:
            _, msk = m(x, regions)
        finally:
            torch.rand = real_rand
        out.append(msk[:, :, 0].numpy().astype(np.int8))
    return np.stack(out)


def main():
    z = {}
    for i, (mode, seed, shape, over) in enumerate(CASES):
        z[f"case{i}"] = reference_masks(mode, seed, shape, over)
    np.savez_compressed(os.path.join(HERE, "masker_modes.npz"), **z)
    print("written", {k: v.shape for k, v in z.items()})


if __name__ == "__main__":
    main()
