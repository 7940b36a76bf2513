"""Shared helpers of the parity tests (test infrastructure)."""
import os

import numpy as np
import torch

from multi_modal_foundation_model_b200.config import default_model_config

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def small_config(**kw):
    return default_model_config(n_layers=1, hidden_size=128, n_heads=4, inter_size=256, **kw)


def load_small():
    z = np.load(os.path.join(GOLDEN, "mm_small.npz"))
    weights = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w/")}
    return z, weights


def oracle_batch(z, mode):
    """batch dict for oracle.mm_oracle.forward from the fixture (masks = the reference's own)."""
    spikes, target = torch.from_numpy(z["in/spikes"]), torch.from_numpy(z["in/target"])
    attn, ts = torch.from_numpy(z["in/attn"]), torch.from_numpy(z["in/ts"])
    b = {}
    for m, x in (("ap", spikes), ("behavior", target)):
        b[m] = dict(inputs=x, targets=x, attn_mask=attn, timestamp=ts, mask=torch.from_numpy(z[f"{mode}/mask/{m}"]))
    return b


def oracle_params(weights):
    """state_dict -> parameter mapping with the shared mod_emb aliased (mm.py:84-87)."""
    P = dict(weights)
    for k in list(P):
        if k.startswith("decoder_embeddings.") and k.endswith("mod_emb.weight"):
            P[k] = P[k.replace("decoder_embeddings.", "encoder_embeddings.")]
    return P


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()
