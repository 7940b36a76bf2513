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


# ------------------------------------------------------------------------------------------------------------
# torch restatement of oracle/philox_ref.py (so that bench-size dropout fields can be produced on the GPU in
# seconds); checked bit for bit against the numpy original in tests/test_host_cpu.py
# ------------------------------------------------------------------------------------------------------------
_M0, _M1, _W0, _W1, _MASK32 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF


def _philox_words_torch(g, site, c3, seed, rounds=7):
    """g: int64 tensor of group indices -> (4, len(g)) int64 words (each < 2^32)."""
    c0, c1 = g & _MASK32, (g >> 32) & _MASK32
    c2 = torch.full_like(g, site & _MASK32)
    c3 = torch.full_like(g, c3)
    k0, k1 = seed & _MASK32, (seed >> 32) & _MASK32
    for _ in range(rounds):
        # products wrap modulo 2^64 in int64; the masks recover the unsigned halves
        p0, p1 = c0 * _M0, c2 * _M1
        hi0, lo0 = (p0 >> 32) & _MASK32, p0 & _MASK32
        hi1, lo1 = (p1 >> 32) & _MASK32, p1 & _MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + _W0) & _MASK32, (k1 + _W1) & _MASK32
    return torch.stack([c0, c1, c2, c3])


def _bytes_of(words):
    """(4, G) int64 words -> (G, 16) uint8, little endian inside each word (= numpy's .view(uint8))."""
    sh = torch.arange(4, device=words.device, dtype=torch.int64) * 8
    return ((words.t()[:, :, None] >> sh) & 0xFF).reshape(words.shape[1], 16).to(torch.uint8)


def philox_bytes_torch(seed, site, rows, cols, device, row0=0):
    from oracle import philox_ref as px  # noqa: F401  (layout documented there)
    gpr = (cols + 15) // 16
    g = torch.arange(row0 * gpr, (row0 + rows) * gpr, device=device, dtype=torch.int64)
    return _bytes_of(_philox_words_torch(g, site, 0, seed)).reshape(rows, gpr * 16)[:, :cols]


def philox_keep_torch(seed, site, rows, cols, p, device, row0=0):
    from oracle import philox_ref as px
    by = philox_bytes_torch(seed, site, rows, cols, device, row0)
    return (by >= px.drop_threshold(p)).float() * px.keep_scale(p)


def philox_prob_bytes_torch(seed, site, rows, cols, device, row0=0):
    nblk = (cols + 63) // 64
    g = torch.arange(row0 * nblk * 4, (row0 + rows) * nblk * 4, device=device, dtype=torch.int64)
    by = _bytes_of(_philox_words_torch(g, site, 1, seed)).reshape(rows, nblk, 4, 8, 2)     # (row, blk, q, n, e)
    return by.permute(0, 1, 3, 2, 4).reshape(rows, nblk * 64)[:, :cols]


def philox_prob_keep_torch(seed, site, rows, cols, p, device, row0=0):
    from oracle import philox_ref as px
    by = philox_prob_bytes_torch(seed, site, rows, cols, device, row0)
    return (by >= px.drop_threshold(p)).float() * px.keep_scale(p)
