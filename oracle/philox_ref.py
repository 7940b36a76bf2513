"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the dropout random stream of the CUDA path.

The reference draws dropout masks from PyTorch's generator (``nn.Dropout`` /
``F.scaled_dot_product_attention(dropout_p=...)``, reference ``src/multi_modal/mm_utils.py:52,111,114``),
whose Philox offsets a foreign kernel cannot reproduce.  The CUDA path therefore defines its OWN
counter-based stream (``csrc/philox.cuh``) and this file restates it bit for bit so that training-mode
activations/gradients can be checked against the oracle with identical masks:

* generator: Philox4x32 with ``PHILOX_ROUNDS`` rounds, key = the 64-bit step seed,
  counter = (group_lo, group_hi, site, 0);
* one call yields 16 bytes = 16 consecutive elements of one row:
  group = row * ceil(cols/16) + col // 16, byte index = col % 16 (little-endian inside each word);
* element is DROPPED iff byte < round(p * 256); kept elements are scaled by 256 / (256 - round(p*256)).

The dropout on the attention probabilities (``F.scaled_dot_product_attention(dropout_p=...)`` in the reference)
uses an interleaved variant so that one generator call covers the 16 elements one thread of the attention kernel
owns inside a 64-column block (``csrc/attention.cu``): for field row r (= (b*n_heads+h)*Sq + i) and key column j,
``blk = j // 64``, ``n = (j % 64) // 8``, ``q = (j % 8) // 2``, ``e = j % 2``; the call has
counter = (g_lo, g_hi, site, 1) with ``g = (r * ceil(Sk/64) + blk) * 4 + q`` and the element reads byte ``2*n + e``.

Deviation from the reference, stated once: probabilities are quantised to 1/256 (0.4 -> 102/256 = 0.3984,
0.2 -> 51/256 = 0.1992) and the survivors are rescaled by the matching 256/(256-t), so the estimator stays
unbiased; bit-level agreement with PyTorch's own Philox offsets is impossible for a foreign kernel.
"""
from __future__ import annotations

import numpy as np

PHILOX_ROUNDS = 7
_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def drop_threshold(p: float) -> int:
    """Byte threshold: an element is dropped iff its random byte < threshold."""
    t = int(round(float(p) * 256.0))
    return max(0, min(255, t))


def keep_scale(p: float) -> float:
    t = drop_threshold(p)
    return 256.0 / (256.0 - t)


def philox4x32(c0, c1, c2, c3, k0: int, k1: int, rounds: int = PHILOX_ROUNDS):
    """Vectorised Philox4x32; counters are uint32 arrays, key two python ints."""
    c0 = c0.astype(np.uint64)
    c1 = c1.astype(np.uint64)
    c2 = c2.astype(np.uint64)
    c3 = c3.astype(np.uint64)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def prob_random_bytes(seed: int, site: int, rows: int, cols: int) -> np.ndarray:
    """(rows, cols) uint8 field of an attention-probability dropout site (interleaved layout, see header)."""
    nblk = (cols + 63) // 64
    g = np.arange(rows * nblk * 4, dtype=np.uint64)
    c0 = (g & _MASK32).astype(np.uint32)
    c1 = (g >> np.uint64(32)).astype(np.uint32)
    c2 = np.full_like(c0, np.uint32(site & 0xFFFFFFFF))
    c3 = np.ones_like(c0)
    w = philox4x32(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    by = np.stack(w, axis=1).view(np.uint8).reshape(rows, nblk, 4, 8, 2)   # (row, blk, q, n, e)
    by = by.transpose(0, 1, 3, 2, 4).reshape(rows, nblk * 64)              # column = blk*64 + n*8 + q*2 + e
    return by[:, :cols]


def prob_keep_mask(seed: int, site: int, rows: int, cols: int, p: float) -> np.ndarray:
    if p <= 0.0:
        return np.ones((rows, cols), dtype=np.float32)
    by = prob_random_bytes(seed, site, rows, cols)
    return (by >= drop_threshold(p)).astype(np.float32) * np.float32(keep_scale(p))


def random_bytes(seed: int, site: int, rows: int, cols: int) -> np.ndarray:
    """The (rows, cols) uint8 random field of one dropout site."""
    gpr = (cols + 15) // 16
    g = np.arange(rows * gpr, dtype=np.uint64)
    c0 = (g & _MASK32).astype(np.uint32)
    c1 = (g >> np.uint64(32)).astype(np.uint32)
    c2 = np.full_like(c0, np.uint32(site & 0xFFFFFFFF))
    c3 = np.zeros_like(c0)
    w = philox4x32(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(w, axis=1)  # (G, 4) uint32, little endian bytes
    by = words.view(np.uint8).reshape(rows, gpr * 16)
    return by[:, :cols]


def keep_mask(seed: int, site: int, rows: int, cols: int, p: float) -> np.ndarray:
    """float32 (rows, cols): 0 where dropped, keep_scale(p) where kept."""
    if p <= 0.0:
        return np.ones((rows, cols), dtype=np.float32)
    by = random_bytes(seed, site, rows, cols)
    return (by >= drop_threshold(p)).astype(np.float32) * np.float32(keep_scale(p))


# Dropout site identifiers (must match csrc/philox.cuh).  site = kind + 16 * layer + 4096 * side
SITE_EMBED = 0        # + modality index in the layer slot
SITE_ATTN_PROB = 1
SITE_ATTN_OUT = 2
SITE_XATTN_PROB = 3
SITE_XATTN_OUT = 4
SITE_MLP = 5
SIDE_ENC = 0
SIDE_DEC = 1


def site_id(kind: int, layer: int, side: int) -> int:
    return kind + 16 * layer + 4096 * side


# ---------------------------------------------------------------------------------------------
# Device-side token masking (the Masker's temporal mode sampled by mmfm_mask_prep, see include/mmfm_b200.h)
# ---------------------------------------------------------------------------------------------
MASK_SITE = 8192


def mask_threshold(ratio: float) -> int:
    """32-bit threshold: element masked iff its random word < floor(ratio * 2^32)."""
    return max(0, min(0xFFFFFFFF, int(float(ratio) * 4294967296.0)))


def mask_bernoulli(seed: int, mod_index: int, B: int, T: int, ratio: float) -> np.ndarray:
    """(B,T) int64 Bernoulli(ratio) field of modality ``mod_index``: element e = b*T + t reads word (e & 3) of
    Philox(counter = (e >> 2, 0, MASK_SITE + mod_index, 2), key = seed).  Stands in for the i.i.d. field of
    models/masker.py:85-86,132 (distribution-identical; the reference draws it from the CPU mt19937 stream)."""
    n = B * T
    g = np.arange((n + 3) // 4, dtype=np.uint64)
    c0 = (g & _MASK32).astype(np.uint32)
    c1 = np.zeros_like(c0)
    c2 = np.full_like(c0, np.uint32(MASK_SITE + mod_index))
    c3 = np.full_like(c0, np.uint32(2))
    w = philox4x32(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(w, axis=1).reshape(-1)[:n]
    return (words < np.uint32(mask_threshold(ratio))).astype(np.int64).reshape(B, T)
