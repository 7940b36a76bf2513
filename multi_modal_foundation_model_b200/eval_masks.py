"""Held-out masks of the evaluation modes (mirror of ``heldout_mask``, reference ``src/utils/eval_utils.py:988-1045``)
in the compact form the B200 path consumes.

The reference materialises a dense (B,T,N) int64 tensor (8 B*T*N bytes: 137 MB at B=256, N=668), multiplies the spikes
by it and hands ``1 - mask`` to the model as ``eval_mask`` -- of which ``MultiModal.forward`` reads column 0 only
(``mm.py:269-270``).  Every mode's mask is separable: a per-neuron vector (``manual``, ``most``, ``inter_region``,
``intra_region``) or a per-time-bin vector (``forward_pred``, ``modal_spike``, ``modal_behavior``).  This module returns

* ``eval_mask``   the (B,T) int64 column the model reads (accepted as is by ``engine.step``: the 2-D compact form),
* ``keep_neurons`` (N,) / ``keep_bins`` (T,) 0/1 vectors -- the separable factor of the reference's dense ``mask``,
* ``spikes``      the masked spikes (broadcast multiply by that factor; no dense int64 tensor),
* ``heldout_idxs`` exactly as the reference returns them,
* ``dense_eval_mask()`` materialises the reference's (B,T,N) tensor on demand (tests compare it bit for bit).

Caller-side helper (the eval scripts call it before ``model(mod_dict)``); nothing here is on the model's hot path.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch


def heldout_mask(spike_data: torch.Tensor, mode: str = "manual", heldout_idxs=np.array([]), n_active: int = 1,
                 target_regions: Optional[Sequence[str]] = None, neuron_regions=None) -> Dict[str, object]:
    B, T, N = spike_data.shape
    dev = spike_data.device
    keep_n = torch.ones(N, dtype=torch.int64)
    keep_t = torch.ones(T, dtype=torch.int64)
    hd = heldout_idxs
    if mode == "manual":                                           # eval_utils.py:1000-1002
        keep_n[hd] = 0
    elif mode == "most":                                           # :1004-1009
        act = spike_data.detach().float().mean(dim=(0, 1)).cpu().numpy()
        hd = np.array(np.argsort(act)[-n_active:])
        keep_n[hd] = 0
    elif mode == "inter_region":                                   # :1011-1018
        out = []
        for region in target_regions:
            idx = np.argwhere(neuron_regions == region).flatten()
            keep_n[idx] = 0
            out.append(idx[heldout_idxs])
        hd = np.stack(out).flatten()
    elif mode == "intra_region":                                   # :1020-1032
        keep_n.zero_()
        out = []
        for region in target_regions:
            idx = np.argwhere(neuron_regions == region).flatten()
            keep_n[idx] = 1
            if len(heldout_idxs) == 0:
                tgt = idx
            else:
                tgt = idx[heldout_idxs]
                keep_n[tgt] = 0
            out.append(tgt)
        hd = np.stack(out).flatten()
    elif mode in ("forward_pred", "modal_spike", "modal_behavior"):   # :1034-1040
        keep_t[hd] = 0
    else:
        raise NotImplementedError("mode not implemented")
    keep_n_d, keep_t_d = keep_n.to(dev), keep_t.to(dev)
    factor = keep_t_d[None, :, None] * keep_n_d[None, None, :]       # (1,T,N) 0/1, broadcast over the batch
    col0 = (1 - keep_t_d * keep_n_d[0])[None, :].expand(B, T).contiguous()

    def dense_eval_mask() -> torch.Tensor:
        return (1 - factor).expand(B, T, N).contiguous()

    return {"spikes": spike_data * factor.to(spike_data.dtype), "heldout_idxs": hd, "eval_mask": col0,
            "keep_neurons": keep_n_d, "keep_bins": keep_t_d, "dense_eval_mask": dense_eval_mask}
