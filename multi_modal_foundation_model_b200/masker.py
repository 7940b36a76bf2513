"""Host-side mask sampler of the B200 path (mirror of the reference ``Masker`` for the mode the model
admits).

The reference ``MultiModal`` only accepts ``masker.mode == 'temporal'`` (``mm.py:68``) and, on its
working path, throws away the masked spikes and every mask column but the first (``mm.py:267,270``).
What has to be reproduced bit for bit is therefore the (B,T) Bernoulli field -- and, for run-to-run
stream parity, the amount of CPU generator state the reference burns around it
(``models/masker.py:81,132,158,160``: 1 + B*T + 2*B*T*C mt19937 draws per call).

Two stream modes:

* ``stream='reference'`` (default): issues the same generator calls, in the same order and with the
  same element counts, as ``models/masker.py:56-168`` so that a process seeded like the reference
  (``utils/utils.py:20-29``) yields identical masks call after call.  The 2*B*T*C discarded draws are
  the price of that contract (0.1-0.6 s per call on the host).
* ``stream='fast'``: draws only what is used (the expand-probability scalar and the (B,T) field).  The
  first call after seeding is still identical to the reference; later calls are not (documented
  divergence, used by ``bench.py`` where the mask only has to be distributed correctly).

Returns the (B,T) int64 mask; the (B,T,C) expansion of the reference is never materialised.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .config import cfg_get


class Masker(nn.Module):
    """Same constructor argument (the ``masker`` sub-config) and attribute names as the reference
    class (``models/masker.py:39-54``) so eval scripts that poke ``model.masker.ratio`` etc.
    (``utils/eval_utils.py:65-67``) keep working."""

    def __init__(self, config, stream: str = "reference"):
        super().__init__()
        self.force_active = bool(cfg_get(config, "force_active", False))
        self.mode = cfg_get(config, "mode")
        self.ratio = cfg_get(config, "ratio")
        self.zero_ratio = cfg_get(config, "zero_ratio")
        self.random_ratio = cfg_get(config, "random_ratio")
        self.expand_prob = cfg_get(config, "expand_prob")
        self.max_timespan = cfg_get(config, "max_timespan")
        self.channels = cfg_get(config, "channels")
        self.timesteps = cfg_get(config, "timesteps")
        self.mask_regions = cfg_get(config, "mask_regions")
        self.target_regions = cfg_get(config, "target_regions")
        self.n_mask_regions = cfg_get(config, "n_mask_regions")
        self.causal_zero = cfg_get(config, "causal_zero")
        self.stream = stream

    # -- reference early-outs, models/masker.py:62-69 ------------------------------------------
    def _inactive(self) -> bool:
        if not self.training and not self.force_active:
            return True
        if self.target_regions is None or self.mask_regions is None:
            return True
        return self.ratio == 0

    def sample_token_mask(self, shape: Tuple[int, int, int], device, neuron_regions=None) -> torch.Tensor:
        """(B,T) int64 mask == column 0 of the reference's (B,T,C) ``targets_mask``."""
        B, T, C = shape
        if self._inactive():
            return torch.zeros(B, T, dtype=torch.int64, device=device)
        if self.mode not in ("temporal", "random_token"):
            raise NotImplementedError(
                f"masking mode {self.mode!r}: the multi-modal model only admits 'temporal' (mm.py:68)")
        # stateful region bookkeeping of the reference (masker.py:72-76); harmless for this mode
        if neuron_regions is not None:
            if "all" in self.mask_regions:
                self.mask_regions = list(np.unique(neuron_regions))
            if "all" in self.target_regions:
                self.target_regions = list(np.unique(neuron_regions))
        # masker.py:81-86 -- one scalar draw, optional randint
        if torch.bernoulli(torch.tensor(self.expand_prob).float()):
            timespan = int(torch.randint(1, self.max_timespan + 1, (1,)).item())
        else:
            timespan = 1
        probs = torch.full((B, T), self.ratio / timespan)
        mask = torch.bernoulli(probs)                                         # masker.py:132 (CPU stream)
        if timespan > 1:                                                      # masker.py:136-137,170-174
            kernel = torch.ones(timespan).view(1, 1, -1)
            mask = (F.conv1d(mask.unsqueeze(1), kernel, padding="same").squeeze(1) >= 1).float()
        if self.stream == "reference":
            # masker.py:158,160: two (B,T,C) Bernoulli fields whose only surviving effect on the
            # working path is the generator state they consume.
            torch.bernoulli(torch.full((B, T, C), float(self.zero_ratio)))
            torch.bernoulli(torch.full((B, T, C), float(self.random_ratio)))
        return mask.to(torch.int64).to(device, non_blocking=True)

    def forward(self, spikes: torch.Tensor, neuron_regions: Optional[np.ndarray] = None):
        """Reference-shaped call (``masker.py:56-60``): returns ``(spikes, mask (B,T,C) int64)``.

        The masked-spike output of the reference is dead on the model's working path
        (``_, mask = self.masker(...)``, mm.py:267); the input is returned untouched.  The device-side
        ``torch.rand`` of masker.py:161 draws from the CUDA generator, not the CPU stream, and is
        skipped."""
        m = self.sample_token_mask(tuple(spikes.shape), spikes.device, neuron_regions)
        return spikes, m.unsqueeze(-1).expand(spikes.shape)
