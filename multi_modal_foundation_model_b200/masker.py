"""Mask sampling of the B200 path (mirror of the reference ``Masker``, ``src/models/masker.py``).

On the model's working path the reference throws away the masked spikes and every mask column but the first
(``_, mask = self.masker(...)``; ``mask[:, :, 0] & inputs_attn_mask`` -- ``mm.py:267,270``).  What has to be
reproduced bit for bit is therefore the **(B,T) column-0 mask** of each masking mode and -- for run-to-run stream
parity -- the amount of CPU generator state the reference burns around it (``masker.py:81,132,158,160``; the python
``random`` module for the region modes, ``:110,120``).

:func:`sample_mask_column` works on any object carrying the reference Masker's attributes (our :class:`Masker` or the
reference's own class when the drop-in of :mod:`dropin` is installed).  Three streams:

* ``'reference'``: issues the same generator calls, in the same order and with the same element counts, as
  ``models/masker.py:56-168`` so that a process seeded like the reference (``utils/utils.py:20-29``) yields identical
  masks call after call, in every mode (``temporal``, ``random_token``, ``causal``, ``neuron``, ``random``,
  ``co-smooth``, ``forward-pred``, ``inter-region``, ``intra-region``).  The 2*B*T*C discarded draws of
  ``:158,160`` are the price of that contract (0.1-0.6 s per call on the host).
* ``'fast'``: the same, minus those two discarded (B,T,C) fields.  The first call after seeding is still identical to
  the reference; later calls are not.
* ``'device'`` (the default; ``temporal`` / ``random_token`` without span expansion): nothing is drawn on
  the host at all -- ``mmfm_mask_prep`` samples the Bernoulli(ratio) field on the GPU from the step's Philox seed
  (restated in ``oracle/philox_ref.py:mask_bernoulli``).  Same distribution, different stream; other modes fall back to
  ``'fast'``.

Returns the (B,T) int64 mask; the (B,T,C) expansion of the reference is never materialised.
"""
from __future__ import annotations

import random
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .config import cfg_get

STREAMS = ("reference", "fast", "device")
TOKEN_MODES = ("temporal", "random_token", "causal")
ALL_MODES = TOKEN_MODES + ("neuron", "random", "co-smooth", "forward-pred", "inter-region", "intra-region")


def inactive(mk) -> bool:
    """The reference's early-outs (models/masker.py:62-69): nothing is masked."""
    if not mk.training and not mk.force_active:
        return True
    if mk.target_regions is None or mk.mask_regions is None:
        return True
    return mk.ratio == 0


def device_samplable(mk) -> bool:
    """True when the mask is an i.i.d. Bernoulli(ratio) field over (B,T): token modes with the span expansion off."""
    return mk.mode in ("temporal", "random_token") and (float(mk.expand_prob) == 0.0 or int(mk.max_timespan) <= 1)


def _expand(mask: torch.Tensor, width: int) -> torch.Tensor:
    """models/masker.py:170-174: a masked bin spreads over `width` neighbours (conv1d, 'same' padding)."""
    kernel = torch.ones(width).view(1, 1, -1)
    return (F.conv1d(mask.unsqueeze(1), kernel, padding="same").squeeze(1) >= 1).float()


def sample_mask_column(mk, shape: Tuple[int, int, int], neuron_regions=None, stream: str = "reference") -> torch.Tensor:
    """(B,T) int64 == ``targets_mask[:, :, 0]`` of the reference's ``Masker.forward`` for an input of ``shape``
    (B,T,C), drawn from the CPU generators exactly as the reference draws it (see the module docstring)."""
    B, T, C = shape
    if inactive(mk):
        return torch.zeros(B, T, dtype=torch.int64)
    mode = mk.mode
    if mode not in ALL_MODES:
        raise Exception(f"Masking mode {mode} not implemented")            # masker.py:129
    # stateful region bookkeeping (masker.py:72-76)
    if "all" in mk.mask_regions:
        mk.mask_regions = list(np.unique(neuron_regions))
    if "all" in mk.target_regions:
        mk.target_regions = list(np.unique(neuron_regions))

    ratio = mk.ratio
    timespan = 1
    target_cols = None                      # intra-region: (B,C) indicator of the target region
    if mode in TOKEN_MODES:                 # masker.py:79-93
        if torch.bernoulli(torch.tensor(mk.expand_prob).float()):
            timespan = int(torch.randint(1, mk.max_timespan + 1, (1,)).item())
        probs = torch.full((B, T), ratio / timespan)
        if mode == "causal":
            timespan = int(torch.randint(1, mk.max_timespan + 1, (1,)).item())
            probs = torch.full((B, T), 0.01)
    elif mode == "neuron":                  # :95-96
        probs = torch.full((B, C), ratio)
    elif mode == "random":                  # :97-98
        probs = torch.full((B, T, C), ratio)
    elif mode == "co-smooth":               # :99-103
        assert mk.channels is not None, "No channels to mask"
        probs = torch.zeros(C)
        probs[list(mk.channels)] = 1
    elif mode == "forward-pred":            # :104-108
        assert mk.timesteps is not None, "No time steps to mask"
        probs = torch.zeros(T)
        probs[list(mk.timesteps)] = 1
    else:                                   # region modes, :109-127 (python `random` picks the regions)
        assert neuron_regions is not None, "Can't mask region without brain region information"
        regions = np.asarray(neuron_regions)
        if mode == "inter-region":
            chosen = random.sample(mk.mask_regions, mk.n_mask_regions)
            probs = torch.zeros(B, C)
            for r in chosen:
                probs[torch.from_numpy(regions == r)] = 1
        else:
            chosen = random.sample(mk.target_regions, mk.n_mask_regions)
            probs = torch.ones(B, C)
            target_cols = torch.zeros(B, C)
            for r in chosen:
                idx = torch.from_numpy(regions == r)
                probs[idx] = ratio
                target_cols[idx] = 1

    field = torch.bernoulli(probs)                                           # masker.py:132 (CPU stream)

    if mode in TOKEN_MODES:                 # :135-146
        if timespan > 1:
            field = _expand(field, timespan)
        col = field                          # causal + causal_zero: the target is the field BEFORE the forward fill (:140,165)
    elif mode in ("neuron", "inter-region", "intra-region"):                  # :148-149: (B,C) broadcast over time
        col0 = field[:, 0]
        if target_cols is not None:          # :167
            col0 = col0 * target_cols[:, 0]
        col = col0[:, None].expand(B, T)
    elif mode == "co-smooth":               # :150-151
        col = field[0].expand(B, T)
    elif mode == "forward-pred":            # :152-153
        col = field[None, :].expand(B, T)
    else:                                   # random, :154-155
        col = field[:, :, 0]

    if stream == "reference":
        # masker.py:158,160: two (B,T,C) Bernoulli fields whose only surviving effect on the working path is the
        # generator state they consume
        torch.bernoulli(torch.full((B, T, C), float(mk.zero_ratio)))
        torch.bernoulli(torch.full((B, T, C), float(mk.random_ratio)))
    return col.to(torch.int64).contiguous()


class Masker(nn.Module):
    """Same constructor argument (the ``masker`` sub-config) and attribute names as the reference class
    (``models/masker.py:39-54``) so eval scripts that poke ``model.masker.ratio`` etc. (``utils/eval_utils.py:65-67``)
    keep working."""

    def __init__(self, config, stream: str = "device"):
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
        assert stream in STREAMS
        self.stream = stream

    def _inactive(self) -> bool:
        return inactive(self)

    def sample_token_mask(self, shape: Tuple[int, int, int], device, neuron_regions=None) -> torch.Tensor:
        """(B,T) int64 mask == column 0 of the reference's (B,T,C) ``targets_mask`` (host streams)."""
        stream = self.stream if self.stream != "device" else "fast"
        return sample_mask_column(self, shape, neuron_regions, stream).to(device, non_blocking=True)

    def forward(self, spikes: torch.Tensor, neuron_regions: Optional[np.ndarray] = None):
        """Reference-shaped call (``masker.py:56-60``): returns ``(spikes, mask (B,T,C) int64)``.

        The masked-spike output of the reference is dead on the model's working path (``_, mask = self.masker(...)``,
        mm.py:267); the input is returned untouched and the mask is the column-0 mask broadcast over the channels (all
        the model reads).  The device-side ``torch.rand`` of masker.py:161 draws from the CUDA generator, not the CPU
        stream, and is skipped."""
        m = self.sample_token_mask(tuple(spikes.shape), spikes.device, neuron_regions)
        return spikes, m.unsqueeze(-1).expand(spikes.shape)
