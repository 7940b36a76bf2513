"""Loss modules for ``MultiModal.loss_mod`` (reference ``mm.py:79-82``) beyond the reference's two.

``TokenCrossEntropy`` is the categorical loss BASELINE.json's north_star names for the choice / block streams.  The
reference has no categorical modality, so this is an extension defined the way the reference's
own ``forward_loss`` would run it: assigned into ``model.loss_mod[mod]`` (a plain dict, ``mm.py:79``), it is called as
``(loss_fn(preds, targets) * targets_mask).sum()`` (``mm.py:230``) and therefore returns the PER-ELEMENT field
``-targets * log_softmax(preds, -1)`` of shape (B,T,K); ``targets`` are one-hot (or class probabilities).

On the B200 path the module is never called: ``adapter.loss_kinds`` maps it to the fused ``MMFM_LOSS_CE`` kernel.  Its
``forward`` exists so that the UNMODIFIED reference model (PyTorch path) computes the same quantity -- that is the
oracle of the extension ("reference classes, re-parameterised").
"""
from __future__ import annotations

import torch
import torch.nn as nn


class TokenCrossEntropy(nn.Module):
    b200_kind = "ce"

    def forward(self, preds: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        return -(targets * torch.log_softmax(preds, dim=-1))


def one_hot_stream(classes: torch.Tensor, n_classes: int) -> torch.Tensor:
    """(B,) per-trial or (B,T) per-bin class indices -> (B,T,K) fp32 one-hot stream (a trial-level label such as the
    IBL choice / block prior, loader/base.py:325-327,447-449, is held constant over the trial's bins)."""
    if classes.dtype not in (torch.int64, torch.int32, torch.uint8):
        raise TypeError("class indices must be integers")
    if classes.dim() == 1:
        raise ValueError("per-trial labels: expand to (B,T) first, e.g. labels[:, None].expand(B, T)")
    return torch.nn.functional.one_hot(classes.long(), n_classes).to(torch.float32)
