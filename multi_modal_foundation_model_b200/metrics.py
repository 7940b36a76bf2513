"""Evaluation metrics on the device.

The reference evaluates its models with bits-per-spike and R^2 computed on the host, one neuron / channel at a time
(``src/utils/eval_utils.py:201,300,405,608,849`` call ``bits_per_spike`` in a Python loop over neurons;
``compute_R2_main``, ``:1539-1549``, calls ``r2_score`` per channel).  Here ONE kernel pass (``mmfm_column_stats``)
reduces the predictions and targets to four float64 sums per column, from which every one of those numbers follows:

    s_y = sum y      s_yy = sum y^2      s_err = sum (y - p)^2      s_nll = sum (rate - y * log rate)

with ``rate = exp(p)`` when the predictions are log-rates (the model's ``mod_preds['ap']``; ``trainer/base.py:231``
exponentiates them before the metric).  ``lgamma(y + 1)`` cancels between the model and the null likelihood and is
never computed.  There is no host fallback.
"""
from __future__ import annotations

import math

import torch

from . import ops


def column_stats(pred: torch.Tensor, y: torch.Tensor, log_rate: bool) -> torch.Tensor:
    """pred, y: (..., C) float32 CUDA tensors of equal shape -> (4, C) float64 [s_y, s_yy, s_err, s_nll]."""
    if pred.shape != y.shape or pred.dtype != torch.float32 or y.dtype != torch.float32 or not pred.is_cuda:
        raise ValueError("column_stats: float32 CUDA tensors of equal shape expected")
    C = pred.shape[-1]
    p2, y2 = pred.reshape(-1, C).contiguous(), y.reshape(-1, C).contiguous()
    out = torch.empty(4, C, device=pred.device, dtype=torch.float64)
    ops.column_stats(p2, y2, out, log_rate=log_rate)
    return out


def bits_per_spike_per_neuron(log_rates: torch.Tensor, spikes: torch.Tensor) -> torch.Tensor:
    """Per-neuron bits per spike (eval_utils.py:1095-1119 applied to one neuron at a time); (N,) float64.
    A neuron without spikes gives +-inf / nan exactly as the reference's division by zero does."""
    st = column_stats(log_rates, spikes, True)
    R = float(spikes.numel() // spikes.shape[-1])
    s_y, s_nll = st[0], st[3]
    mean = s_y / R
    null_rate = torch.where(mean == 0, torch.full_like(mean, 1e-9), mean)      # eval_utils.py:1086-1091
    nll_null = R * null_rate - s_y * torch.log(null_rate)
    return (nll_null - s_nll) / s_y / math.log(2.0)


def bits_per_spike(log_rates: torch.Tensor, spikes: torch.Tensor) -> torch.Tensor:
    """Population bits per spike over all neurons (eval_utils.py:514 form); 0-d float64."""
    st = column_stats(log_rates, spikes, True)
    R = float(spikes.numel() // spikes.shape[-1])
    s_y, s_nll = st[0], st[3]
    mean = s_y / R
    null_rate = torch.where(mean == 0, torch.full_like(mean, 1e-9), mean)
    nll_null = (R * null_rate - s_y * torch.log(null_rate)).sum()
    return (nll_null - s_nll.sum()) / s_y.sum() / math.log(2.0)


def r2_per_channel(pred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Coefficient of determination per channel of the last axis (sklearn ``r2_score`` semantics); (C,) float64."""
    st = column_stats(pred, y, False)
    R = float(y.numel() // y.shape[-1])
    ss_tot = st[1] - st[0] * st[0] / R
    return 1.0 - st[2] / ss_tot
