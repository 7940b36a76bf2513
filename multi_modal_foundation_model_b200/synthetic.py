"""Seeded synthetic IBL-shaped trials (SURVEY.md section 8d "Synthetic inputs").

There is no network for the reference's HF datasets (``train_multi_modal.py:97-113``), so tests,
``bench.py`` and ``smoke()`` use this generator.  The batch layout is the one the reference loader hands
to the trainer (``loader/base.py:436-450``): ``spikes_data`` (B,T,N) fp32 counts, ``target`` (B,T,nb)
fp32, ``time_attn_mask`` (B,T) int64, ``spikes_timestamps`` (B,T) int64, ``neuron_regions``.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

_REGIONS = ["CA1", "DG", "LP", "PO", "VISa", "VPM", "ZI", "MRN"]


def make_batch(batch: int, n_neurons: int, n_behaviors: int = 2, n_bins: int = 100, step: int = 0,
               pad_bins: int = 0, pin: bool = False) -> Dict[str, object]:
    """One trainer batch on the host.

    spikes ~ Poisson(rate_n), rate_n = exp(N(log 0.15, 0.8^2)) clipped to [0.005, 3] per 20 ms bin;
    behaviours = unit-variance Gaussian noise low-pass filtered along time (5-bin moving average);
    ``pad_bins`` > 0 gives the right-padded variant (attention mask 0, inputs -1 there,
    ``train_multi_modal.py:125-126``)."""
    g = torch.Generator().manual_seed(1234 + step)
    rate = torch.exp(torch.randn(n_neurons, generator=g) * 0.8 + float(np.log(0.15))).clamp_(0.005, 3.0)
    spikes = torch.poisson(rate[None, None, :].expand(batch, n_bins, n_neurons).contiguous(), generator=g)
    g2 = torch.Generator().manual_seed(4321 + step)
    noise = torch.randn(batch, n_behaviors, n_bins + 4, generator=g2)
    beh = torch.nn.functional.avg_pool1d(noise, kernel_size=5, stride=1) * (5.0 ** 0.5)
    beh = beh.transpose(1, 2).contiguous()
    attn = torch.ones(batch, n_bins, dtype=torch.int64)
    if pad_bins > 0:
        attn[:, n_bins - pad_bins:] = 0
        spikes[:, n_bins - pad_bins:, :] = -1.0
        beh[:, n_bins - pad_bins:, :] = -1.0
    ts = torch.arange(n_bins, dtype=torch.int64)[None, :].expand(batch, n_bins).contiguous()
    regions = [[_REGIONS[i % len(_REGIONS)]] * batch for i in range(n_neurons)]  # loader layout: N lists of B
    out = {
        # (B, N) region matrix, converted once: the trainer redoes np.asarray(batch['neuron_regions']).T every step
        # (trainer/base.py:73), ~10 ms of pure Python at N=668 x B=256 that is outside the hot path
        "_regions_T": np.asarray(regions).T,
        "spikes_data": spikes.float(),
        "target": beh.float(),
        "time_attn_mask": attn,
        "spikes_timestamps": ts,
        "neuron_regions": regions,
        "eid": ["synthetic-session"] * batch,
    }
    if pin:
        for k, v in out.items():
            if torch.is_tensor(v):
                out[k] = v.pin_memory()
    return out


_MOD_INDEX: Dict[tuple, torch.Tensor] = {}


def _modality_index(idx: int, device) -> torch.Tensor:
    key = (str(device), idx)
    t = _MOD_INDEX.get(key)
    if t is None:
        t = _MOD_INDEX[key] = torch.tensor(idx, device=device)
    return t


def make_mod_dict(batch: Dict[str, object], avail_mod, training_mode: Optional[str], device="cpu",
                  extra_behaviors: int = 0, compact_masks: bool = False) -> Dict[str, Dict[str, object]]:
    """Build ``mod_dict`` exactly as the trainer does (``trainer/base.py:51-103``) for multi-modal
    training: ``encoding`` (ap fully masked), ``decoding`` (behavior fully masked),
    ``token_masking`` (eval_mask None -> Masker samples) or ``None`` (single-modality output).

    ``compact_masks`` replaces the trainer's dense (B,T,N) int64 all-ones / all-zeros ``eval_mask`` tensors (137 MB each
    at B=256, N=668; only column 0 is ever read, mm.py:270) by the scalars 1 / 0 the B200 path also accepts, and the
    per-step blocking copies of the modality indices by cached device scalars (no host synchronisation in the step)."""
    spikes = batch["spikes_data"].to(device, non_blocking=True)
    target = batch["target"].to(device, non_blocking=True)
    attn = batch["time_attn_mask"].to(device, non_blocking=True)
    ts = batch["spikes_timestamps"].to(device, non_blocking=True)
    mod_dict: Dict[str, Dict[str, object]] = {}
    for idx, mod in enumerate(avail_mod):
        d: Dict[str, object] = {}
        if compact_masks:
            # the trainer's `torch.tensor(idx).to(device)` (trainer/base.py:60-61) is a blocking copy from pageable memory:
            # four stream synchronisations per step, each of which drains the launch queue (tools/e2e_probe.py: host
            # enqueue time = device time, +0.25 ms per step).  The wire-format dict keeps one device scalar per index.
            d["inputs_modality"] = d["targets_modality"] = _modality_index(idx, device)
        else:
            d["inputs_modality"] = torch.tensor(idx, device=device)
            d["targets_modality"] = torch.tensor(idx, device=device)
        d["inputs_attn_mask"] = attn
        d["inputs_timestamp"] = ts
        d["targets_timestamp"] = ts
        d["eid"] = batch["eid"][0]
        d["num_neuron"] = spikes.shape[2]
        d["masking_mode"] = None
        if mod == "ap":
            d["inputs"] = spikes.clone()
            d["targets"] = spikes.clone()
            d["inputs_regions"] = batch["_regions_T"] if "_regions_T" in batch else np.asarray(batch["neuron_regions"]).T
        else:
            # 'behavior' (all nb channels) or an extra single-channel stream 'behN' (config 5 extension)
            if mod == "behavior":
                x = target
            else:
                k = int(mod[3:])
                x = target[:, :, k:k + 1]
            d["inputs"] = x.clone()
            d["targets"] = x.clone()
        if training_mode == "encoding":
            like = spikes
            d["eval_mask"] = (int(mod == "ap") if compact_masks else
                              (torch.ones_like(like) if mod == "ap" else torch.zeros_like(like)).to(torch.int64))
        elif training_mode == "decoding":
            like = target
            d["eval_mask"] = (int(mod != "ap") if compact_masks else
                              (torch.zeros_like(like) if mod == "ap" else torch.ones_like(like)).to(torch.int64))
        elif training_mode == "token_masking":
            d["eval_mask"] = None
        else:
            raise Exception("Training objective not implemented yet.")
        mod_dict[mod] = d
    return mod_dict


class DevicePrefetcher:
    """Double-buffered host->device staging of trainer batches (pinned host memory): batch i+1 is copied on a side
    stream while step i computes, so the H2D transfer leaves the critical path.  ``put`` enqueues the copy of the next
    batch, ``get`` makes the compute stream wait for it and hands the device batch over."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._next = None
        self._event = None

    def put(self, host_batch: Dict[str, object]) -> None:
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self._next = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v)
                          for k, v in host_batch.items()}
            self._event = torch.cuda.Event()
            self._event.record(self.stream)

    def get(self) -> Dict[str, object]:
        if self._next is None:
            raise RuntimeError("DevicePrefetcher.get() without a pending put()")
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._event)
        batch, self._next = self._next, None
        for v in batch.values():
            if torch.is_tensor(v):
                v.record_stream(cur)
        return batch
