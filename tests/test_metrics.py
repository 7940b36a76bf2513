"""Evaluation metrics: the numpy restatement against the reference's own functions
(golden vectors), and the device path (mmfm_column_stats + metrics.py) against the restatement."""
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN
from oracle import metrics_ref as mr


def _golden():
    return np.load(os.path.join(GOLDEN, "metrics.npz"))


def test_metrics_oracle_matches_reference_golden():
    # the reference computes in the dtype of its inputs (float32 arrays from the model); the restatement in float64
    z = _golden()
    rates = np.exp(z["log_rates"])
    assert np.isclose(mr.bits_per_spike(rates, z["spikes"]), z["bps_all"], rtol=2e-6)
    got = mr.bits_per_spike_per_neuron(rates, z["spikes"])
    fin = np.isfinite(z["bps_n"])
    assert np.allclose(got[fin], z["bps_n"][fin], rtol=2e-5, atol=1e-7) and np.array_equal(np.isfinite(got), fin)
    assert np.allclose(mr.r2(z["beh"], z["beh_pred"]), z["r2"], rtol=1e-6)


@pytest.mark.gpu
def test_metrics_device_matches_golden_and_oracle():
    from multi_modal_foundation_model_b200 import metrics
    z = _golden()
    lr, sp = torch.from_numpy(z["log_rates"]).cuda(), torch.from_numpy(z["spikes"]).cuda()
    bps_n = metrics.bits_per_spike_per_neuron(lr, sp).cpu().numpy()
    fin = np.isfinite(z["bps_n"])
    assert np.allclose(bps_n[fin], z["bps_n"][fin], rtol=2e-5, atol=1e-6) and np.array_equal(np.isfinite(bps_n), fin)
    assert np.isclose(metrics.bits_per_spike(lr, sp).item(), float(z["bps_all"]), rtol=2e-5)
    r2 = metrics.r2_per_channel(torch.from_numpy(z["beh_pred"]).cuda(), torch.from_numpy(z["beh"]).cuda()).cpu().numpy()
    assert np.allclose(r2, z["r2"], rtol=1e-5)
    # a session-sized batch: 256 trials x 100 bins x 668 neurons, against the restatement
    g = torch.Generator().manual_seed(3)
    rate = torch.exp(torch.randn(668, generator=g) * 0.8 - 1.9)
    spikes = torch.poisson(rate.expand(256, 100, 668).contiguous(), generator=g)
    logr = torch.log(rate)[None, None, :] + 0.2 * torch.randn(256, 100, 668, generator=g)
    got = metrics.bits_per_spike_per_neuron(logr.cuda(), spikes.cuda()).cpu().numpy()
    ref = mr.bits_per_spike_per_neuron(np.exp(logr.numpy().astype(np.float64)), spikes.numpy())
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin) and np.allclose(got[fin], ref[fin], rtol=1e-4, atol=1e-5)
