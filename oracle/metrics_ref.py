"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's evaluation metrics.

Parity status: PINNED by ``tests/golden/metrics.npz`` (produced by ``tests/golden/make_metrics_golden.py`` from the
unmodified reference functions).  Nothing in the product package imports this file.

* ``bits_per_spike`` / ``neg_log_likelihood``: ``src/utils/eval_utils.py:1052-1119`` (copied there from the NLB tools):
  Poisson NLL ``r - n log r + lgamma(n + 1)`` of the model rates against the per-neuron mean-rate null model, in bits
  per spike; zero rates are replaced by 1e-9; the callers evaluate it one neuron at a time
  (``eval_utils.py:201,300,405,608,849``).
* ``r2``: ``sklearn.metrics.r2_score`` as called per channel by ``compute_R2_main`` (``eval_utils.py:1539-1549``).
"""
import numpy as np
from scipy.special import gammaln


def neg_log_likelihood(rates, spikes):
    rates = np.where(rates == 0, 1e-9, rates).astype(np.float64)
    spikes = spikes.astype(np.float64)
    return np.sum(rates - spikes * np.log(rates) + gammaln(spikes + 1.0))


def bits_per_spike(rates, spikes):
    """rates, spikes: (..., N).  Null model = per-neuron mean rate over all leading axes."""
    nll_model = neg_log_likelihood(rates, spikes)
    null = np.broadcast_to(spikes.mean(axis=tuple(range(spikes.ndim - 1)), keepdims=True), spikes.shape)
    nll_null = neg_log_likelihood(null, spikes)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (nll_null - nll_model) / np.sum(spikes, dtype=np.float64) / np.log(2)


def bits_per_spike_per_neuron(rates, spikes):
    return np.array([bits_per_spike(rates[..., [n]], spikes[..., [n]]) for n in range(spikes.shape[-1])])


def r2(y, y_pred):
    """Per channel of the last axis: 1 - SS_res / SS_tot."""
    y = y.reshape(-1, y.shape[-1]).astype(np.float64)
    p = y_pred.reshape(-1, y.shape[-1]).astype(np.float64)
    return 1.0 - ((y - p) ** 2).sum(0) / ((y - y.mean(0)) ** 2).sum(0)
