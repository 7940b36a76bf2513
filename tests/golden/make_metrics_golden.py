"""Golden vectors for the evaluation metrics.

Runs the UNMODIFIED reference functions ``neg_log_likelihood`` / ``bits_per_spike`` (src/utils/eval_utils.py:1052-1119,
extracted from the source text because the module itself imports plotting packages that are absent here) and
``sklearn.metrics.r2_score`` (what ``compute_R2_main``, eval_utils.py:1539-1549, calls per channel) on seeded inputs
and stores inputs + outputs in ``metrics.npz``.  Re-run only in the build container (needs /root/reference)."""
import ast
import logging
import os

import numpy as np
from scipy.special import gammaln  # noqa: F401  (used by the extracted reference code)
from sklearn.metrics import r2_score

SRC = "/root/reference/src/utils/eval_utils.py"
tree = ast.parse(open(SRC).read())
ns = {"np": np, "gammaln": gammaln, "logger": logging.getLogger("ref")}
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in ("neg_log_likelihood", "bits_per_spike"):
        exec(compile(ast.Module([node], []), SRC, "exec"), ns)

rng = np.random.default_rng(7)
B, T, N = 6, 10, 7
rate = np.exp(rng.normal(np.log(0.3), 0.7, size=N)).astype(np.float32)
spikes = rng.poisson(rate[None, None, :], size=(B, T, N)).astype(np.float32)
spikes[:, :, 3] = 0.0                                     # a silent neuron (null rate 0 -> 1e-9 in the reference)
log_rates = (np.log(rate)[None, None, :] + 0.3 * rng.normal(size=(B, T, N))).astype(np.float32)
rates = np.exp(log_rates)
bps_all = ns["bits_per_spike"](rates.copy(), spikes.copy())
with np.errstate(divide="ignore", invalid="ignore"):
    bps_n = np.array([ns["bits_per_spike"](rates[:, :, [n]].copy(), spikes[:, :, [n]].copy()) for n in range(N)])
beh = rng.normal(size=(B, T, 2)).astype(np.float32)
beh_pred = (beh + 0.4 * rng.normal(size=beh.shape)).astype(np.float32)
r2 = np.array([r2_score(beh[:, :, c].flatten(), beh_pred[:, :, c].flatten()) for c in range(2)])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics.npz")
np.savez(out, spikes=spikes, log_rates=log_rates, bps_all=bps_all, bps_n=bps_n, beh=beh, beh_pred=beh_pred, r2=r2)
print("wrote", out, "bps", bps_all, bps_n, "r2", r2)
