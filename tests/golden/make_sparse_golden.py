"""Golden vectors for the sparse -> dense batch assembly.

Runs the UNMODIFIED reference functions ``get_sparse_from_binned_spikes`` / ``get_binned_spikes_from_sparse``
(src/utils/dataset_utils.py:29-43, extracted from the source text: the module imports the HF ``datasets`` package, absent
here) on a seeded batch and stores the CSR lists + the dense result in ``sparse_batch.npz``."""
import ast
import os

import numpy as np
from scipy.sparse import csr_array  # noqa: F401  (used by the extracted reference code)

SRC = "/root/reference/src/utils/dataset_utils.py"
tree = ast.parse(open(SRC).read())
ns = {"np": np, "csr_array": csr_array}
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in ("get_sparse_from_binned_spikes", "get_binned_spikes_from_sparse"):
        exec(compile(ast.Module([node], []), SRC, "exec"), ns)

rng = np.random.default_rng(11)
B, T, N = 5, 12, 37
spikes = rng.poisson(0.2, size=(B, T, N)).astype(np.float32)
spikes[2, 4, :] = 0                                        # an empty row
spikes[3] = 0                                              # an empty trial
_, data, indices, indptr, shape = ns["get_sparse_from_binned_spikes"](spikes)
dense = ns["get_binned_spikes_from_sparse"](data, indices, indptr, shape)
assert np.array_equal(dense, spikes)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sparse_batch.npz")
np.savez(out, dense=dense.astype(np.uint8), n_trials=B,
         **{f"data{i}": np.asarray(data[i], dtype=np.uint8) for i in range(B)},
         **{f"indices{i}": np.asarray(indices[i], dtype=np.int32) for i in range(B)},
         **{f"indptr{i}": np.asarray(indptr[i], dtype=np.int64) for i in range(B)},
         shape=np.asarray(shape[0]))
print("wrote", out, dense.shape, int(dense.sum()))
