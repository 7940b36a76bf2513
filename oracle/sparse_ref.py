"""TEST INFRASTRUCTURE ONLY -- numpy / scipy restatement of the reference's sparse -> dense batch assembly
(``get_binned_spikes_from_sparse``, src/utils/dataset_utils.py:38-43: one ``csr_array((data, indices, indptr),
shape).toarray()`` per trial, stacked).  Parity status: PINNED by ``tests/golden/sparse_batch.npz`` (produced by
``tests/golden/make_sparse_golden.py`` from the unmodified reference functions)."""
import numpy as np
from scipy.sparse import csr_array


def binned_spikes_from_sparse(data_list, indices_list, indptr_list, shape_list):
    return np.array([csr_array((data_list[i], indices_list[i], indptr_list[i]), shape=tuple(shape_list[i])).toarray()
                     for i in range(len(data_list))])
