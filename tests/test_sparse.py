"""Sparse (CSR ubyte) -> dense batch assembly: the restatement and the host concatenation
against the reference's own round trip (golden vectors), and the device scatter against both."""
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN
from oracle import sparse_ref


def _golden():
    z = np.load(os.path.join(GOLDEN, "sparse_batch.npz"))
    B = int(z["n_trials"])
    data = [z[f"data{i}"] for i in range(B)]
    indices = [z[f"indices{i}"] for i in range(B)]
    indptr = [z[f"indptr{i}"] for i in range(B)]
    shape = [tuple(int(v) for v in z["shape"])] * B
    return z["dense"], data, indices, indptr, shape


def test_sparse_oracle_and_concat_match_reference_golden():
    from multi_modal_foundation_model_b200.sparse import concat_trials
    dense, data, indices, indptr, shape = _golden()
    assert np.array_equal(sparse_ref.binned_spikes_from_sparse(data, indices, indptr, shape), dense)
    d, idx, row_ptr, (B, T, N) = concat_trials(data, indices, indptr, shape)
    assert (B, T, N) == dense.shape and d.dtype == np.uint8 and idx.dtype == np.int32
    rebuilt = np.zeros((B * T, N), np.uint8)                 # row_ptr semantics, checked with a plain loop
    for r in range(B * T):
        rebuilt[r, idx[row_ptr[r]:row_ptr[r + 1]]] = d[row_ptr[r]:row_ptr[r + 1]]
    assert np.array_equal(rebuilt.reshape(B, T, N), dense)
    with pytest.raises(ValueError):
        concat_trials(data, indices, indptr, shape[:-1] + [(T, N + 1)])


@pytest.mark.gpu
def test_sparse_densify_on_device_and_through_the_model():
    from multi_modal_foundation_model_b200.model import build_model
    from multi_modal_foundation_model_b200.sparse import concat_trials, densify
    from multi_modal_foundation_model_b200.synthetic import make_batch, make_mod_dict
    from _util import small_config
    from scipy.sparse import csr_array
    dense, data, indices, indptr, shape = _golden()
    d, idx, row_ptr, shp = concat_trials(data, indices, indptr, shape)
    got = densify(torch.from_numpy(d).cuda(), torch.from_numpy(idx).cuda(), torch.from_numpy(row_ptr).cuda(), shp)
    assert torch.equal(got.cpu(), torch.from_numpy(dense))
    # a trainer-sized batch, then the model on the densified bytes == the model on the fp32 batch
    batch = make_batch(4, 80, 2, 100, step=5)
    sp = batch["spikes_data"].numpy()
    mats = [csr_array(sp[i], dtype=np.ubyte) for i in range(sp.shape[0])]
    d, idx, row_ptr, shp = concat_trials([m.data for m in mats], [m.indices for m in mats], [m.indptr for m in mats],
                                         [m.shape for m in mats])
    dev = densify(torch.from_numpy(d).cuda(), torch.from_numpy(idx).cuda(), torch.from_numpy(row_ptr).cuda(), shp)
    assert torch.equal(dev.cpu(), batch["spikes_data"].to(torch.uint8))
    model = build_model(80, 2, small_config()).cuda().eval()
    md32 = make_mod_dict(batch, ["ap", "behavior"], "encoding", device="cuda")
    md8 = make_mod_dict(batch, ["ap", "behavior"], "encoding", device="cuda")
    md8["ap"]["inputs"], md8["ap"]["targets"] = dev, dev.clone()
    assert model(md8).loss.item() == model(md32).loss.item()
