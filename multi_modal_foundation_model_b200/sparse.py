"""Batch assembly from the datasets' sparse spike format.

The reference stores every trial as a CSR matrix of unsigned bytes -- ``spikes_sparse_data`` / ``_indices`` /
``_indptr`` / ``_shape`` (``src/utils/dataset_utils.py:15,29-36``) -- and its loader rebuilds the dense (T, N) array of
each trial on the host with scipy, one trial at a time (``get_binned_spikes_from_sparse``, ``:38-43``), before the
fp32 (B, T, N) batch is copied to the device.  Here the host only concatenates the CSR pieces of the batch
(:func:`concat_trials`, a few numpy concatenations); the bytes and column indices go to the device as they are and
``mmfm_csr_to_dense_u8`` scatters them into a dense uint8 (B, T, N) tensor, which ``MultiModal`` accepts directly (the
uint8 wire format, expanded by ``mmfm_u8_expand``).  H2D per step at the default shape: ~9 MB instead of 69 MB.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch

from . import ops


def concat_trials(data_list: Sequence, indices_list: Sequence, indptr_list: Sequence,
                  shape_list: Sequence) -> Tuple[np.ndarray, np.ndarray, np.ndarray, Tuple[int, int, int]]:
    """Concatenate the per-trial CSR pieces of a batch: returns (data uint8 [nnz], indices int32 [nnz],
    row_ptr int64 [B*T + 1] -- global offsets of every (trial, bin) row, (B, T, N))."""
    B = len(data_list)
    if B == 0:
        raise ValueError("empty batch")
    T, N = (int(v) for v in shape_list[0])
    row_ptr = np.zeros(B * T + 1, dtype=np.int64)
    base = 0
    for i in range(B):
        if tuple(int(v) for v in shape_list[i]) != (T, N):
            raise ValueError(f"trial {i}: shape {tuple(shape_list[i])} != {(T, N)} (one session per batch)")
        ip = np.asarray(indptr_list[i], dtype=np.int64)
        if ip.shape[0] != T + 1:
            raise ValueError(f"trial {i}: indptr has {ip.shape[0]} entries, expected {T + 1}")
        row_ptr[i * T + 1:(i + 1) * T + 1] = base + ip[1:]
        base += int(ip[-1])
    data = np.concatenate([np.asarray(d, dtype=np.uint8) for d in data_list]) if base else np.zeros(0, np.uint8)
    indices = np.concatenate([np.asarray(d, dtype=np.int32) for d in indices_list]) if base else np.zeros(0, np.int32)
    if data.shape[0] != base or indices.shape[0] != base:
        raise ValueError("data / indices length does not match indptr")
    return data, indices, row_ptr, (B, T, N)


def densify(data: torch.Tensor, indices: torch.Tensor, row_ptr: torch.Tensor, shape: Tuple[int, int, int],
            out: torch.Tensor = None) -> torch.Tensor:
    """Device scatter: CSR pieces (CUDA tensors: uint8, int32, int64) -> dense uint8 (B, T, N)."""
    B, T, N = shape
    if not (data.is_cuda and indices.is_cuda and row_ptr.is_cuda):
        raise ValueError("densify runs on the device: move the CSR pieces to CUDA first (there is no host fallback)")
    if data.dtype != torch.uint8 or indices.dtype != torch.int32 or row_ptr.dtype != torch.int64:
        raise ValueError("densify: expected uint8 data, int32 indices, int64 row_ptr")
    if out is None:
        out = torch.empty(B, T, N, device=data.device, dtype=torch.uint8)
    ops.csr_to_dense_u8(data, indices, row_ptr, out, n_rows=B * T, n_cols=N)
    return out
