"""Drop-in for the hot-path-adjacent part of ``uibk/deep_preconditioning/utils.py``."""

from __future__ import annotations

import torch

from . import _lib


def sparse_matvec_mul(spconv_batch, vector_batch: torch.Tensor, transpose: bool) -> torch.Tensor:
    """Batched COO sparse matrix-vector product (utils.py:15-43), fp32, on the GPU (``dp_coo_spmv_batch_f32``).

    Args:
        spconv_batch: a batch as a ``SparseConvTensor``-like object (``features [nnz,1]``, ``indices [nnz,3]``).
        vector_batch: ``[B, N]`` vectors to multiply with the matrices.
        transpose: multiply with the transposed matrices instead.
    """
    device = vector_batch.device if vector_batch.is_cuda else torch.device("cuda")
    indices = spconv_batch.indices.to(device=device, dtype=torch.int32).contiguous()
    feats = spconv_batch.features.detach().to(device=device, dtype=torch.float32)
    feats = (feats[:, 0] if feats.dim() == 2 else feats).contiguous()
    vec = vector_batch.detach().to(device=device, dtype=torch.float32).contiguous()
    out = torch.empty_like(vec)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().dp_coo_spmv_batch_f32(_lib.ptr(indices), _lib.ptr(feats), indices.shape[0], vec.shape[0],
                                                    vec.shape[1], _lib.ptr(vec), int(bool(transpose)), _lib.ptr(out),
                                                    _lib.stream_ptr(device)), "dp_coo_spmv_batch_f32")
    return out if vector_batch.is_cuda else out.to(vector_batch.device)
