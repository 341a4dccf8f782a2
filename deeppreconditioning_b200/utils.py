"""Drop-in for the hot-path-adjacent part of ``uibk/deep_preconditioning/utils.py``."""

from __future__ import annotations

import torch

from . import _lib


def _coo_spmv(indices: torch.Tensor, feats: torch.Tensor, vec: torch.Tensor, transpose: bool) -> torch.Tensor:
    """``dp_coo_spmv_batch_f32`` on prepared device tensors (int32 ``[nnz,3]``, fp32 ``[nnz]``, fp32 ``[B,N]``)."""
    out = torch.empty_like(vec)
    with torch.cuda.device(vec.device):
        _lib.check(_lib.lib().dp_coo_spmv_batch_f32(_lib.ptr(indices), _lib.ptr(feats), indices.shape[0], vec.shape[0],
                                                    vec.shape[1], _lib.ptr(vec), int(bool(transpose)), _lib.ptr(out),
                                                    _lib.stream_ptr(vec.device)), "dp_coo_spmv_batch_f32")
    return out


class _SparseMatvecMul(torch.autograd.Function):
    """``out[b, row] = sum feat * vec[b, col]`` with the gradients the training losses need (``metrics.py:28-29``):
    ``d/d vec`` is the transposed product (the same kernel), ``d/d feat[e] = grad[b_e, row_e] * vec[b_e, col_e]``."""

    @staticmethod
    def forward(ctx, feats, vec, indices, transpose):
        ctx.save_for_backward(feats, vec, indices)
        ctx.transpose = bool(transpose)
        return _coo_spmv(indices, feats, vec, transpose)

    @staticmethod
    def backward(ctx, grad):
        feats, vec, indices = ctx.saved_tensors
        grad = grad.contiguous()
        grad_feats = grad_vec = None
        b = indices[:, 0].long()
        r = indices[:, 2 if ctx.transpose else 1].long()
        c = indices[:, 1 if ctx.transpose else 2].long()
        inside = (r < vec.shape[1]) & (c < vec.shape[1]) & (b < vec.shape[0])
        if ctx.needs_input_grad[0]:
            grad_feats = torch.where(inside, grad[b.clamp_max(vec.shape[0] - 1), r.clamp_max(vec.shape[1] - 1)]
                                     * vec[b.clamp_max(vec.shape[0] - 1), c.clamp_max(vec.shape[1] - 1)], 0.0)
        if ctx.needs_input_grad[1]:
            grad_vec = _coo_spmv(indices, feats, grad, not ctx.transpose)
        return grad_feats, grad_vec, None, None


def sparse_matvec_mul(spconv_batch, vector_batch: torch.Tensor, transpose: bool) -> torch.Tensor:
    """Batched COO sparse matrix-vector product (utils.py:15-43), fp32, on the GPU (``dp_coo_spmv_batch_f32``).

    Differentiable with respect to the features and the vectors, like the reference's gather / ``scatter_reduce`` form
    (its callers are the training losses, ``metrics.py:28-29``).

    Args:
        spconv_batch: a batch as a ``SparseConvTensor``-like object (``features [nnz,1]``, ``indices [nnz,3]``).
        vector_batch: ``[B, N]`` vectors to multiply with the matrices.
        transpose: multiply with the transposed matrices instead.
    """
    device = vector_batch.device if vector_batch.is_cuda else torch.device("cuda")
    indices = spconv_batch.indices.to(device=device, dtype=torch.int32).contiguous()
    feats = spconv_batch.features.to(device=device, dtype=torch.float32)
    feats = (feats[:, 0] if feats.dim() == 2 else feats).contiguous()
    vec = vector_batch.to(device=device, dtype=torch.float32).contiguous()
    if torch.is_grad_enabled() and (feats.requires_grad or vec.requires_grad):
        out = _SparseMatvecMul.apply(feats, vec, indices, bool(transpose))
    else:
        out = _coo_spmv(indices, feats.detach(), vec.detach(), transpose)
    return out if vector_batch.is_cuda else out.to(vector_batch.device)
