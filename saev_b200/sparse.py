"""Sparse views of the TopK forward state.

saev's inference dump (/root/reference/src/saev/framework/inference.py:196-244) materialises the dense `out.f_x[B, d_sae]`
of every batch, copies it to the host (4.3 GB per batch at the c3 shape) and calls `scipy.sparse.csr_array` on it
(:236).  The CUDA path never forms that matrix: the forward leaves `topk_idx / topk_val [B, K]`, and the CSR block of
the batch is just those lists with every row sorted by column.
"""

from __future__ import annotations

import torch
from torch import Tensor


def topk_to_csr_parts(topk_idx: Tensor, topk_val: Tensor, d_sae: int, row_mask: Tensor | None = None):
    """(indptr int64[B + 1], indices int32[nnz], data float32[nnz]) of the matrix
    `f[b, topk_idx[b, k]] = topk_val[b, k]` in canonical CSR form (columns ascending within a row, no explicit
    zeros, empty slots `idx < 0` dropped) -- what `scipy.sparse.csr_array(f_x)` yields for the dense f_x.
    `row_mask[b] = False` empties row b (inference.py:234: `f_x[~mask_b, :] = 0.0`).  Works on any device."""
    idx = topk_idx.to(torch.int64)
    keep = (idx >= 0) & (topk_val != 0)
    if row_mask is not None:
        keep = keep & row_mask.to(keep.device)[:, None]
    key = torch.where(keep, idx, torch.full_like(idx, d_sae))  # dropped slots sort to the end of their row
    key, order = torch.sort(key, dim=1, stable=True)
    val = torch.gather(topk_val, 1, order)
    indptr = torch.zeros(idx.shape[0] + 1, dtype=torch.int64, device=idx.device)
    indptr[1:] = torch.cumsum(keep.sum(dim=1), dim=0)
    sel = key < d_sae
    return indptr, key[sel].to(torch.int32), val[sel]


def topk_to_csr(topk_idx: Tensor, topk_val: Tensor, d_sae: int, row_mask: Tensor | None = None):
    """`scipy.sparse.csr_array` of shape [B, d_sae]; only the [B, K] lists cross PCIe."""
    import scipy.sparse

    indptr, indices, data = topk_to_csr_parts(topk_idx, topk_val, d_sae, row_mask)
    return scipy.sparse.csr_array((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()),
                                  shape=(topk_idx.shape[0], d_sae))
