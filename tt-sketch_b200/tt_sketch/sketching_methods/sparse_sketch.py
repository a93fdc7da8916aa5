"""Omega / Psi of a COO tensor from explicit per-nonzero DRM rows (operator level).

Mirror of tt_sketch/sketching_methods/sparse_sketch.py:39-69 (reference).  The reference
scans all nonzeros once per slice j (boolean mask, O(n_mu * nnz)); here the nonzeros are
bucketed by i_mu once and every slice is a register-tile accumulation in one kernel
(csrc/ttsk_sparse.cu).  `stream_sketch` does not come through here for sparse input -- it
uses the fused `ttsk_sparse_sketch` entry point that never materialises the DRM rows.
"""
from typing import Optional, Tuple

import numpy as np

from tt_sketch import _backend as be
from tt_sketch.tensor import SparseTensor


def omega_sparse_device(left, right, *, tensor: SparseTensor, out, **kwargs):
    """out (rL, rR) += (left * entries) @ right.T ; left/right are (r, nnz) device views."""
    return be.sparse_omega(tensor.device()["entries"], left, right, out)


def psi_sparse_device(left, right, *, tensor: SparseTensor, mu: int, out, **kwargs):
    """out (r1, n_mu, r2) += per-slice sums over the nonzeros with i_mu == j."""
    if left is None and right is None:
        raise ValueError("sparse Psi needs a left or a right sketch")
    dev = tensor.device()
    return be.sparse_psi(dev["indices"][mu], tensor.shape[mu], dev["entries"], left, right, out)


def sketch_omega_sparse(left_sketch, right_sketch, *, tensor: SparseTensor, **kwargs):
    L = be.to_device(left_sketch, np.float64)
    R = be.to_device(right_sketch, np.float64)
    out = be.zeros((L.shape[0], R.shape[0]))
    return be.to_host(omega_sparse_device(L, R, tensor=tensor, out=out))


def sketch_psi_sparse(left_sketch: Optional[np.ndarray], right_sketch: Optional[np.ndarray], *,
                      tensor: SparseTensor, mu: int, psi_shape: Tuple[int, int, int], **kwargs):
    d = tensor.ndim
    if (mu > 0 and left_sketch is None) or (mu < d - 1 and right_sketch is None):
        raise ValueError("missing left/right sketch for an interior core")
    L = be.to_device(left_sketch, np.float64) if (left_sketch is not None and mu > 0) else None
    R = be.to_device(right_sketch, np.float64) if (right_sketch is not None and mu < d - 1) else None
    out = be.zeros(tuple(psi_shape))
    return be.to_host(psi_sparse_device(L, R, tensor=tensor, mu=mu, out=out))
