"""Omega / Psi of a CPTensor input.  Mirror of tt_sketch/sketching_methods/cp_sketch.py:6-36
(reference).  The three-operand Hadamard einsum "ij,kj,jm->ikm" (which NumPy evaluates without
BLAS) becomes a Khatri-Rao operand kernel followed by one GEMM.  left (R_cp, rL),
right (R_cp, rR) as produced by `DRM.sketch_cp`."""
from typing import Optional

import numpy as np

from tt_sketch import _backend as be
from tt_sketch.tensor import CPTensor


def omega_cp_device(left, right, *, out, **kwargs):
    return be.gemm(left.T, right, out=out, beta=1.0)


def psi_cp_device(left, right, *, tensor: CPTensor, mu: int, out, **kwargs):
    a = tensor.device()["cores"][mu]  # (n, R)
    n, R = a.shape
    if left is None:
        be.gemm(a, right, out=out.reshape(n, -1), beta=1.0)
    elif right is None:
        be.gemm(left.T, a.T, out=out.reshape(-1, n), beta=1.0)
    else:
        kr = be.khatri_rao(a, right)                                  # kr[j, k, m] = a[k, j] right[j, m]
        rR = kr.shape[2]
        be.gemm(left.T, kr.reshape(R, n * rR), out=out.reshape(-1, n * rR), beta=1.0)
    return out


def sketch_omega_cp(left_sketch, right_sketch, **kwargs):
    L, R = be.to_device(left_sketch, np.float64), be.to_device(right_sketch, np.float64)
    return be.to_host(omega_cp_device(L, R, out=be.zeros((L.shape[1], R.shape[1]))))


def sketch_psi_cp(left_sketch: Optional[np.ndarray], right_sketch: Optional[np.ndarray], *, tensor: CPTensor,
                  mu: int, **kwargs):
    L = be.to_device(left_sketch, np.float64) if left_sketch is not None else None
    R = be.to_device(right_sketch, np.float64) if right_sketch is not None else None
    n = tensor.cores[mu].shape[0]
    shape = (L.shape[1] if L is not None else 1, n, R.shape[1] if R is not None else 1)
    return be.to_host(psi_cp_device(L, R, tensor=tensor, mu=mu, out=be.zeros(shape)))
