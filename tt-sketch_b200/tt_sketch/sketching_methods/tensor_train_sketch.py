"""Omega / Psi of a TensorTrain input.  Mirror of
tt_sketch/sketching_methods/tensor_train_sketch.py:8-35 (reference); each einsum there is one
or two strided FP64 GEMMs here.  left (r_T, rL), right (r_T', rR) as produced by
`DRM.sketch_tt`."""
from typing import Optional

import numpy as np

from tt_sketch import _backend as be
from tt_sketch.tensor import TensorTrain


def omega_tt_device(left, right, *, out, **kwargs):
    return be.gemm(left.T, right, out=out, beta=1.0)


def psi_tt_device(left, right, *, tensor: TensorTrain, mu: int, out, **kwargs):
    c = tensor.device()["cores"][mu]
    r0, n, r1 = c.shape
    if left is None:
        be.gemm(c.reshape(n, r1), right, out=out.reshape(n, -1), beta=1.0)
    elif right is None:
        be.gemm(left.T, c.reshape(r0, n), out=out.reshape(-1, n), beta=1.0)
    else:
        t1 = be.gemm(left.T, c.reshape(r0, n * r1))                  # (rL, n*r1)
        rL = t1.shape[0]
        be.gemm(t1.reshape(rL * n, r1), right, out=out.reshape(rL * n, -1), beta=1.0)
    return out


def sketch_omega_tt(left_sketch, right_sketch, **kwargs):
    L, R = be.to_device(left_sketch, np.float64), be.to_device(right_sketch, np.float64)
    return be.to_host(omega_tt_device(L, R, out=be.zeros((L.shape[1], R.shape[1]))))


def sketch_psi_tt(left_sketch: Optional[np.ndarray], right_sketch: Optional[np.ndarray], *, tensor: TensorTrain,
                  mu: int, **kwargs):
    L = be.to_device(left_sketch, np.float64) if left_sketch is not None else None
    R = be.to_device(right_sketch, np.float64) if right_sketch is not None else None
    r0, n, r1 = tensor.cores[mu].shape
    shape = (L.shape[1] if L is not None else r0, n, R.shape[1] if R is not None else r1)
    return be.to_host(psi_tt_device(L, R, tensor=tensor, mu=mu, out=be.zeros(shape)))
