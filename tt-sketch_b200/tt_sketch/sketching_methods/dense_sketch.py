"""Omega / Psi of a DenseTensor input.  Mirror of
tt_sketch/sketching_methods/dense_sketch.py:7-52 (reference).  left (rL, prod(shape[:mu+1])),
right (rR, prod(shape[mu+1:])) are the dense DRM unfoldings from `DRM.sketch_dense`, used as
flat arrays against the C-order unfolding of X exactly like the reference (including its
reversed-mode right unfolding).  X is streamed once per bond through the
split-K GEMM  Y = X^{<mu>} R^T ; the small left factor is applied to Y."""
from typing import Optional

import numpy as np

from tt_sketch import _backend as be
from tt_sketch.tensor import DenseTensor


def _x(tensor: DenseTensor):
    x = tensor.device()["data"]
    return x if x.is_contiguous() else x.contiguous()


def omega_dense_device(left, right, *, tensor: DenseTensor, mu: int, out, **kwargs):
    x = _x(tensor)
    rows = int(np.prod(tensor.shape[: mu + 1]))
    y = be.gemm(x.reshape(rows, -1), right.T)                         # (P, rR)
    return be.gemm(left, y, out=out, beta=1.0)


def psi_dense_device(left, right, *, tensor: DenseTensor, mu: int, out, **kwargs):
    x = _x(tensor)
    n = tensor.shape[mu]
    if left is None:
        be.gemm(x.reshape(n, -1), right.T, out=out.reshape(n, -1), beta=1.0)
    elif right is None:
        be.gemm(left, x.reshape(-1, n), out=out.reshape(-1, n), beta=1.0)
    else:
        pre = int(np.prod(tensor.shape[:mu]))
        y = be.gemm(x.reshape(pre * n, -1), right.T)                  # (pre*n, rR)
        rR = y.shape[1]
        be.gemm(left, y.reshape(pre, n * rR), out=out.reshape(-1, n * rR), beta=1.0)
    return out


def sketch_omega_dense(left_sketch, rigth_sketch, *, tensor: DenseTensor, mu: int, **kwargs):
    L, R = be.to_device(left_sketch, np.float64), be.to_device(rigth_sketch, np.float64)
    out = be.zeros((L.shape[0], R.shape[0]))
    return be.to_host(omega_dense_device(L, R, tensor=tensor, mu=mu, out=out))


def sketch_psi_dense(left_sketch: Optional[np.ndarray], right_sketch: Optional[np.ndarray], *,
                     tensor: DenseTensor, mu: int, **kwargs):
    L = be.to_device(left_sketch, np.float64) if left_sketch is not None else None
    R = be.to_device(right_sketch, np.float64) if right_sketch is not None else None
    shape = (L.shape[0] if L is not None else 1, tensor.shape[mu], R.shape[0] if R is not None else 1)
    return be.to_host(psi_dense_device(L, R, tensor=tensor, mu=mu, out=be.zeros(shape)))
