"""Omega / Psi of a TuckerTensor input.  Mirror of tt_sketch/sketching_methods/tucker_sketch.py:9-46 (reference).

left (prod(s_1..s_{mu+1}), rL) and right (prod of the trailing Tucker ranks, rR) are the DRMs contracted with the
Tucker factors (`DRM.sketch_tucker`); what remains is small GEMMs against the flat core (ttsk_gemm on strided
views): Omega_mu = left^T core_(<=mu) right, Psi_mu = ((left^T x core x right) contracted with factor mu).  Like the
reference, the core is used as a FLAT C-order array against the right sketch as it arrives (whose rows run over the
trailing modes in reversed order because a right DRM contracts the mode-reversed tensor) -- reproduced, not
"fixed", so results equal the reference's."""
from typing import Optional

import numpy as np

from tt_sketch import _backend as be
from tt_sketch.tensor import TuckerTensor


def omega_tucker_device(left, right, *, tensor: TuckerTensor, mu: int, out, **kwargs):
    core = tensor.device()["core"]
    a = left.shape[0]
    w = be.gemm(left.T, core.reshape(a, -1))
    return be.gemm(w, right, out=out, beta=1.0)


def psi_tucker_device(left, right, *, tensor: TuckerTensor, mu: int, out, **kwargs):
    dev = tensor.device()
    core, U = dev["core"], dev["factors"][mu]  # U: (s_mu, n_mu)
    s, n = U.shape
    a = left.shape[0] if left is not None else 1
    b = right.shape[0] if right is not None else 1
    w = core.reshape(a, s * b)
    if left is not None:
        w = be.gemm(left.T, w)                       # (r1, s*b)
    r1 = w.shape[0]
    w = w.reshape(r1 * s, b)
    if right is not None:
        w = be.gemm(w, right)                        # (r1*s, r2)
    r2 = w.shape[1]
    w = w.reshape(r1, s, r2)
    # out[i] (n x r2) += U^T (n x s) @ w[i] (s x r2), batched over the left rank index
    be.gemm_batched(U.T.unsqueeze(0).expand(r1, n, s), w, out, beta=1.0)
    return out


def sketch_omega_tucker(left_sketch, right_sketch, *, tensor: TuckerTensor, mu: int, **kwargs):
    L, R = be.to_device(left_sketch, np.float64), be.to_device(right_sketch, np.float64)
    return be.to_host(omega_tucker_device(L, R, tensor=tensor, mu=mu, out=be.zeros((L.shape[1], R.shape[1]))))


def sketch_psi_tucker(left_sketch: Optional[np.ndarray], right_sketch: Optional[np.ndarray], *, tensor: TuckerTensor,
                      mu: int, **kwargs):
    L = be.to_device(left_sketch, np.float64) if left_sketch is not None else None
    R = be.to_device(right_sketch, np.float64) if right_sketch is not None else None
    shape = (L.shape[1] if L is not None else 1, tensor.shape[mu], R.shape[1] if R is not None else 1)
    return be.to_host(psi_tucker_device(L, R, tensor=tensor, mu=mu, out=be.zeros(shape)))
