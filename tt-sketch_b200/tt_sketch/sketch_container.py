"""Output record of a sketch: the d order-3 cores Psi_mu and the d-1 matrices Omega_mu.

Same attributes as tt_sketch/sketch_container.py:11-89 of the reference (Psi_cores, Omega_mats,
shape, left_rank, right_rank, zero, +, .T).  Scalar multiplication works here (the reference's
`__mul__` raises UnboundLocalError, sketch_container.py:78).  `pack` / `unpack` give the flat
layout [Psi_0|...|Psi_{d-1}|Omega_0|...|Omega_{d-2}] shared with libttsk and with the
multi-GPU all-reduce.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from tt_sketch.utils import ArrayList


class SketchContainer:
    def __init__(self, Psi_cores: ArrayList, Omega_mats: ArrayList, shape: Optional[Tuple[int, ...]] = None,
                 left_rank: Optional[Tuple[int, ...]] = None, right_rank: Optional[Tuple[int, ...]] = None) -> None:
        self.Psi_cores = Psi_cores
        self.Omega_mats = Omega_mats
        self.shape = tuple(P.shape[1] for P in Psi_cores) if shape is None else shape
        self.left_rank = tuple(P.shape[0] for P in Psi_cores[1:]) if left_rank is None else left_rank
        self.right_rank = tuple(P.shape[2] for P in Psi_cores[:-1]) if right_rank is None else right_rank

    @staticmethod
    def layout(shape, left_rank, right_rank):
        """[(offset, shape)] of every Psi then every Omega in the packed buffer, and its length."""
        d = len(shape)
        items, off = [], 0
        for mu in range(d):
            shp = (left_rank[mu - 1] if mu > 0 else 1, shape[mu], right_rank[mu] if mu < d - 1 else 1)
            items.append((off, shp))
            off += int(np.prod(shp))
        for mu in range(d - 1):
            shp = (left_rank[mu], right_rank[mu])
            items.append((off, shp))
            off += int(np.prod(shp))
        return items, off

    @classmethod
    def zero(cls, shape, left_rank, right_rank) -> "SketchContainer":
        items, _ = cls.layout(shape, left_rank, right_rank)
        d = len(shape)
        arrays = [np.zeros(shp) for _, shp in items]
        return cls(arrays[:d], arrays[d:], tuple(shape), tuple(left_rank), tuple(right_rank))

    @classmethod
    def unpack(cls, flat: np.ndarray, shape, left_rank, right_rank, copy: bool = True) -> "SketchContainer":
        """Split a packed buffer into Psi cores and Omega matrices.  With copy=False the arrays are
        views of `flat` (which must then be a fresh buffer owned by nobody else)."""
        items, total = cls.layout(shape, left_rank, right_rank)
        if flat.size != total:
            raise ValueError("packed sketch has the wrong length")
        d = len(shape)
        arrays = [flat[o:o + int(np.prod(s))].reshape(s) for o, s in items]
        if copy:
            arrays = [a.copy() for a in arrays]
        return cls(arrays[:d], arrays[d:])

    def pack(self) -> np.ndarray:
        return np.concatenate([a.reshape(-1) for a in list(self.Psi_cores) + list(self.Omega_mats)])

    def __add__(self, other: "SketchContainer") -> "SketchContainer":
        return SketchContainer([a + b for a, b in zip(self.Psi_cores, other.Psi_cores)],
                               [a + b for a, b in zip(self.Omega_mats, other.Omega_mats)])

    @property
    def T(self) -> "SketchContainer":
        return SketchContainer([P.transpose(2, 1, 0) for P in reversed(self.Psi_cores)],
                               [O.T for O in reversed(self.Omega_mats)])

    def __mul__(self, other: float) -> "SketchContainer":
        return SketchContainer([P * other for P in self.Psi_cores], [O * other for O in self.Omega_mats])

    __rmul__ = __mul__

    def __neg__(self):
        return self * -1

    def __sub__(self, other):
        return self + (-other)

    def __truediv__(self, other: float):
        return self * (1 / other)
