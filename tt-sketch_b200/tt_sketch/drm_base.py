"""Dimension-reduction-map base classes: rank / slice / seed bookkeeping and orientation.

Host-side mirror of tt_sketch/drm_base.py of the reference: DRM (:14-63), DRM.T (:66-73),
CanSlice.slice (:92-109), CanIncreaseRank.increase_rank (:118-119), handle_transpose (:122-145).
Semantics kept bit-for-bit because they define output shapes and DRM seeds:
  * `rank = rank_max - rank_min` per bond; a blocked sketch slices [rank_min, rank_max) out of
    a DRM whose full size is `true_rank`;
  * a right DRM (`transpose=True`) stores true_rank / rank_min / rank_max / rank REVERSED, i.e.
    in the order of the transposed tensor's bonds; `bond_*` accessors undo that;
  * `seed` is reduced mod 2**32 - 1 (as a Python int, so NumPy >= 2 is fine).
"""
from __future__ import annotations

from abc import ABC
from copy import deepcopy
from typing import Callable, Optional, Tuple

import numpy as np

from tt_sketch.utils import TTRank, process_tt_rank


class DRM(ABC):
    rank: Tuple[int, ...]
    rank_min: Tuple[int, ...]
    rank_max: Tuple[int, ...]
    true_rank: Tuple[int, ...]
    shape: Tuple[int, ...]
    transpose: bool
    seed: int

    def __init__(self, rank: TTRank, shape: Tuple[int, ...], transpose: bool, seed: Optional[int] = None,
                 rank_min: Optional[Tuple[int, ...]] = None, rank_max: Optional[Tuple[int, ...]] = None,
                 true_rank: Optional[Tuple[int, ...]] = None, **kwargs) -> None:
        full = process_tt_rank(rank, shape, trim=False)
        nb = len(shape) - 1
        lo = tuple(rank_min) if rank_min is not None else (0,) * nb
        hi = tuple(rank_max) if rank_max is not None else full
        tr = tuple(true_rank) if true_rank is not None else full
        if transpose:  # stored in the transposed tensor's bond order
            lo, hi, tr = lo[::-1], hi[::-1], tr[::-1]
        self.transpose = bool(transpose)
        self.rank_min, self.rank_max, self.true_rank = lo, hi, tr
        self.rank = tuple(b - a for a, b in zip(lo, hi))
        self.shape = tuple(shape)
        if seed is None:
            seed = hash(np.random.uniform())
        self.seed = int(np.mod(seed, 2**32 - 1))

    # ---- user (bond) orientation, independent of `transpose`
    def _bond(self, t):
        return tuple(t[::-1]) if self.transpose else tuple(t)

    @property
    def bond_rank(self):
        return self._bond(self.rank)

    @property
    def bond_rank_min(self):
        return self._bond(self.rank_min)

    @property
    def bond_rank_max(self):
        return self._bond(self.rank_max)

    @property
    def bond_true_rank(self):
        return self._bond(self.true_rank)

    @property
    def T(self):
        other = deepcopy(self)
        other.transpose = not self.transpose
        for name in ("true_rank", "rank_min", "rank_max", "rank"):
            setattr(other, name, getattr(other, name)[::-1])
        return other

    def __repr__(self) -> str:
        side = "Right" if self.transpose else "Left"
        return f"<{side} {type(self).__name__} of rank {self.rank} and shape {self.shape} at {hex(id(self))}>"


class CanSlice(DRM):
    """The DRM can hand out the sub-DRM made of columns [start_rank, end_rank) of every bond
    (both given in user/bond order)."""

    def slice(self, start_rank, end_rank) -> DRM:
        return type(self)(rank=self.rank, shape=self.shape, transpose=self.transpose, seed=self.seed,
                          rank_min=start_rank, rank_max=end_rank, true_rank=self.bond_true_rank,
                          **self._slice_kwargs())

    def _slice_kwargs(self):
        return {}


class CanIncreaseRank(CanSlice):
    """Growing the rank keeps the existing columns (same seed, larger column range)."""

    def increase_rank(self, new_rank) -> DRM:
        return type(self)(new_rank, self.shape, self.transpose, self.seed)


def handle_transpose(sketch: Callable) -> Callable:
    """A DRM only implements the LEFT contraction.  For a right DRM the wrapper feeds it the
    mode-reversed tensor and reverses the produced list so that item mu belongs to bond mu."""

    def wrapper(self, tensor):
        if tuple(self.shape) != tuple(tensor.shape):
            raise ValueError(f"Shape {self.shape} of DRM doesn't match tensor's shape {tensor.shape}")
        if not self.transpose:
            yield from sketch(self, tensor)
        else:
            yield from list(sketch(self, tensor.T))[::-1]

    wrapper.__name__ = getattr(sketch, "__name__", "sketch")
    wrapper.__doc__ = sketch.__doc__
    return wrapper
