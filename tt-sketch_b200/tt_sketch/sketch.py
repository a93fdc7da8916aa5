"""User entry points: streaming, orthogonal, HMT and blocked sketches of a tensor, and the
assembly of a tensor train from a streaming sketch.

Mirror of tt_sketch/sketch.py (reference): hmt_sketch (:44-78), orthogonal_sketch (:81-151),
stream_sketch (:154-229), SketchedTensorTrain (:232-361), _blocked_stream_sketch_components
(:364-397), assemble_sketched_tt (:400-443), _assemble_blocked_stream_sketches (:446-473),
blocked_stream_sketch (:493-525).  Same signatures, defaults (TensorTrainDRM for every input),
rank-trimming and seed rules; the arithmetic runs on the GPU through sketch_dispatch.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Type

import numpy as np

from tt_sketch import _backend as be
from tt_sketch.drm import ALL_DRM, SparseGaussianDRM, TensorTrainDRM
from tt_sketch.drm_base import DRM, CanIncreaseRank, CanSlice
from tt_sketch.sketch_container import SketchContainer
from tt_sketch.sketch_dispatch import SketchMethod, general_sketch
from tt_sketch.sketching_methods.abstract_methods import (CansketchCP, CansketchDense, CansketchSparse,
                                                          CansketchTT)
from tt_sketch.tensor import Tensor, TensorTrain
from tt_sketch.utils import ArrayList, TTRank, process_tt_rank

BlockedSketch = Dict[Tuple[int, int], SketchContainer]


def _default_seed(seed):
    return np.mod(hash(np.random.uniform()), 2**32) if seed is None else seed


def _right_seed(seed, d: int):
    # NB: hash(str(d)) is randomised per process unless PYTHONHASHSEED is set (reference :132,:210)
    return np.mod(seed + hash(str(d)), 2**32)


def _make_drms(tensor, left_rank, right_rank, seed, left_drm_type, right_drm_type, left_drm, right_drm,
               trim_left: bool, trim_right: bool):
    """Shared DRM construction / validation of stream_sketch and orthogonal_sketch."""
    d = len(tensor.shape)
    seed = _default_seed(seed)
    if left_drm is None:
        if left_drm_type is None:
            left_drm_type = right_drm_type if right_drm_type is not None else TensorTrainDRM
        left_rank = process_tt_rank(left_rank, tensor.shape, trim=trim_left)
        left_drm = left_drm_type(left_rank, transpose=False, shape=tensor.shape, seed=seed)
    elif left_drm.rank != left_rank:
        raise ValueError(f"Left rank {left_rank} does not match the rank of the DRM {left_drm.rank}.")
    if right_drm is None:
        if right_drm_type is None:
            right_drm_type = left_drm_type if left_drm_type is not None else TensorTrainDRM
        right_rank = process_tt_rank(right_rank, tensor.shape, trim=trim_right)
        right_drm = right_drm_type(right_rank, transpose=True, shape=tensor.shape, seed=_right_seed(seed, d))
    elif tuple(right_drm.rank[::-1]) != right_rank:
        raise ValueError(f"Right rank {right_rank} does not match the rank of the DRM {right_drm.rank}.")
    return left_drm, right_drm


def hmt_sketch(tensor: Tensor, rank: TTRank, seed: Optional[int] = None, drm_type: Optional[Type[DRM]] = None,
               drm: Optional[DRM] = None, return_drm: bool = False) -> TensorTrain:
    """One-sided (Halko-Martinsson-Tropp style) sketch: right DRM only, QR after every core."""
    seed = _default_seed(seed)
    if drm is None:
        drm_type = drm_type or TensorTrainDRM
        rank = process_tt_rank(rank, tensor.shape, trim=True)
        drm = drm_type(rank, transpose=True, shape=tensor.shape, seed=seed)
    elif tuple(drm.rank[::-1]) != tuple(process_tt_rank(rank, tensor.shape, trim=False)):
        raise ValueError(f"Right rank {rank} does not match the rank of the DRM {drm.rank}.")
    sketch = general_sketch(tensor, None, drm, method=SketchMethod.hmt)
    tt = TensorTrain(sketch.Psi_cores)
    return (tt, drm) if return_drm else tt  # type: ignore[return-value]


def orthogonal_sketch(tensor: Tensor, left_rank: TTRank, right_rank: TTRank, seed: Optional[int] = None,
                      left_drm_type: Optional[Type[DRM]] = None, right_drm_type: Optional[Type[DRM]] = None,
                      left_drm: Optional[DRM] = None, right_drm: Optional[DRM] = None,
                      return_drm: bool = False) -> TensorTrain:
    """Two-sided sketch with orthogonalisation after every core (left ranks < right ranks)."""
    if not bool(np.all(np.array(left_rank) < np.array(right_rank))):
        raise ValueError(f"The right rank needs to be larger than the left rank. Left rank: {left_rank}, "
                         f"right rank: {right_rank}")
    left_drm, right_drm = _make_drms(tensor, left_rank, right_rank, seed, left_drm_type, right_drm_type, left_drm,
                                     right_drm, trim_left=True, trim_right=False)
    sketch = general_sketch(tensor, left_drm, right_drm, method=SketchMethod.orthogonal)
    tt = TensorTrain(sketch.Psi_cores)
    return (tt, left_drm, right_drm) if return_drm else tt  # type: ignore[return-value]


def stream_sketch(tensor: Tensor, left_rank: TTRank, right_rank: TTRank, seed: Optional[int] = None,
                  left_drm_type: Optional[Type[DRM]] = None, right_drm_type: Optional[Type[DRM]] = None,
                  left_drm: Optional[DRM] = None, right_drm: Optional[DRM] = None,
                  return_drm: bool = False) -> "SketchedTensorTrain":
    """Streaming (linear, one-pass) two-sided sketch.  One side's ranks must be strictly larger
    than the other's on every bond; the smaller side is trimmed to the lossless maximum."""
    l_big = bool(np.all(np.array(left_rank) > np.array(right_rank)))
    r_big = bool(np.all(np.array(left_rank) < np.array(right_rank)))
    if not (l_big or r_big):
        raise ValueError(f"Left ranks or right ranks must be conistently larger or smaller than the other. "
                         f"Left rank: {left_rank}, right rank: {right_rank}")
    left_drm, right_drm = _make_drms(tensor, left_rank, right_rank, seed, left_drm_type, right_drm_type, left_drm,
                                     right_drm, trim_left=r_big, trim_right=l_big)
    sketch = general_sketch(tensor, left_drm, right_drm, method=SketchMethod.streaming)
    stt = SketchedTensorTrain(sketch, left_drm, right_drm)
    return (stt, left_drm, right_drm) if return_drm else stt  # type: ignore[return-value]


class SketchedTensorTrain(Tensor):
    """A streaming sketch together with the DRMs that produced it: can be turned into a TT
    (`to_tt`), updated with another tensor (`+`), grown in rank (`increase_rank`)."""

    def __init__(self, sketch_: SketchContainer, left_drm: DRM, right_drm: DRM) -> None:
        self.sketch_ = sketch_
        self.left_drm = left_drm
        self.right_drm = right_drm
        self.shape = sketch_.shape

    @property
    def left_rank(self) -> Tuple[int, ...]:
        return self.left_drm.rank

    @property
    def right_rank(self) -> Tuple[int, ...]:
        return self.right_drm.rank[::-1]

    @property
    def Psi_cores(self) -> ArrayList:
        return self.sketch_.Psi_cores

    @property
    def Omega_mats(self) -> ArrayList:
        return self.sketch_.Omega_mats

    @property
    def size(self) -> int:
        return sum(a.size for a in self.Psi_cores) + sum(a.size for a in self.Omega_mats)

    def C_cores(self, direction="auto") -> ArrayList:
        return assemble_sketched_tt(self.sketch_, direction=direction)

    @property
    def T(self) -> "SketchedTensorTrain":
        return SketchedTensorTrain(self.sketch_.T, self.right_drm.T, self.left_drm.T)

    def to_tt(self) -> TensorTrain:
        return TensorTrain(self.C_cores())

    def to_numpy(self):
        return self.to_tt().to_numpy()

    def __repr__(self) -> str:
        return (f"<Sketched tensor train of shape {self.shape} with left-rank {self.left_rank} and "
                f"right-rank {self.right_rank} at {hex(id(self))}>")

    def __add__(self, other: Tensor) -> "SketchedTensorTrain":
        """Streaming update: sketch `other` with the stored DRMs and add the sketches."""
        upd = stream_sketch(other, self.left_rank, self.right_rank, left_drm=self.left_drm, right_drm=self.right_drm)
        return SketchedTensorTrain(self.sketch_ + upd.sketch_, self.left_drm, self.right_drm)

    def __mul__(self, other: float) -> "SketchedTensorTrain":
        return SketchedTensorTrain(self.sketch_ * other, self.left_drm, self.right_drm)

    def increase_rank(self, tensor: Tensor, new_left_rank: TTRank, new_right_rank: TTRank) -> "SketchedTensorTrain":
        """Grow the sketch ranks re-using the block already computed (needs CanSlice DRMs)."""
        new_left_rank = process_tt_rank(new_left_rank, tensor.shape, trim=False)
        new_right_rank = process_tt_rank(new_right_rank, tensor.shape, trim=False)
        for drm in (self.left_drm, self.right_drm):
            if not isinstance(drm, CanIncreaseRank):
                raise ValueError(f"Increasing rank is not supported for DRM {type(drm).__name__}")
        nb = len(tensor.shape) - 1
        l_slices = [(0,) * nb, tuple(self.left_drm.rank), tuple(new_left_rank)]
        r_slices = [(0,) * nb, tuple(self.right_drm.rank[::-1]), tuple(new_right_rank)]
        left = self.left_drm.increase_rank(new_left_rank)
        right = self.right_drm.increase_rank(new_right_rank)
        blocks = _blocked_stream_sketch_components(tensor, left, right, l_slices, r_slices, excluded_entries=[(0, 0)])
        blocks[(0, 0)] = self.sketch_
        sketch = _assemble_blocked_stream_sketches(l_slices, r_slices, tensor.shape, blocks)
        return SketchedTensorTrain(sketch, left, right)


def _blocked_stream_sketch_components(tensor: Tensor, left_rm: CanSlice, right_drm: CanSlice,
                                      left_rank_slices: List[Tuple[int, ...]],
                                      right_rank_slices: List[Tuple[int, ...]],
                                      excluded_entries: Optional[Sequence[Tuple[int, int]]] = None) -> BlockedSketch:
    skip = set(excluded_entries or [])
    lefts = [left_rm.slice(a, b) for a, b in zip(left_rank_slices[:-1], left_rank_slices[1:])]
    rights = [right_drm.slice(a, b) for a, b in zip(right_rank_slices[:-1], right_rank_slices[1:])]
    out: BlockedSketch = {}
    for i, lb in enumerate(lefts):
        for j, rb in enumerate(rights):
            if (i, j) not in skip:
                out[(i, j)] = general_sketch(tensor, lb, rb, method=SketchMethod.streaming)
    return out


def assemble_sketched_tt(sketch: SketchContainer, direction="auto") -> ArrayList:
    """TT cores C_mu = Psi_mu Omega_mu^+ ("right") or Omega_{mu-1}^+ Psi_mu ("left"); the
    pseudo-inverse of the small Omega is a Jacobi SVD on the GPU with gelsd's cut-off, applied to
    the r*n right-hand sides by GEMM."""
    if direction == "auto":
        bigger = np.all(np.array(sketch.left_rank) > np.array(sketch.right_rank))
        direction = "left" if bigger else "right"
    cores: ArrayList = []
    if direction == "right":
        for Psi, Omega in zip(sketch.Psi_cores[:-1], sketch.Omega_mats):
            r1, n, r2 = Psi.shape
            c = be.gemm(be.to_device(Psi.reshape(r1 * n, r2), np.float64), be.pinv(be.to_device(Omega, np.float64)))
            cores.append(be.to_host(c).reshape(r1, n, Omega.shape[0]))
        cores.append(sketch.Psi_cores[-1])
    elif direction == "left":
        cores.append(sketch.Psi_cores[0])
        for Psi, Omega in zip(sketch.Psi_cores[1:], sketch.Omega_mats):
            r1, n, r2 = Psi.shape
            c = be.gemm(be.pinv(be.to_device(Omega, np.float64)), be.to_device(Psi.reshape(r1, n * r2), np.float64))
            cores.append(be.to_host(c).reshape(Omega.shape[1], n, r2))
    else:
        raise ValueError(f"Unknown direction {direction}")
    return cores


def _assemble_blocked_stream_sketches(left_rank_slices: List[Tuple[int, ...]],
                                      right_rank_slices: List[Tuple[int, ...]], shape: Tuple[int, ...],
                                      sketch_dict: BlockedSketch) -> SketchContainer:
    """Paste block (i, j) into rows [l_i, l_{i+1}) x columns [r_j, r_{j+1}) of every Psi / Omega."""
    full = SketchContainer.zero(shape, tuple(left_rank_slices[-1]), tuple(right_rank_slices[-1]))
    d = len(shape)
    for (i, j), blk in sketch_dict.items():
        l0, l1 = (0,) + tuple(left_rank_slices[i]), (1,) + tuple(left_rank_slices[i + 1])
        r0, r1 = tuple(right_rank_slices[j]) + (0,), tuple(right_rank_slices[j + 1]) + (1,)
        for mu in range(d):
            full.Psi_cores[mu][l0[mu]:l1[mu], :, r0[mu]:r1[mu]] = blk.Psi_cores[mu]
        for mu in range(d - 1):
            full.Omega_mats[mu][l0[mu + 1]:l1[mu + 1], r0[mu]:r1[mu]] = blk.Omega_mats[mu]
    return full


def get_drm_capabilities():
    caps = (CanSlice, CanIncreaseRank, CansketchSparse, CansketchDense, CansketchTT, CansketchCP)
    return {drm.__name__: {c.__name__: issubclass(drm, c) for c in caps} for drm in ALL_DRM}


def blocked_stream_sketch(tensor: Tensor, left_drm: CanSlice, right_drm: CanSlice,
                          left_rank_slices: List[Tuple[int, ...]],
                          right_rank_slices: List[Tuple[int, ...]]) -> SketchContainer:
    """Streaming sketch computed block by block over slices of the DRM ranks; equals the
    unblocked sketch.  Blocks are independent (tt_sketch.distributed can spread them or the
    nonzeros over GPUs)."""
    for drm in (left_drm, right_drm):
        if not isinstance(drm, CanSlice):
            raise ValueError(f"Blocked sketch not supported for DRM {type(drm).__name__}")
    merged = merged_block_drms(left_drm, right_drm, left_rank_slices, right_rank_slices)
    if merged is not None:
        return general_sketch(tensor, merged[0], merged[1], method=SketchMethod.streaming)
    blocks = _blocked_stream_sketch_components(tensor, left_drm, right_drm, left_rank_slices, right_rank_slices)
    return _assemble_blocked_stream_sketches(left_rank_slices, right_rank_slices, tensor.shape, blocks)


def merged_block_drms(left_drm: CanSlice, right_drm: CanSlice, left_rank_slices, right_rank_slices):
    """Block (i, j) of a blocked sketch is the streaming sketch under the column slices
    [l_i, l_{i+1}) x [r_j, r_{j+1}) of the DRMs, pasted into those rows / columns (reference sketch.py:364-397,
    446-473), so the assembled result IS the streaming sketch under the slices [l_0, l_last) x [r_0, r_last).
    For the library's own DRMs -- whose slices are column ranges of one fixed map: a TensorTrainDRM slice
    recomputes the whole chain and keeps some columns (reference tensor_train_drm.py:60-69), which is why the
    reference pays blocks x full cost -- the chain is therefore computed ONCE and every block's columns come from
    it.  Returns the two merged DRMs, or None when the per-block path must run (third-party DRM classes, slices
    that are not consecutive, merged rank above the fused kernels' 64)."""
    from tt_sketch.drm import SparseGaussianDRM, TensorTrainDRM

    if type(left_drm) not in (SparseGaussianDRM, TensorTrainDRM) or type(right_drm) not in (SparseGaussianDRM, TensorTrainDRM):
        return None
    ls, rs = [tuple(x) for x in left_rank_slices], [tuple(x) for x in right_rank_slices]
    if len(ls) < 2 or len(rs) < 2:
        return None
    for sl in (ls, rs):
        for a, b in zip(sl[:-1], sl[1:]):
            if any(y < x for x, y in zip(a, b)):
                return None
    if max(max(y - x for x, y in zip(ls[0], ls[-1])), max(y - x for x, y in zip(rs[0], rs[-1]))) > 64:
        return None
    return left_drm.slice(ls[0], ls[-1]), right_drm.slice(rs[0], rs[-1])
