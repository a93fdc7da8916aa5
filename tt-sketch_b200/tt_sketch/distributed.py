"""Multi-GPU streaming sketch: shard the input, sketch the shard on this rank's GPU, combine the
partial sketches with ONE all-reduce of the packed buffer.

The reference has no distributed code (tt_sketch/sketch.py:500-503 says blocked sketching "would
only be faster in a distributed setting (which isn't properly supported)").  What makes sharding
exact is the sketch's linearity in the tensor, which the reference pins with
tests/test_sketching_matrix.py:599-631 (sketch(X1)+sketch(X2) == sketch(X1+X2)) and
tensor.py:215-234 (`SparseTensor.split`):

  * SparseTensor : contiguous nonzero ranges of equal size (`SparseTensor.split` semantics);
  * TensorSum    : summands dealt round-robin (sparse summands are additionally range-split so
                   one huge sparse term does not land on a single rank);
  * TensorTrain / CPTensor / DenseTensor given alone do not shard usefully (inputs of a few MB,
    sequential chains): rank 0 sketches them and the others contribute zeros ("replicas only").

DRMs are seed-defined, so every rank generates identical DRM entries with no communication.
One process per GPU (`torchrun`), `torch.distributed` with the NCCL backend over NVLink; the
all-reduce moves 8 * sketch_size bytes (131 MB for BASELINE config 4, < 1 MB for configs 1-3).
Host-side logic (partitioning, packing, the reduction) is backend-agnostic and is tested on CPU
with gloo at world_size 2 (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np

from tt_sketch.sketch_container import SketchContainer
from tt_sketch.tensor import SparseTensor, Tensor, TensorSum


def shard_bounds(n: int, world: int, rank: int):
    """[lo, hi) of an equal contiguous split of range(n) (remainder spread over the first ranks)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sparse(tensor: SparseTensor, world: int, rank: int) -> SparseTensor:
    lo, hi = shard_bounds(tensor.nnz, world, rank)
    idx = np.asarray(tensor.indices)[:, lo:hi]
    part = SparseTensor(tensor.shape, idx, tensor.entries[lo:hi])
    if tensor.__dict__.get("_checked"):
        part.__dict__["_checked"] = True
    return part


def shard_tensor(tensor: Tensor, world: int, rank: int) -> Optional[Tensor]:
    """The part of `tensor` this rank sketches (None: nothing -- contributes a zero sketch)."""
    if world == 1:
        return tensor
    if isinstance(tensor, SparseTensor):
        part = shard_sparse(tensor, world, rank)
        return part if part.nnz > 0 else None
    if isinstance(tensor, TensorSum):
        mine: List[Tensor] = []
        dense_like = [X for X in _flatten(tensor) if not isinstance(X, SparseTensor)]
        sparse = [X for X in _flatten(tensor) if isinstance(X, SparseTensor)]
        mine.extend(X for i, X in enumerate(dense_like) if i % world == rank)
        for X in sparse:
            part = shard_sparse(X, world, rank)
            if part.nnz > 0:
                mine.append(part)
        return TensorSum(mine, shape=tensor.shape) if mine else None
    return tensor if rank == 0 else None


def _flatten(tensor: Tensor) -> List[Tensor]:
    if isinstance(tensor, TensorSum):
        out: List[Tensor] = []
        for X in tensor.tensors:
            out.extend(_flatten(X))
        return out
    return [tensor]


def _gpu_local_sketch(part: Optional[Tensor], left_drm, right_drm, total: int):
    """Packed partial sketch of this rank's shard as a device tensor (the production path)."""
    from tt_sketch import _backend as be
    from tt_sketch.sketch_dispatch import streaming_sketch_device

    if part is None:
        return be.zeros(total)
    packed, _ = streaming_sketch_device(part, left_drm, right_drm)
    return packed


def distributed_stream_sketch(tensor: Tensor, left_drm, right_drm, group=None,
                              local_sketch: Optional[Callable] = None) -> SketchContainer:
    """Streaming sketch of `tensor` computed by all ranks of `group`; every rank returns the
    full SketchContainer.  `local_sketch(part, left_drm, right_drm, total)` must return the packed
    partial sketch as a torch tensor on the backend's device (default: this rank's GPU)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    shape = tuple(tensor.shape)
    rL, rR = tuple(left_drm.bond_rank), tuple(right_drm.bond_rank)
    _, total = SketchContainer.layout(shape, rL, rR)
    part = shard_tensor(tensor, world, rank)
    packed = (local_sketch or _gpu_local_sketch)(part, left_drm, right_drm, total)
    if not isinstance(packed, torch.Tensor):
        packed = torch.from_numpy(np.ascontiguousarray(packed, dtype=np.float64))
    if packed.numel() != total:
        raise ValueError("local sketch has the wrong packed length")
    if world > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    if packed.is_cuda:
        from tt_sketch import _backend as be

        return SketchContainer.unpack(be.to_host_pinned(packed), shape, rL, rR, copy=False)
    return SketchContainer.unpack(packed.detach().numpy(), shape, rL, rR)


def distributed_blocked_stream_sketch(tensor: Tensor, left_drm, right_drm, left_rank_slices, right_rank_slices,
                                      group=None, local_sketch: Optional[Callable] = None) -> SketchContainer:
    """blocked_stream_sketch (sketch.py:493-525) with the work spread over the ranks of `group`.
    Data (nonzeros / summands) is sharded for every block, because slices of a TT-DRM cost as much as
    the full DRM; each block is all-reduced and pasted like _assemble_blocked_stream_sketches."""
    from tt_sketch.drm_base import CanSlice
    from tt_sketch.sketch import _assemble_blocked_stream_sketches

    for drm in (left_drm, right_drm):
        if not isinstance(drm, CanSlice):
            raise ValueError(f"Blocked sketch not supported for DRM {type(drm).__name__}")
    blocks = {}
    for i, (a, b) in enumerate(zip(left_rank_slices[:-1], left_rank_slices[1:])):
        for j, (c, d) in enumerate(zip(right_rank_slices[:-1], right_rank_slices[1:])):
            blocks[(i, j)] = distributed_stream_sketch(tensor, left_drm.slice(a, b), right_drm.slice(c, d), group,
                                                       local_sketch)
    return _assemble_blocked_stream_sketches(left_rank_slices, right_rank_slices, tensor.shape, blocks)
