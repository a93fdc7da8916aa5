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
  * DenseTensor   : slabs along the first mode, X[lo:hi]; a rank sketches its slab with the DRMs restricted
                   to it (`TensorTrainDRM.restrict_first_mode`): rows lo:hi of Psi_0 are its own, every
                   other Psi / Omega is a partial sum;
  * TensorTrain / CPTensor given alone do not shard usefully (inputs of a few MB, sequential chains):
    rank 0 sketches them and the others contribute zeros ("replicas only").

`distributed_orthogonal_sketch` shards a TensorSum the same way for `orthogonal_sketch`: Omega is reduced once,
and every Psi_mu is all-reduced BEFORE its QR (reference sketch_dispatch.py:251-271 sums the summands' Psi
before `orth_step`), i.e. d small all-reduces instead of one.

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
from tt_sketch.tensor import DenseTensor, SparseTensor, Tensor, TensorSum


class DenseSlab:
    """Rows [lo, hi) of the first mode of a DenseTensor: what one rank sketches."""

    def __init__(self, tensor: DenseTensor, lo: int, hi: int) -> None:
        self.lo, self.hi = lo, hi
        self.full_shape = tuple(tensor.shape)
        self.tensor = DenseTensor(tensor.data[lo:hi])
        self.shape = self.full_shape


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
    if isinstance(tensor, DenseTensor):
        lo, hi = shard_bounds(tensor.shape[0], world, rank)
        return DenseSlab(tensor, lo, hi) if hi > lo else None
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
    from tt_sketch.sketch_dispatch import streaming_sketch

    if part is None:
        return be.zeros(total)
    packed, _ = streaming_sketch(part, left_drm, right_drm)
    return packed


def _slab_sketch(slab: DenseSlab, left_drm, right_drm, total: int, local_sketch: Callable):
    """Packed partial sketch of a dense slab in the layout of the FULL tensor: the slab's Psi_0 block lands in
    rows lo:hi of Psi_0, every other block is a partial sum (dense_sketch.py:7-52 of the reference is linear in X).
    `local_sketch` sketches the slab as a tensor of its own with the DRMs restricted to it."""
    import torch

    for drm in (left_drm, right_drm):
        if not hasattr(drm, "restrict_first_mode"):
            raise ValueError(f"DRM {type(drm).__name__} cannot be restricted to a slab of the first mode")
    n0, rows = slab.full_shape[0], slab.hi - slab.lo
    r0 = int(right_drm.bond_rank[0])
    part = local_sketch(slab.tensor, left_drm.restrict_first_mode(slab.lo, slab.hi),
                        right_drm.restrict_first_mode(slab.lo, slab.hi), total - (n0 - rows) * r0)
    if not isinstance(part, torch.Tensor):
        part = torch.from_numpy(np.ascontiguousarray(part, dtype=np.float64))
    packed = torch.zeros(total, dtype=torch.float64, device=part.device)
    packed[slab.lo * r0:slab.hi * r0] = part[:rows * r0]
    packed[n0 * r0:] = part[rows * r0:]
    return packed


def check_same_drms(left_drm, right_drm, group=None) -> None:
    """Every rank must hold the same DRMs (kind, seed, column ranges): a DRM built with seed=None, or the
    default right seed of stream_sketch (`hash(str(d))`, randomised per process), differs between ranks and the
    all-reduce would then silently sum sketches made with different maps."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return

    def fp(d):
        cores = getattr(d, "cores", None)
        core_sum = None
        if cores is not None and len(cores) and isinstance(cores[0], np.ndarray):
            core_sum = float(sum(np.asarray(c).ravel()[:: max(1, c.size // 64)].sum() for c in cores))
        return (type(d).__name__, bool(d.transpose), int(d.seed), tuple(d.rank_min), tuple(d.rank_max),
                tuple(d.true_rank), tuple(d.shape), core_sum)

    mine = (fp(left_drm) if left_drm is not None else None, fp(right_drm))
    everyone = [None] * dist.get_world_size(group)
    dist.all_gather_object(everyone, mine, group=group)
    if any(e != everyone[0] for e in everyone):
        raise ValueError("ranks hold different DRMs (pass explicit DRMs with explicit seeds to every rank): "
                         f"{everyone}")


def distributed_stream_sketch(tensor: Tensor, left_drm, right_drm, group=None,
                              local_sketch: Optional[Callable] = None, dst: Optional[int] = None) -> Optional[SketchContainer]:
    """Streaming sketch of `tensor` computed by all ranks of `group`; every rank returns the
    full SketchContainer -- or, with `dst` given, only that rank does (one NCCL reduce instead of an all-reduce and
    ONE device->host copy of the packed sketch instead of one per rank: the ranks of a box share the host's
    memory bandwidth); the others return None.  `local_sketch(part, left_drm, right_drm, total)` must return the
    packed partial sketch as a torch tensor on the backend's device (default: this rank's GPU)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    shape = tuple(tensor.shape)
    rL, rR = tuple(left_drm.bond_rank), tuple(right_drm.bond_rank)
    _, total = SketchContainer.layout(shape, rL, rR)
    check_same_drms(left_drm, right_drm, group)
    part = shard_tensor(tensor, world, rank)
    if isinstance(part, DenseSlab):
        packed = _slab_sketch(part, left_drm, right_drm, total, local_sketch or _gpu_local_sketch)
    else:
        packed = (local_sketch or _gpu_local_sketch)(part, left_drm, right_drm, total)
    if not isinstance(packed, torch.Tensor):
        packed = torch.from_numpy(np.ascontiguousarray(packed, dtype=np.float64))
    if packed.numel() != total:
        raise ValueError("local sketch has the wrong packed length")
    if world > 1:
        if dst is None:
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(packed, dst=dst, op=dist.ReduceOp.SUM, group=group)
            if rank != dst:
                return None
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
    from tt_sketch.sketch import merged_block_drms

    merged = merged_block_drms(left_drm, right_drm, left_rank_slices, right_rank_slices)
    if merged is not None:  # one pass over the shard for all blocks (see merged_block_drms)
        return distributed_stream_sketch(tensor, merged[0], merged[1], group, local_sketch)
    blocks = {}
    for i, (a, b) in enumerate(zip(left_rank_slices[:-1], left_rank_slices[1:])):
        for j, (c, d) in enumerate(zip(right_rank_slices[:-1], right_rank_slices[1:])):
            blocks[(i, j)] = distributed_stream_sketch(tensor, left_drm.slice(a, b), right_drm.slice(c, d), group,
                                                       local_sketch)
    return _assemble_blocked_stream_sketches(left_rank_slices, right_rank_slices, tensor.shape, blocks)


def allreduce_blocked_stream_sketch(local_tensor: Tensor, left_drm, right_drm, left_rank_slices, right_rank_slices,
                                    group=None, dst: Optional[int] = None) -> Optional[SketchContainer]:
    """blocked_stream_sketch when every rank ALREADY holds its part of the data (its sparse shard, its share of
    the summands -- BASELINE configs[4]): local sketch on this rank's GPU, one NCCL reduction of the packed
    buffer.  With `dst` given only that rank receives (and copies to the host) the result; others return None."""
    import torch.distributed as dist

    from tt_sketch import _backend as be
    from tt_sketch.sketch import _assemble_blocked_stream_sketches, merged_block_drms
    from tt_sketch.sketch_dispatch import streaming_sketch as streaming_sketch_device

    on = dist.is_initialized() and dist.get_world_size(group) > 1
    check_same_drms(left_drm, right_drm, group)
    merged = merged_block_drms(left_drm, right_drm, left_rank_slices, right_rank_slices)
    pairs = [((0, 0), merged)] if merged is not None else [
        ((i, j), (left_drm.slice(a, b), right_drm.slice(c, d)))
        for i, (a, b) in enumerate(zip(left_rank_slices[:-1], left_rank_slices[1:]))
        for j, (c, d) in enumerate(zip(right_rank_slices[:-1], right_rank_slices[1:]))]
    blocks = {}
    for key, (lb, rb) in pairs:
        packed, (shape, rL, rR) = streaming_sketch_device(local_tensor, lb, rb)
        if on:
            if dst is None:
                dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
            else:
                dist.reduce(packed, dst=dst, op=dist.ReduceOp.SUM, group=group)
        if dst is None or not on or dist.get_rank(group) == dst:
            blocks[key] = SketchContainer.unpack(be.to_host_pinned(packed), shape, rL, rR, copy=False)
    if not blocks:
        return None
    if merged is not None:
        return blocks[(0, 0)]
    return _assemble_blocked_stream_sketches(left_rank_slices, right_rank_slices, local_tensor.shape, blocks)


# ------------------------------------------------------------------ orthogonal sketch of a TensorSum
class _GpuSequentialOps:
    """This rank's contributions to the orthogonal sketch, on its GPU (see sketch_dispatch._sequential_sketch)."""

    def __init__(self, part: Optional[Tensor], left_drm, right_drm, shape):
        from tt_sketch import _backend as be
        from tt_sketch.sketch_dispatch import (OrthogTTDRM, _check_supported, _summands, device_operators,
                                               get_sketch_method)

        self.be, self.shape, self.d = be, tuple(shape), len(shape)
        self.parts = _summands(part) if part is not None else []
        self.rL, self.rR = tuple(left_drm.bond_rank), tuple(right_drm.bond_rank)
        self.ops = device_operators
        for X in self.parts:
            _check_supported(X, left_drm)
            _check_supported(X, right_drm)
        self.Rc = [list(get_sketch_method(X, right_drm, device=True)(X)) for X in self.parts]
        self.Lc = [list(get_sketch_method(X, left_drm, device=True)(X)) for X in self.parts]
        self.left_psi = OrthogTTDRM(self.rL, TensorSum(self.parts, shape=self.shape)) if self.parts else None

    def omegas(self):
        out = []
        for mu in range(self.d - 1):
            o = self.be.zeros((self.rL[mu], self.rR[mu]))
            for s, X in enumerate(self.parts):
                self.ops(type(X))[0](self.Lc[s][mu], self.Rc[s][mu], tensor=X, mu=mu, out=o)
            out.append(o)
        self.Lc = None
        return out

    def psi(self, mu: int, prev_core):
        r1 = self.rL[mu - 1] if mu > 0 else 1
        r2 = self.rR[mu] if mu < self.d - 1 else 1
        P = self.be.zeros((r1, self.shape[mu], r2))
        lefts = [None] * len(self.parts)
        if mu > 0 and self.parts:
            self.left_psi.add_core(prev_core)
            lefts = next(self.left_psi)
        for s, X in enumerate(self.parts):
            self.ops(type(X))[1](lefts[s], self.Rc[s][mu] if mu < self.d - 1 else None, tensor=X, mu=mu, out=P)
        return P

    def orth(self, P, Omega):
        from tt_sketch.sketch_dispatch import orth_step_device

        return orth_step_device(P, Omega)

    def to_host(self, t):
        return self.be.to_host(t)


def distributed_orthogonal_sketch(tensor: Tensor, left_drm, right_drm, group=None, ops_factory=None):
    """`orthogonal_sketch` (reference sketch.py:81-151 -> general_sketch method="orthogonal") of a TensorSum /
    SparseTensor with the summands and nonzeros spread over the ranks of `group`.  Returns the list of
    orthogonalised TT cores (host arrays), identical on every rank.  `ops_factory(part, left, right, shape)`
    supplies the per-rank contractions (default: this rank's GPU)."""
    import torch
    import torch.distributed as dist

    on = dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    shape = tuple(tensor.shape)
    d = len(shape)
    check_same_drms(left_drm, right_drm, group)
    part = shard_tensor(tensor, world, rank)
    if isinstance(part, DenseSlab):
        raise ValueError("distributed_orthogonal_sketch shards TensorSum / SparseTensor inputs")
    ops = (ops_factory or _GpuSequentialOps)(part, left_drm, right_drm, shape)

    def reduce(ts):
        if world == 1:
            return ts
        flat = torch.cat([torch.as_tensor(t).reshape(-1) for t in ts])  # one all-reduce for the whole list
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        out, off = [], 0
        for t in ts:
            n = int(np.prod(t.shape))
            out.append(flat[off:off + n].reshape(t.shape))
            off += n
        return out

    Omega = reduce(ops.omegas())
    cores = []
    prev = None
    for mu in range(d):
        P = reduce([ops.psi(mu, prev)])[0]
        if mu < d - 1:
            P = ops.orth(P, Omega[mu])
        prev = P
        cores.append(P)
    return [np.asarray(ops.to_host(c)) for c in cores]
