"""Host-side rank bookkeeping and RNG helpers (kept on the host on purpose: they define output
shapes and the TT-DRM core values, which must match the reference value-for-value).

Mirrors the parts of tt_sketch/utils.py (reference) the sketching path needs:
process_tt_rank (:155-175), trim_ranks (:121-152), random_normal (:178-227), matricize (:63-82),
right_mul_pinv / left_mul_pinv (:98-109, here evaluated on the GPU).
"""
from __future__ import annotations

import math
import multiprocessing
from typing import Generator, List, Optional, Sequence, Tuple, Union

import numpy as np
import numpy.typing as npt

ArrayList = List[npt.NDArray[np.float64]]
ArrayGenerator = Generator[npt.NDArray[np.float64], None, None]
TTRank = Union[int, Tuple[int, ...]]


def trim_ranks(dims: Tuple[int, ...], ranks: Tuple[int, ...]) -> Tuple[int, ...]:
    """Largest TT rank <= `ranks` a tensor of shape `dims` can have without redundancy: every
    bond is capped by the product of the modes on either side, then neighbouring bonds are
    relaxed until r_{k+1} <= r_k * n and r_k <= n * r_{k+1} hold everywhere."""
    d = len(dims)
    r = [1] * (d + 1)
    for k in range(d - 1):
        left = math.prod(dims[: k + 1])
        right = math.prod(dims[k + 1:])
        r[k + 1] = min(ranks[k], left, right)
    for _ in range(100):
        moved = False
        for k, n in enumerate(dims):
            if r[k + 1] > r[k] * n:
                r[k + 1], moved = r[k] * n, True
            if r[k] > n * r[k + 1]:
                r[k], moved = n * r[k + 1], True
        if not moved:
            break
    return tuple(r[1:-1])


def process_tt_rank(rank: TTRank, shape: Tuple[int, ...], trim: bool) -> Tuple[int, ...]:
    """Normalise an int / iterable TT rank to a tuple of len(shape)-1 entries."""
    try:
        out = tuple(rank)  # type: ignore[arg-type]
    except TypeError:
        out = (rank,) * (len(shape) - 1)  # type: ignore[assignment]
    if len(out) != len(shape) - 1:
        raise ValueError(f"TT-rank {out} doesn't have right number of elements")
    return trim_ranks(shape, out) if trim else out


class MultithreadedRNG:
    """Standard-normal fill with one PCG64 stream per CPU thread (SeedSequence.spawn); stream i
    owns the i-th contiguous chunk of the flat output.  The VALUES depend on the thread count,
    exactly as in the reference, so `threads` defaults to cpu_count() there and here."""

    def __init__(self, shape, seed=None, threads: Optional[int] = None):
        self.threads = threads or multiprocessing.cpu_count()
        self.shape = shape
        n = int(np.prod(shape))
        step = int(np.ceil(n / self.threads))
        flat = np.empty(n)
        children = np.random.SeedSequence(seed).spawn(self.threads)
        for i, child in enumerate(children):  # chunks are independent: order does not matter
            np.random.default_rng(child).standard_normal(out=flat[i * step:(i + 1) * step])
        self.values = flat.reshape(shape)


def random_normal(shape, seed=None):
    return MultithreadedRNG(shape, seed).values


def matricize(A: npt.NDArray, mode: Union[int, Sequence[int]], mat_shape: bool = False):
    """Move `mode` axes to the front and flatten the rest (C order)."""
    modes = (mode,) if isinstance(mode, int) else tuple(mode)
    rest = tuple(i for i in range(A.ndim) if i not in modes)
    B = np.transpose(A, modes + rest)
    tail = int(np.prod(B.shape[len(modes):], dtype=np.int64))
    if mat_shape:
        return B.reshape(int(np.prod(B.shape[: len(modes)], dtype=np.int64)), tail)
    return B.reshape(B.shape[: len(modes)] + (tail,))


def right_mul_pinv(A, B, cond=None):
    """A @ pinv(B), evaluated on the GPU (Jacobi-SVD pseudo-inverse + GEMM).  NumPy in/out."""
    from tt_sketch import _backend as be

    rc = -1.0 if cond is None else float(cond)
    out = be.gemm(be.to_device(A, np.float64), be.pinv(be.to_device(B, np.float64), rc))
    return be.to_host(out)


def left_mul_pinv(A, B, cond=None):
    """pinv(A) @ B on the GPU.  NumPy in/out."""
    from tt_sketch import _backend as be

    rc = -1.0 if cond is None else float(cond)
    out = be.gemm(be.pinv(be.to_device(A, np.float64), rc), be.to_device(B, np.float64))
    return be.to_host(out)
