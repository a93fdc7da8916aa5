"""Input tensor containers of the sketching path (the drop-in boundary's data model).

Same class names, constructors and attributes as tt_sketch/tensor.py of the reference for
DenseTensor (:140), SparseTensor (:186), TensorTrain (:294), TensorSum (:612) and CPTensor
(:674).  Only what the sketching hot path and its tests need is provided; the tensor
algebra that is not sketching (round, dot, gather, svdvals, orthogonalize, TuckerTensor) is
out of scope.  Containers additionally cache device copies of their
arrays (`.device()`), keyed by CUDA device and by the identity of the host arrays.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Iterable, List, Optional, Tuple, Union

import numpy as np
import numpy.typing as npt
from numpy.random import SeedSequence

from tt_sketch.utils import ArrayList, TTRank, process_tt_rank, random_normal


class Tensor(ABC):
    shape: Tuple[int, ...]

    @property
    @abstractmethod
    def T(self):
        """Tensor with the order of the modes reversed."""

    @abstractmethod
    def to_numpy(self) -> npt.NDArray[np.float64]:
        ...

    @abstractmethod
    def __mul__(self, other: float):
        ...

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def dense(self) -> "DenseTensor":
        return DenseTensor(self.to_numpy())

    def __rmul__(self, other: float):
        return self.__mul__(other)

    def __truediv__(self, other: float):
        return self.__mul__(1 / other)

    def __neg__(self):
        return self * -1

    def __add__(self, other) -> "TensorSum":
        mine = self.tensors if isinstance(self, TensorSum) else [self]
        theirs = other.tensors if isinstance(other, TensorSum) else [other]
        return TensorSum(list(mine) + list(theirs))

    def __sub__(self, other):
        return self + (-other)

    def norm(self) -> float:
        return float(np.linalg.norm(self.to_numpy()))

    def error(self, other, relative: bool = False, rmse: bool = False, fast: bool = False) -> float:
        """Frobenius distance to `other` through dense reconstruction (small tensors only)."""
        b = other if isinstance(other, np.ndarray) else other.to_numpy()
        err = float(np.linalg.norm(self.to_numpy() - b))
        if relative:
            nb = float(np.linalg.norm(b))
            if nb == 0:
                return float("inf")
            err /= nb
        if rmse:
            err /= float(np.sqrt(np.prod(self.shape)))
        return err

    # ---- device residency ----
    # The upload is cached per (CUDA device, identity of the host arrays).  Replacing an array
    # (`t.entries = ...`, `tt[i] = core`, `tt.cores = [...]`) or switching the current device is
    # detected and triggers a fresh upload, like the reference always reads the current arrays.
    # Editing an array IN PLACE (`t.entries[3] = 0`) cannot be seen without hashing the data:
    # call `invalidate_device()` after such an edit.
    def device(self):
        from tt_sketch import _backend as be

        key = (be.device_index(), self._host_arrays_id())
        if self.__dict__.get("_dev") is None or self.__dict__.get("_dev_key") != key:
            self.__dict__["_dev"] = self._upload()
            self.__dict__["_dev_key"] = key
        return self.__dict__["_dev"]

    def invalidate_device(self) -> None:
        self.__dict__.pop("_dev", None)
        self.__dict__.pop("_dev_key", None)

    def _device_if_current(self):
        """The cached upload if it still matches the host arrays and the current device, else None."""
        from tt_sketch import _backend as be

        if self.__dict__.get("_dev") is not None and \
                self.__dict__.get("_dev_key") == (be.device_index(), self._host_arrays_id()):
            return self.__dict__["_dev"]
        return None

    def _host_arrays_id(self):
        raise NotImplementedError

    def _upload(self):
        raise NotImplementedError


class DenseTensor(Tensor):
    def __init__(self, data: npt.NDArray) -> None:
        self.data = data
        self.shape = tuple(data.shape)

    @property
    def T(self) -> "DenseTensor":
        return DenseTensor(np.transpose(self.data))

    @property
    def size(self) -> int:
        return int(np.prod(self.shape))

    def to_numpy(self):
        return self.data

    def to_sparse(self) -> "SparseTensor":
        idx = np.indices(self.shape).reshape(self.ndim, -1)
        return SparseTensor(self.shape, idx, self.data.reshape(-1))

    def __mul__(self, other: float) -> "DenseTensor":
        return DenseTensor(self.data * other)

    @classmethod
    def random(cls, shape: Tuple[int, ...]) -> "DenseTensor":
        return cls(random_normal(shape))

    def __repr__(self) -> str:
        return f"<Dense tensor of shape {self.shape} at {hex(id(self))}>"

    def _host_arrays_id(self):
        return (id(self.data),)

    def _upload(self):
        from tt_sketch import _backend as be

        return {"data": be.to_device(self.data, np.float64)}


class SparseTensor(Tensor):
    """COO tensor: `indices` is (d, nnz) int64 (a tuple of index vectors is stacked),
    `entries` (nnz,) float64.  Duplicate coordinates are kept; sketching sums them."""

    def __init__(self, shape, indices, entries) -> None:
        self.shape = tuple(int(n) for n in shape)
        self.indices = np.stack(indices) if isinstance(indices, tuple) else indices
        self.entries = entries

    @property
    def T(self) -> "SparseTensor":
        t = SparseTensor(self.shape[::-1], self.indices[::-1], self.entries)
        t.__dict__["_dev_parent"] = self
        if self.__dict__.get("_checked"):
            t.__dict__["_checked"] = True
        return t

    @property
    def nnz(self) -> int:
        return len(self.entries)

    @property
    def size(self) -> int:
        return self.nnz * (self.ndim + 1)

    def split(self, n_summands: int) -> "TensorSum":
        """Contiguous nonzero ranges of equal length (the last takes the remainder)."""
        block = self.nnz // n_summands
        parts: List[Tensor] = []
        for i in range(n_summands):
            hi = (i + 1) * block if i < n_summands - 1 else self.nnz
            sl = slice(i * block, hi)
            parts.append(SparseTensor(self.shape, tuple(row[sl] for row in self.indices), self.entries[sl]))
        return TensorSum(parts)

    def to_numpy(self):
        X = np.zeros(self.shape)
        X[tuple(self.indices)] = self.entries
        return X

    def norm(self) -> float:
        return float(np.linalg.norm(self.entries))

    def __mul__(self, other: float) -> "SparseTensor":
        return SparseTensor(self.shape, self.indices, self.entries * other)

    @classmethod
    def random(cls, shape: Tuple[int, ...], nnz: int, seed: Optional[int] = None) -> "SparseTensor":
        if seed is not None:
            np.random.seed(seed)
        flat = np.random.choice(int(np.prod(shape)), size=nnz, replace=False)
        return cls(shape, np.unravel_index(flat, shape), random_normal(shape=(nnz,), seed=seed))

    def __repr__(self) -> str:
        return f"<Sparse tensor of shape {self.shape} with {self.nnz} non-zero entries at {hex(id(self))}>"

    def check_indices(self):
        """Host-side bounds check (the kernels trust their input)."""
        if self.__dict__.get("_checked"):
            return
        idx = np.asarray(self.indices)
        if idx.shape[0] != self.ndim:
            raise ValueError("indices must have one row per mode")
        if idx.shape[1] and (idx.min() < 0 or np.any(idx.max(axis=1) >= np.array(self.shape))):
            raise ValueError("sparse index out of range for the tensor shape")
        self.__dict__["_checked"] = True

    def _host_arrays_id(self):
        return (id(self.indices), id(self.entries))

    def _upload(self):
        from tt_sketch import _backend as be

        self.__dict__.pop("_checked", None) if self.__dict__.get("_checked_for") != self._host_arrays_id() else None
        self.check_indices()
        self.__dict__["_checked_for"] = self._host_arrays_id()
        parent = self.__dict__.get("_dev_parent")
        pd = parent._device_if_current() if parent is not None else None
        # derive on the device (no second upload) when this is still the transpose of the parent's current arrays
        if pd is not None and self.entries is parent.entries and getattr(self.indices, "base", None) is parent.indices:
            return {"indices": pd["indices"].flip(0), "entries": pd["entries"]}
        return {"indices": be.to_device(self.indices, np.int64), "entries": be.to_device(self.entries, np.float64)}


class TensorTrain(Tensor):
    def __init__(self, cores: ArrayList) -> None:
        self.cores = cores
        self.shape = tuple(C.shape[1] for C in cores)
        self.rank = tuple(C.shape[0] for C in cores[1:])

    @property
    def T(self) -> "TensorTrain":
        t = TensorTrain([np.transpose(C, (2, 1, 0)) for C in reversed(self.cores)])
        t.__dict__["_dev_parent"] = self
        return t

    @property
    def size(self) -> int:
        return sum(C.size for C in self.cores)

    def to_numpy(self):
        X = self.cores[0][0]
        for C in self.cores[1:]:
            X = np.tensordot(X, C, axes=(-1, 0))
        return X[..., 0]

    @classmethod
    def random(cls, shape, rank: TTRank, seed: Optional[int] = None, orthog: bool = False,
               trim: Optional[bool] = None, norm_goal: str = "norm-1") -> "TensorTrain":
        """Gaussian TT cores; core i is drawn as an (r_i*n_i, r_{i+1}) matrix from the stream
        SeedSequence(seed).generate_state(d)[i] and scaled by 1/sqrt(r_i n_i) ('norm-1') or
        1/sqrt(r_i) ('norm-preserve', used for TensorTrainDRM)."""
        d = len(shape)
        if trim is None:
            trim = bool(orthog)
        if orthog and not trim:
            raise ValueError("Trimming must be enabled if orthogonalization is enabled.")
        ranks = (1,) + tuple(process_tt_rank(rank, shape, trim=trim)) + (1,)
        seeds = SeedSequence(seed).generate_state(d)
        cores = []
        for i, n in enumerate(shape):
            r1, r2 = ranks[i], ranks[i + 1]
            M = random_normal(shape=(r1 * n, r2), seed=seeds[i])
            if orthog and i < d - 1:
                M, _ = np.linalg.qr(M, mode="reduced")
            elif norm_goal == "norm-1":
                M /= np.sqrt(r1 * n)
            elif norm_goal == "norm-preserve":
                M /= np.sqrt(r1)
            else:
                raise ValueError(f"Unknown norm goal: {norm_goal}")
            cores.append(M.reshape(r1, n, r2))
        return cls(cores)

    @classmethod
    def zero(cls, shape, rank: TTRank) -> "TensorTrain":
        ranks = (1,) + tuple(process_tt_rank(rank, shape, trim=False)) + (1,)
        return cls([np.zeros((ranks[i], n, ranks[i + 1])) for i, n in enumerate(shape)])

    def __getitem__(self, i: int):
        return self.cores[i]

    def __setitem__(self, i: int, data) -> None:
        self.cores[i] = data

    def __mul__(self, other: float) -> "TensorTrain":
        cores = [c.copy() for c in self.cores]
        cores[-1] = cores[-1] * other
        return TensorTrain(cores)

    def __repr__(self) -> str:
        return f"<Tensor train of shape {self.shape} with rank {self.rank} at {hex(id(self))}>"

    def _host_arrays_id(self):
        return tuple(id(c) for c in self.cores)

    def _upload(self):
        from tt_sketch import _backend as be

        parent = self.__dict__.get("_dev_parent")
        pd = parent._device_if_current() if parent is not None else None
        if pd is not None and len(parent.cores) == len(self.cores) and \
                all(getattr(c, "base", None) is pc for c, pc in zip(self.cores, reversed(parent.cores))):
            return {"cores": [c.permute(2, 1, 0).contiguous() for c in reversed(pd["cores"])]}
        return {"cores": [be.to_device(c, np.float64) for c in self.cores]}


class CPTensor(Tensor):
    """CP format; `cores[i]` has shape (shape[i], rank)."""

    def __init__(self, cores: ArrayList) -> None:
        self.cores = cores
        self.rank = cores[0].shape[1]
        self.shape = tuple(C.shape[0] for C in cores)

    @property
    def T(self) -> "CPTensor":
        t = CPTensor(list(reversed(self.cores)))
        t.__dict__["_dev_parent"] = self
        return t

    def size(self) -> int:
        return sum(C.size for C in self.cores)

    def to_numpy(self):
        letters = "abcdefghijklmnopqrstuvwxy"[: self.ndim]
        spec = ",".join(l + "z" for l in letters) + "->" + letters
        return np.einsum(spec, *self.cores, optimize=True)

    @classmethod
    def random(cls, shape, rank: int, seed: Optional[int] = None) -> "CPTensor":
        seeds = SeedSequence(seed).generate_state(len(shape))
        return cls([random_normal(shape=(n, rank), seed=s) / np.sqrt(n) for n, s in zip(shape, seeds)])

    def __getitem__(self, i: int):
        return self.cores[i]

    def __setitem__(self, i: int, data) -> None:
        self.cores[i] = data

    def __mul__(self, other: float) -> "CPTensor":
        cores = list(self.cores)
        cores[0] = cores[0] * other
        return CPTensor(cores)

    def __repr__(self) -> str:
        return f"<CP tensor of shape {self.shape} and rank {self.rank} at {hex(id(self))}>"

    def _host_arrays_id(self):
        return tuple(id(c) for c in self.cores)

    def _upload(self):
        from tt_sketch import _backend as be

        parent = self.__dict__.get("_dev_parent")
        pd = parent._device_if_current() if parent is not None else None
        if pd is not None and len(parent.cores) == len(self.cores) and \
                all(c is pc for c, pc in zip(self.cores, reversed(parent.cores))):
            return {"cores": list(reversed(pd["cores"]))}
        return {"cores": [be.to_device(c, np.float64) for c in self.cores]}


class TensorSum(Tensor):
    """Lazy sum of tensors of one shape; sketched summand by summand (the sketch is linear)."""

    def __init__(self, tensors: List[Tensor], shape=None) -> None:
        self.tensors = tensors
        self.shape = tuple(shape) if shape is not None else tuple(tensors[0].shape)

    @property
    def T(self) -> "TensorSum":
        return TensorSum([X.T for X in self.tensors])

    @property
    def size(self) -> int:
        return sum(X.size if not callable(X.size) else X.size() for X in self.tensors)

    @property
    def num_summands(self) -> int:
        return len(self.tensors)

    def to_numpy(self):
        out = np.zeros(self.shape)
        for X in self.tensors:
            out += X.to_numpy()
        return out

    def __iadd__(self, other) -> "TensorSum":
        self.tensors.extend(other.tensors if isinstance(other, TensorSum) else [other])
        return self

    def __mul__(self, other: Union[float, Iterable[float]]) -> "TensorSum":
        try:
            coeffs = list(other)  # type: ignore[arg-type]
        except TypeError:
            return TensorSum([X * other for X in self.tensors])
        if len(coeffs) != len(self.tensors):
            raise ValueError("one coefficient per summand expected")
        return TensorSum([X * c for X, c in zip(self.tensors, coeffs)])

    def __repr__(self) -> str:
        return f"<Sum of {self.num_summands} tensors of shape {self.shape} at {hex(id(self))}>"
