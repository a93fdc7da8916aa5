"""Input tensor containers of the sketching path (the drop-in boundary's data model).

Same class names, constructors and attributes as tt_sketch/tensor.py of the reference for
DenseTensor (:140), SparseTensor (:186), TensorTrain (:294), TensorSum (:612) and CPTensor
(:674) and TuckerTensor (:746).  What the sketching hot path, its tests and the validation step after it
need is provided (dot / norm / error / gather; TensorTrain.gather runs on the device); round, svdvals and
orthogonalize are out of scope.  Containers additionally cache device copies of their
arrays (`.device()`), keyed by CUDA device and by the identity of the host arrays.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Iterable, List, Optional, Tuple, Union

import numpy as np
import numpy.typing as npt
from numpy.random import SeedSequence

from tt_sketch.utils import ArrayList, TTRank, process_tt_rank, random_normal


def _is_transposed_view(c, pc) -> bool:
    """`c` is `pc` with its axes reversed, on the same memory (what `.T` of a container hands out)."""
    return (isinstance(c, np.ndarray) and isinstance(pc, np.ndarray) and c.dtype == pc.dtype
            and c.shape == pc.shape[::-1] and c.strides == pc.strides[::-1]
            and c.__array_interface__["data"][0] == pc.__array_interface__["data"][0])


def _copy_into(dev_tensors, host_arrays) -> bool:
    """Host arrays -> existing device tensors of the same shapes (one H2D copy each); False if anything differs."""
    import torch

    if len(dev_tensors) != len(host_arrays):
        return False
    for t, a in zip(dev_tensors, host_arrays):
        if not isinstance(a, np.ndarray) or tuple(t.shape) != tuple(a.shape) or not t.is_contiguous():
            return False
    for t, a in zip(dev_tensors, host_arrays):
        t.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)))
    return True


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

    def dot(self, other, reverse: bool = False) -> float:
        """Inner product (reference tensor.py:121-129): a TensorSum distributes, the other operand is asked first
        (a SparseTensor only needs `other.gather`), dense reconstruction is the last resort."""
        if isinstance(other, TensorSum):
            return other.dot(self)
        if not reverse:
            return other.dot(self, reverse=True)
        return float(np.dot(self.to_numpy().reshape(-1), other.to_numpy().reshape(-1)))

    def norm(self) -> float:
        return float(np.sqrt(np.abs(self.dot(self))))

    def error(self, other, relative: bool = False, rmse: bool = False, fast: bool = False) -> float:
        """Frobenius distance to `other` (reference tensor.py:52-87): through dense reconstruction, or with
        `fast=True` through the inner-product formula (norms and one dot product; for a SparseTensor against a
        TensorTrain that is one device gather at the nonzeros -- how errors are sampled on tensors too large to
        densify; inaccurate below relative errors of about 1e-8, like the reference says)."""
        if isinstance(other, np.ndarray):
            other = DenseTensor(other)
        other_norm = other.norm()
        if fast:
            self_norm = self.norm()
            dot = self.dot(other)
            norm_sum = self_norm ** 2 + other_norm ** 2
            err = float(np.sqrt(norm_sum) * np.sqrt(np.abs(1 - 2 * dot / norm_sum)))
        else:
            err = float(np.linalg.norm(self.to_numpy() - other.to_numpy()))
        if relative:
            if other_norm == 0:
                return float("inf")
            err /= other_norm
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
        dev = self.__dict__.get("_dev")
        if dev is None or self.__dict__.get("_dev_key") != key:
            slot = self.__dict__.get("_dev_slot")
            if dev is None or slot != key[0] or not self._refresh_in_place(dev):
                self.__dict__["_dev"] = self._upload()
            self.__dict__["_dev_key"] = key
            self.__dict__["_dev_slot"] = key[0]
        return self.__dict__["_dev"]

    def invalidate_device(self) -> None:
        """The host arrays were edited: the next `.device()` uploads them again -- into the SAME device arrays when the
        shapes still fit (so CUDA graphs captured on them stay valid), into new ones otherwise."""
        self.__dict__.pop("_dev_key", None)

    def _refresh_in_place(self, dev) -> bool:
        """Copy the current host arrays into the existing device arrays `dev`; False if that is not possible."""
        return False

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
        t = DenseTensor(np.transpose(self.data))
        t.__dict__["_dev_parent"] = self
        return t

    @property
    def size(self) -> int:
        return int(np.prod(self.shape))

    def to_numpy(self):
        return self.data

    def norm(self) -> float:
        return float(np.linalg.norm(self.data))

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

    def _refresh_in_place(self, dev) -> bool:
        return "_dev_parent" not in self.__dict__ and _copy_into([dev["data"]], [self.data])

    def _upload(self):
        from tt_sketch import _backend as be

        parent = self.__dict__.get("_dev_parent")
        pd = parent._device_if_current() if parent is not None else None
        if pd is not None and _is_transposed_view(self.data, parent.data):  # transpose of the parent's upload, on the device
            x = pd["data"]
            return {"data": x.permute(*reversed(range(x.dim()))).contiguous()}
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

    def dot(self, other, reverse: bool = False) -> float:
        """sum_p entries[p] * other[indices[:, p]] when `other` can gather (reference tensor.py:250-255)."""
        if hasattr(other, "gather"):
            return float(np.dot(other.gather(self.indices), self.entries))
        return super().dot(other, reverse=reverse)

    def gather(self, indices) -> npt.NDArray[np.float64]:
        """Entries at the given multi-indices, 0 where nothing is stored (reference tensor.py:275-291; duplicates of
        a stored index: the last one wins, like the reference's dict)."""
        flat = np.ravel_multi_index(tuple(np.asarray(r) for r in indices), self.shape)
        own = np.ravel_multi_index(tuple(np.asarray(r) for r in self.indices), self.shape)
        order = np.argsort(own, kind="stable")
        own_sorted = own[order]
        pos = np.searchsorted(own_sorted, flat, side="right") - 1
        hit = (pos >= 0) & (own_sorted[np.clip(pos, 0, None)] == flat)
        out = np.zeros(len(flat))
        out[hit] = np.asarray(self.entries)[order[pos[hit]]]
        return out

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

    def gather(self, idx) -> npt.NDArray[np.float64]:
        """One entry of the represented tensor per column of `idx` (d x N): the chained products of the core slices
        at the indices, left to right (reference tensor.py:414-440, which loops over every slice of every mode).
        Runs on the GPU with the per-nonzero chain kernel of the TT-DRM (`ttsk_ttdrm_sparse_step`): this is how errors
        are sampled on sparse tensors too large to densify."""
        from tt_sketch import _backend as be

        idx = np.ascontiguousarray(np.stack([np.asarray(r) for r in idx]) if not isinstance(idx, np.ndarray) else idx,
                                   dtype=np.int64)
        if idx.shape[0] != self.ndim:
            raise ValueError(f"gather needs {self.ndim} index rows, got {idx.shape[0]}")
        if idx.shape[1] == 0:
            return np.zeros(0)
        d_idx = be.to_device(idx, np.int64)
        cores = self.device()["cores"]
        v = None
        for mu in range(self.ndim):
            v = be.ttdrm_sparse_step(d_idx[mu], v, cores[mu])
        return be.to_host(v).reshape(-1)

    def dot(self, other, reverse: bool = False) -> float:
        """TT . TT in a left-to-right sweep on the device (reference tensor.py:542-560: two GEMMs per mode);
        anything else through the base class."""
        if isinstance(other, TensorTrain):
            from tt_sketch import _backend as be

            mine, theirs = self.device()["cores"], other.device()["cores"]
            res = None
            for c1, c2 in zip(mine, theirs):
                r1, n, a = c1.shape
                r2, _, b = c2.shape
                if res is None:
                    res = be.gemm(c1.reshape(n, a).T, c2.reshape(n, b))                  # (a, b)
                else:
                    w = be.gemm(res.T, c1.reshape(r1, n * a))                            # (r2, n * a)
                    res = be.gemm(w.reshape(r2 * n, a).T, c2.reshape(r2 * n, b))         # (a, b)
            return float(be.to_host(res).sum())
        return super().dot(other, reverse=reverse)

    # ---- the step after the sketch: orthogonalisation, rounding, norms of TTs (reference tensor.py:442-609) on the
    # device: QR by ttsk_qr_q (LAPACK's Householder signs, so Q equals np.linalg.qr's), R = Q^T A and the core
    # updates by ttsk_gemm, SVDs by ttsk_svd (one-sided Jacobi).
    def orthogonalize(self) -> "TensorTrain":
        """QR sweep: every core but the last gets orthonormal columns as an (r n, r') matrix (reference :562-575)."""
        from tt_sketch import _backend as be

        cores = self.device()["cores"]
        new, R = [], None
        for mu, C in enumerate(cores):
            r0, n, r1 = C.shape
            if R is not None:
                C = be.gemm(R, C.reshape(r0, n * r1))
                r0 = C.shape[0]
            A = C.reshape(r0 * n, r1)
            if mu == len(cores) - 1:
                new.append(A.reshape(r0, n, r1))
                break
            m = r0 * n
            k = min(m, r1)
            Q = A[:, :k].clone() if k < r1 else A.clone()   # a wide unfolding: Q of its leading square block
            be.qr_q_inplace(Q)
            R = be.gemm(Q.T, A)
            new.append(Q.reshape(r0, n, k))
        return TensorTrain([be.to_host(c) for c in new])

    def norm(self) -> float:
        return float(np.linalg.norm(self.orthogonalize().cores[-1]))

    def add(self, other: "TensorTrain") -> "TensorTrain":
        """Direct sum of the cores (reference :519-540); `+` stays the lazy TensorSum."""
        cores = [np.concatenate((self.cores[0], other.cores[0]), axis=2)]
        for a, b in zip(self.cores[1:-1], other.cores[1:-1]):
            blk = np.zeros((a.shape[0] + b.shape[0], a.shape[1], a.shape[2] + b.shape[2]))
            blk[: a.shape[0], :, : a.shape[2]] = a
            blk[a.shape[0]:, :, a.shape[2]:] = b
            cores.append(blk)
        cores.append(np.concatenate((self.cores[-1], other.cores[-1]), axis=0))
        return TensorTrain(cores)

    def _rl_svd_sweep(self, eps, max_rank, orthogonalized: bool, keep_cores: bool):
        from tt_sketch import _backend as be

        tt = self if orthogonalized else self.orthogonalize()
        cores = tt.device()["cores"]
        d = len(cores)
        new, svals, US = [], [], None
        for mu in range(d - 1, -1, -1):
            C = cores[mu]
            r0, n, r1 = C.shape
            if US is not None:
                C = be.gemm(C.reshape(r0 * n, r1), US).reshape(r0, n, US.shape[1])
                r1 = C.shape[2]
            if mu > 0:
                U, S, Vt = be.svd(C.reshape(r0, n * r1), u_times_s=True)
                s_host = be.to_host(S)
                svals.append(s_host)
                r = len(s_host) if eps is None else max(1, min(int(np.sum(s_host > s_host[0] * eps)), int(max_rank[mu - 1])))
                US = U[:, :r]
                if keep_cores:
                    new.append(be.to_host(Vt[:r]).reshape(r, n, r1))
            else:
                if keep_cores:
                    new.append(be.to_host(C))
                else:  # svdvals of the first unfolding (r0 n, r1)
                    svals.append(be.to_host(be.svd(C.reshape(r0 * n, r1))[1]))
        return new[::-1], svals[::-1]

    def round(self, eps: Optional[float] = None, max_rank=None, orthogonalized: bool = False) -> "TensorTrain":
        """TT-SVD rounding: left-orthogonalise, then truncate in a right-to-left SVD sweep (reference :446-484;
        singular values above eps * s_max are kept, at most max_rank)."""
        if eps is None:
            eps = 0
        if max_rank is None:
            max_rank = self.rank
        max_rank = process_tt_rank(max_rank, self.shape, trim=True)
        cores, _ = self._rl_svd_sweep(eps, max_rank, orthogonalized, keep_cores=True)
        return TensorTrain(cores)

    def svdvals(self):
        """Singular values of every unfolding (reference :486-506)."""
        return self._rl_svd_sweep(None, None, False, keep_cores=False)[1]

    def error(self, other, relative: bool = False, rmse: bool = False, fast: bool = False) -> float:
        """Against another TT (or anything with `to_tt`): ||self - other|| from the orthogonalised difference, exact
        without densifying (reference :577-609); otherwise the generic rule."""
        if hasattr(other, "to_tt") and not isinstance(other, TensorTrain):
            other = other.to_tt()
        if isinstance(other, TensorTrain):
            err = self.add(-other).norm()
            if relative:
                on = other.norm()
                if on == 0:
                    return float("inf")
                err /= on
            if rmse:
                err /= float(np.sqrt(np.prod(self.shape)))
            return float(err)
        return super().error(other, relative=relative, rmse=rmse, fast=fast)

    def __mul__(self, other: float) -> "TensorTrain":
        cores = [c.copy() for c in self.cores]
        cores[-1] = cores[-1] * other
        return TensorTrain(cores)

    def __repr__(self) -> str:
        return f"<Tensor train of shape {self.shape} with rank {self.rank} at {hex(id(self))}>"

    def _host_arrays_id(self):
        return tuple(id(c) for c in self.cores)

    def _refresh_in_place(self, dev) -> bool:
        return "_dev_parent" not in self.__dict__ and _copy_into(dev["cores"], list(self.cores))

    def _upload(self):
        from tt_sketch import _backend as be

        parent = self.__dict__.get("_dev_parent")
        pd = parent._device_if_current() if parent is not None else None
        if pd is not None and len(parent.cores) == len(self.cores) and \
                all(_is_transposed_view(c, pc) for c, pc in zip(self.cores, reversed(parent.cores))):
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

    def gather(self, idx) -> npt.NDArray[np.float64]:
        """Values at the given indices (reference tensor.py:726-732)."""
        res = 1
        for C, i in zip(self.cores, idx):
            res = res * np.asarray(C)[np.asarray(i)]
        return np.sum(res, axis=1)

    def __mul__(self, other: float) -> "CPTensor":
        cores = list(self.cores)
        cores[0] = cores[0] * other
        return CPTensor(cores)

    def __repr__(self) -> str:
        return f"<CP tensor of shape {self.shape} and rank {self.rank} at {hex(id(self))}>"

    def _host_arrays_id(self):
        return tuple(id(c) for c in self.cores)

    def _refresh_in_place(self, dev) -> bool:
        return "_dev_parent" not in self.__dict__ and _copy_into(dev["cores"], list(self.cores))

    def _upload(self):
        from tt_sketch import _backend as be

        parent = self.__dict__.get("_dev_parent")
        pd = parent._device_if_current() if parent is not None else None
        if pd is not None and len(parent.cores) == len(self.cores) and \
                all(c is pc for c, pc in zip(self.cores, reversed(parent.cores))):
            return {"cores": list(reversed(pd["cores"]))}
        return {"cores": [be.to_device(c, np.float64) for c in self.cores]}


class TuckerTensor(Tensor):
    """Tucker format (reference tensor.py:746-816): a core of shape `rank` = (s_1..s_d) and d factor matrices
    `factors[i]` of shape (s_i, n_i)."""

    def __init__(self, factors: ArrayList, core: npt.NDArray) -> None:
        self.core = core
        self.factors = factors
        self.shape = tuple(U.shape[1] for U in factors)
        self.rank = tuple(U.shape[0] for U in factors)

    @property
    def T(self) -> "TuckerTensor":
        t = TuckerTensor(self.factors[::-1], np.transpose(self.core))
        t.__dict__["_dev_parent"] = self
        return t

    @property
    def size(self) -> int:
        return int(self.core.size + sum(U.size for U in self.factors))

    def to_numpy(self):
        out = self.core
        for i, U in enumerate(self.factors):  # mode-i product with U_i^T, one mode at a time
            out = np.moveaxis(np.tensordot(out, U, axes=([i], [0])), -1, i)
        return out

    def __mul__(self, other: float) -> "TuckerTensor":
        return TuckerTensor(self.factors, self.core * other)

    def __repr__(self) -> str:
        return f"<Tucker tensor of shape {self.shape} and rank {self.rank} at {hex(id(self))}>"

    @classmethod
    def random(cls, shape, rank, seed: Optional[int] = None) -> "TuckerTensor":
        """Gaussian core, orthonormal-row factors; same draws as the reference (tensor.py:792-816: the core seed is
        the first word of the SeedSequence state, the factor seeds its first d words)."""
        d = len(shape)
        try:
            ranks = tuple(rank)
        except TypeError:
            ranks = (rank,) * d
        ranks = tuple(min(int(r), int(n)) for r, n in zip(ranks, shape))
        seq = SeedSequence(seed)
        core = random_normal(shape=ranks, seed=seq.generate_state(1)[0])
        factors = [np.linalg.qr(random_normal(shape=(r, n), seed=s).T)[0].T
                   for r, n, s in zip(ranks, shape, seq.generate_state(d))]
        return cls(factors, core)

    def _host_arrays_id(self):
        return (id(self.core),) + tuple(id(U) for U in self.factors)

    def _refresh_in_place(self, dev) -> bool:
        return "_dev_parent" not in self.__dict__ and \
            _copy_into([dev["core"]] + list(dev["factors"]), [self.core] + list(self.factors))

    def _upload(self):
        from tt_sketch import _backend as be

        parent = self.__dict__.get("_dev_parent")
        pd = parent._device_if_current() if parent is not None else None
        if pd is not None and _is_transposed_view(self.core, parent.core) and len(self.factors) == len(parent.factors) \
                and all(a is b for a, b in zip(self.factors, reversed(parent.factors))):
            c = pd["core"]
            return {"core": c.permute(*reversed(range(c.dim()))).contiguous(), "factors": list(reversed(pd["factors"]))}
        return {"core": be.to_device(self.core, np.float64),
                "factors": [be.to_device(U, np.float64) for U in self.factors]}


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

    def dot(self, other, reverse: bool = False) -> float:
        return float(sum(X.dot(other, reverse) for X in self.tensors))

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
