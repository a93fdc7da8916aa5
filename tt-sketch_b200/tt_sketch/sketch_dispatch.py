"""Driver of the sketching path: picks the DRM contraction and the Omega/Psi kernels for a
tensor type and runs the streaming / orthogonal / HMT loops on the GPU.

Mirror of tt_sketch/sketch_dispatch.py (reference): registries DRM_SKETCH_METHOD_DISPATCH /
OMEGA_METHODS / PSI_METHODS (:59-82), TensorSum handling (:85-147), get_sketch_method (:150-157),
orth_step (:160-174), OrthogTTDRM (:177-193), SketchMethod (:196-199), general_sketch (:202-275).

Execution model (B200): the whole sketch lives in ONE packed device buffer
[Psi_0|...|Psi_{d-1}|Omega_0|...|Omega_{d-2}]; every summand of a TensorSum accumulates into it
(the sketch is linear), and it is copied to the host once at the end (or all-reduced across
GPUs first, see tt_sketch/distributed.py).  Sparse summands of a streaming sketch go through the
fused `ttsk_sparse_sketch` entry point (DRM entries generated on the fly, nothing
materialised); everything else runs the per-bond device operators registered below.
"""
from __future__ import annotations

import enum
from ctypes import byref
from functools import partial
from typing import Callable, List, Optional

import numpy as np

from tt_sketch import _backend as be
from tt_sketch.drm import SparseGaussianDRM, TensorTrainDRM
from tt_sketch.drm_base import DRM
from tt_sketch.sketch_container import SketchContainer
from tt_sketch.sketching_methods.abstract_methods import (CansketchCP, CansketchDense, CansketchSparse,
                                                          CanSketchTucker, CansketchTT)
from tt_sketch.sketching_methods.cp_sketch import (omega_cp_device, psi_cp_device, sketch_omega_cp,
                                                   sketch_psi_cp)
from tt_sketch.sketching_methods.dense_sketch import (omega_dense_device, psi_dense_device,
                                                      sketch_omega_dense, sketch_psi_dense)
from tt_sketch.sketching_methods.sparse_sketch import (omega_sparse_device, psi_sparse_device,
                                                       sketch_omega_sparse, sketch_psi_sparse)
from tt_sketch.sketching_methods.tensor_train_sketch import (omega_tt_device, psi_tt_device,
                                                             sketch_omega_tt, sketch_psi_tt)
from tt_sketch.sketching_methods.tucker_sketch import (omega_tucker_device, psi_tucker_device,
                                                       sketch_omega_tucker, sketch_psi_tucker)
from tt_sketch.tensor import CPTensor, DenseTensor, SparseTensor, Tensor, TensorSum, TensorTrain, TuckerTensor
from tt_sketch.utils import right_mul_pinv  # noqa: F401  (re-exported like the reference)

ABSTRACT_TENSOR_SKETCH_DISPATCH = {
    SparseTensor: CansketchSparse,
    TensorTrain: CansketchTT,
    DenseTensor: CansketchDense,
    CPTensor: CansketchCP,
    TuckerTensor: CanSketchTucker,
}

# tensor class -> name of the DRM method that contracts the DRM with it (host arrays)
DRM_SKETCH_METHOD_DISPATCH = {
    SparseTensor: "sketch_sparse",
    TensorTrain: "sketch_tt",
    DenseTensor: "sketch_dense",
    CPTensor: "sketch_cp",
    TuckerTensor: "sketch_tucker",
}

# the plug-in point: NumPy-in / NumPy-out operators, mutable like the reference's dicts
OMEGA_METHODS = {
    SparseTensor: sketch_omega_sparse,
    TensorTrain: sketch_omega_tt,
    DenseTensor: sketch_omega_dense,
    CPTensor: sketch_omega_cp,
    TuckerTensor: sketch_omega_tucker,
}
PSI_METHODS = {
    SparseTensor: sketch_psi_sparse,
    TensorTrain: sketch_psi_tt,
    DenseTensor: sketch_psi_dense,
    CPTensor: sketch_psi_cp,
    TuckerTensor: sketch_psi_tucker,
}

# device-resident twins used by general_sketch (accumulate into `out`)
OMEGA_DEVICE = {
    SparseTensor: omega_sparse_device,
    TensorTrain: omega_tt_device,
    DenseTensor: omega_dense_device,
    CPTensor: omega_cp_device,
    TuckerTensor: omega_tucker_device,
}
PSI_DEVICE = {
    SparseTensor: psi_sparse_device,
    TensorTrain: psi_tt_device,
    DenseTensor: psi_dense_device,
    CPTensor: psi_cp_device,
    TuckerTensor: psi_tucker_device,
}


# ------------------------------------------------------------------ TensorSum (host-level API)
def sketch_omega_sum(left_sketch_array, right_sketch_array, *, tensor: TensorSum, omega_shape, **kwargs):
    total = np.zeros(omega_shape)
    for X, L, R in zip(tensor.tensors, left_sketch_array, right_sketch_array):
        total += OMEGA_METHODS[type(X)](L, R, tensor=X, omega_shape=omega_shape, **kwargs)
    return total


def sketch_psi_sum(left_sketch_array, right_sketch_array, *, tensor: TensorSum, psi_shape, **kwargs):
    n = tensor.num_summands
    Ls = left_sketch_array if left_sketch_array is not None else (None,) * n
    Rs = right_sketch_array if right_sketch_array is not None else (None,) * n
    total = np.zeros(psi_shape)
    for X, L, R in zip(tensor.tensors, Ls, Rs):
        total += PSI_METHODS[type(X)](L, R, tensor=X, psi_shape=psi_shape, **kwargs)
    return total


OMEGA_METHODS[TensorSum] = sketch_omega_sum
PSI_METHODS[TensorSum] = sketch_psi_sum


# The public dicts are the reference's plug-in point: replacing (or adding) an entry must change what the sketching
# loops run.  `general_sketch` takes the device twin of a tensor type only while its public entries are still the
# stock ones; otherwise the registered NumPy-in / NumPy-out operator is called on host copies of the DRM
# contractions and its result is added to the device buffer (and the fused single-call paths are skipped).
_STOCK_OMEGA, _STOCK_PSI, _STOCK_DRM_METHOD = dict(OMEGA_METHODS), dict(PSI_METHODS), dict(DRM_SKETCH_METHOD_DISPATCH)


def _is_stock(t) -> bool:
    return (t in _STOCK_OMEGA and OMEGA_METHODS.get(t) is _STOCK_OMEGA[t] and PSI_METHODS.get(t) is _STOCK_PSI[t]
            and DRM_SKETCH_METHOD_DISPATCH.get(t) == _STOCK_DRM_METHOD[t] and t in OMEGA_DEVICE)


def _host(x):
    return None if x is None else np.ascontiguousarray(be.to_host(x))


def device_operators(t):
    """(omega, psi) with the device-twin signature `(L, R, *, tensor, mu, out)` for tensor class `t`."""
    if _is_stock(t):
        return OMEGA_DEVICE[t], PSI_DEVICE[t]
    if t not in OMEGA_METHODS or t not in PSI_METHODS:
        raise ValueError(f"no Omega / Psi method registered for {t}")

    def omega(L, R, *, tensor, mu, out, **kw):
        res = OMEGA_METHODS[t](_host(L), _host(R), tensor=tensor, mu=mu, omega_shape=tuple(out.shape))
        out += be.to_device(np.asarray(res, dtype=np.float64).reshape(tuple(out.shape)))
        return out

    def psi(L, R, *, tensor, mu, out, **kw):
        res = PSI_METHODS[t](_host(L), _host(R), tensor=tensor, mu=mu, psi_shape=tuple(out.shape))
        out += be.to_device(np.asarray(res, dtype=np.float64).reshape(tuple(out.shape)))
        return out

    return omega, psi


def sum_sketch(tensor: TensorSum, *, drm: DRM):
    """Per-summand DRM contractions advanced in lock step: yields one tuple per bond."""
    gens = [get_sketch_method(X, drm)(X) for X in tensor.tensors]
    for _ in range(len(tensor.shape) - 1):
        yield tuple(next(g) for g in gens)


def get_sketch_method(tensor: Tensor, drm: DRM, device: bool = False) -> Callable:
    name = DRM_SKETCH_METHOD_DISPATCH.get(type(tensor))
    if name is not None:
        if device and not hasattr(drm, name + "_device"):  # a plugged-in host-only contraction: upload what it yields
            host_method = getattr(drm, name)
            return lambda X: (be.to_device(np.ascontiguousarray(m), np.float64) for m in host_method(X))
        return getattr(drm, name + "_device" if device else name)
    if isinstance(tensor, TensorSum):
        return partial(sum_sketch, drm=drm)
    raise ValueError(f"DRM of type {type(drm)} can't sketch {type(tensor)}")


# ------------------------------------------------------------------ orthogonalisation step
def orth_step_device(Psi, Omega):
    """Psi (r1, n, rR) device, Omega (rL, rR) device or None -> Q factor reshaped (r1, n, rL|rR).
    Psi_mat @ pinv(Omega) by Jacobi-SVD pseudo-inverse + GEMM, then Householder QR."""
    r1, n, r2 = Psi.shape
    mat = Psi.reshape(r1 * n, r2)
    if Omega is not None:
        mat = be.gemm(mat, be.pinv(Omega))
    else:
        mat = mat.clone()
    if mat.shape[0] < mat.shape[1]:
        raise ValueError(f"cannot orthogonalise a {tuple(mat.shape)} unfolding: rank exceeds r*n (trim the rank)")
    be.qr_q_inplace(mat)
    return mat.reshape(r1, n, mat.shape[1])


def orth_step(Psi: np.ndarray, Omega: Optional[np.ndarray]) -> np.ndarray:
    """NumPy-level twin of the reference's orth_step (sketch_dispatch.py:160-174)."""
    q = orth_step_device(be.to_device(Psi, np.float64), be.to_device(Omega, np.float64) if Omega is not None else None)
    return be.to_host(q)


class OrthogTTDRM:
    """Left DRM of the orthogonal / HMT loops: a TensorTrainDRM whose cores are the
    orthogonalised Psi cores produced so far (device tensors), contracted lazily."""

    def __init__(self, rank, tensor):
        self.rank = rank
        self.drm = TensorTrainDRM(rank, tensor.shape, transpose=False, cores=[])
        self.tensor = tensor
        self.generators = None

    def add_core(self, core):
        self.drm.cores.append(core)
        if self.generators is None:
            self.generators = [get_sketch_method(X, self.drm, device=True)(X) for X in _summands(self.tensor)]

    def __next__(self):
        return [next(g) for g in self.generators]


class SketchMethod(enum.Enum):
    streaming = "streaming"
    orthogonal = "orthogonal"
    hmt = "hmt"


def _summands(tensor: Tensor) -> List[Tensor]:
    if isinstance(tensor, TensorSum):
        out: List[Tensor] = []
        for X in tensor.tensors:
            out.extend(_summands(X))
        return out
    return [tensor]


def _check_supported(X: Tensor, drm: DRM):
    if type(X) not in DRM_SKETCH_METHOD_DISPATCH:
        raise ValueError(f"DRM of type {type(drm)} can't sketch {type(X)}")
    name = DRM_SKETCH_METHOD_DISPATCH[type(X)]
    if not hasattr(drm, name + "_device"):
        getattr(drm, name)  # AttributeError if the DRM lacks the capability (like the reference)


def drm_descriptor(drm: DRM):
    """ctypes `ttsk_drm` for a SparseGaussianDRM / TensorTrainDRM (plus keep-alive refs)."""
    desc = be.TtskDrm()
    desc.kind = drm.kind
    desc.right = 1 if drm.transpose else 0
    desc.seed = int(drm.seed)
    keep = []
    for mu, (lo, hi) in enumerate(zip(drm.bond_rank_min, drm.bond_rank_max)):
        desc.rank_min[mu], desc.rank_max[mu] = int(lo), int(hi)
    if drm.kind == be.DRM_TT:
        for k in range(len(drm.cores)):
            c = drm.device_core(k)
            keep.append(c)
            desc.d_cores[k] = c.data_ptr()
            desc.core_r0[k], desc.core_r1[k] = int(c.shape[0]), int(c.shape[2])
    return desc, keep


def _fused_sparse(X: SparseTensor, left_drm: DRM, right_drm: DRM, packed, accumulate: bool):
    dev = X.device()
    idx, val = dev["indices"], dev["entries"]
    ld, lkeep = drm_descriptor(left_drm)
    rd, rkeep = drm_descriptor(right_drm)
    be.check(be.lib().ttsk_sparse_sketch(be.ctx(), X.ndim, be.as_i64(X.shape), X.nnz, be.ptr(idx), idx.stride(0),
                                         be.ptr(val), byref(ld), byref(rd), be.ptr(packed),
                                         1 if accumulate else 0, be.stream()))
    del lkeep, rkeep


def _fused_tt(X: TensorTrain, left_drm: DRM, right_drm: DRM, packed):
    """One C call for a TensorTrain summand under TT DRMs (ttsk_tt_sketch): the same GEMMs as
    TensorTrainDRM.sketch_tt_device + omega_tt_device / psi_tt_device, without a host round trip each."""
    from ctypes import c_void_p

    cores = [c if c.is_contiguous() else c.contiguous() for c in X.device()["cores"]]
    ld, lkeep = drm_descriptor(left_drm)
    rd, rkeep = drm_descriptor(right_drm)
    ptrs = (c_void_p * len(cores))(*[c.data_ptr() for c in cores])
    be.check(be.lib().ttsk_tt_sketch(be.ctx(), X.ndim, be.as_i64(X.shape), be.as_i32((1,) + tuple(X.rank) + (1,)), ptrs,
                                     byref(ld), byref(rd), be.ptr(packed), be.stream()))
    del lkeep, rkeep, cores


def _fused_dense(X: DenseTensor, left_drm: DRM, right_drm: DRM, packed):
    """One C call for a DenseTensor summand under TT DRMs (ttsk_dense_sketch): X is read once (TMA-staged first
    pass), the left DRM is swept, the right unfoldings are used as flat arrays like the reference's dense path."""
    x = X.device()["data"]
    x = x if x.is_contiguous() else x.contiguous()
    ld, lkeep = drm_descriptor(left_drm)
    rd, rkeep = drm_descriptor(right_drm)
    be.check(be.lib().ttsk_dense_sketch(be.ctx(), X.ndim, be.as_i64(X.shape), be.ptr(x), byref(ld), byref(rd),
                                        be.ptr(packed), be.stream()))
    del lkeep, rkeep, x


def _fusable_dense(X: Tensor, left_drm: DRM, right_drm: DRM) -> bool:
    """DenseTensor under two unsliced TensorTrainDRMs (the reference's dense sketch ignores rank slices)."""
    if not (isinstance(X, DenseTensor) and type(left_drm) is TensorTrainDRM and type(right_drm) is TensorTrainDRM
            and tuple(left_drm.shape) == tuple(X.shape) == tuple(right_drm.shape) and X.ndim >= 2):
        return False
    for drm in (left_drm, right_drm):
        if any(int(lo) != 0 for lo in drm.bond_rank_min) or tuple(drm.bond_rank_max) != tuple(drm.bond_true_rank):
            return False
    return True


def _fusable_tt(X: Tensor, left_drm: DRM, right_drm: DRM) -> bool:
    return (isinstance(X, TensorTrain) and type(left_drm) is TensorTrainDRM and type(right_drm) is TensorTrainDRM
            and tuple(left_drm.shape) == tuple(X.shape) == tuple(right_drm.shape))


_warned_wide_rank = False


def _fusable(X: Tensor, left_drm: DRM, right_drm: DRM) -> bool:
    global _warned_wide_rank
    if not (isinstance(X, SparseTensor) and type(left_drm) in (SparseGaussianDRM, TensorTrainDRM)
            and type(right_drm) in (SparseGaussianDRM, TensorTrainDRM) and X.nnz > 0):
        return False
    if max(max(left_drm.rank), max(right_drm.rank)) > 64:
        if not _warned_wide_rank:  # say so once: the operator-level path materialises (rank x nnz) DRM matrices
            import warnings

            warnings.warn("tt_sketch: a DRM rank above 64 takes the operator-level sparse path (DRM contractions "
                          "materialised as (rank x nnz) device arrays) instead of the fused sparse kernels",
                          RuntimeWarning, stacklevel=3)
            _warned_wide_rank = True
        return False
    return True


def streaming_sketch_device(tensor: Tensor, left_drm: DRM, right_drm: DRM, packed=None):
    """Accumulate the streaming sketch of `tensor` into the packed device buffer (allocated and
    zeroed if None) and return it with its layout.  Used by general_sketch and by the multi-GPU
    driver, which all-reduces the buffer before it is unpacked."""
    shape = tuple(tensor.shape)
    d = len(shape)
    rL, rR = tuple(left_drm.bond_rank), tuple(right_drm.bond_rank)
    items, total = SketchContainer.layout(shape, rL, rR)
    if packed is None:
        packed = be.zeros(total)
    views = [packed[o:o + int(np.prod(s))].reshape(s) for o, s in items]
    for X in _summands(tensor):
        if tuple(X.shape) != shape:
            raise ValueError(f"Shape {left_drm.shape} of DRM doesn't match tensor's shape {X.shape}")
        stock = _is_stock(type(X))
        if stock and _fusable(X, left_drm, right_drm):
            if tuple(left_drm.shape) != shape or tuple(right_drm.shape) != shape:
                raise ValueError(f"Shape {left_drm.shape} of DRM doesn't match tensor's shape {shape}")
            _fused_sparse(X, left_drm, right_drm, packed, accumulate=True)
            continue
        if stock and _fusable_tt(X, left_drm, right_drm):
            _fused_tt(X, left_drm, right_drm, packed)
            continue
        if stock and _fusable_dense(X, left_drm, right_drm):
            _fused_dense(X, left_drm, right_drm, packed)
            continue
        _check_supported(X, left_drm)
        _check_supported(X, right_drm)
        Lc = list(get_sketch_method(X, left_drm, device=True)(X))
        Rc = list(get_sketch_method(X, right_drm, device=True)(X))
        om, ps = device_operators(type(X))
        for mu in range(d - 1):
            om(Lc[mu], Rc[mu], tensor=X, mu=mu, out=views[d + mu])
        for mu in range(d):
            ps(Lc[mu - 1] if mu > 0 else None, Rc[mu] if mu < d - 1 else None, tensor=X, mu=mu, out=views[mu])
    return packed, (shape, rL, rR)


def _sequential_device(tensor: Tensor, left_drm: Optional[DRM], right_drm: DRM, method: SketchMethod):
    """orthogonal / HMT: Psi_mu depends on the QR of Psi_{mu-1}, so bonds are processed in order;
    all intermediates stay on the device.  Returns (Psi, Omega) as lists of device tensors."""
    shape = tuple(tensor.shape)
    d = len(shape)
    parts = _summands(tensor)
    rR = tuple(right_drm.bond_rank)
    for X in parts:
        _check_supported(X, right_drm)
    Rc = [list(get_sketch_method(X, right_drm, device=True)(X)) for X in parts]
    Omega = []
    if method == SketchMethod.orthogonal:
        rL = tuple(left_drm.bond_rank)
        for X in parts:
            _check_supported(X, left_drm)
        Lc = [list(get_sketch_method(X, left_drm, device=True)(X)) for X in parts]
        for mu in range(d - 1):
            o = be.zeros((rL[mu], rR[mu]))
            for s, X in enumerate(parts):
                device_operators(type(X))[0](Lc[s][mu], Rc[s][mu], tensor=X, mu=mu, out=o)
            Omega.append(o)
        del Lc
    else:
        rL = rR  # HMT: the left rank is only needed for shapes (reference :220-222)
    left_psi = OrthogTTDRM(rL, tensor)
    Psi = []
    for mu in range(d):
        r1 = rL[mu - 1] if mu > 0 else 1
        r2 = rR[mu] if mu < d - 1 else 1
        lefts = [None] * len(parts)
        if mu > 0:
            left_psi.add_core(Psi[-1])
            lefts = next(left_psi)
        P = be.zeros((r1, shape[mu], r2))
        for s, X in enumerate(parts):
            device_operators(type(X))[1](lefts[s], Rc[s][mu] if mu < d - 1 else None, tensor=X, mu=mu, out=P)
        if mu < d - 1:
            P = orth_step_device(P, Omega[mu] if method == SketchMethod.orthogonal else None)
        Psi.append(P)
    return Psi, Omega


# ------------------------------------------------------------------ CUDA graphs for the sequential loops
# An orthogonal / HMT sketch of TT / CP / Tucker / dense input is a fixed chain of several hundred small launches
# (C2: 645) whose cost is the launch path, not the kernels.  The chain depends only on the device arrays of the
# input and of the DRMs, so it is captured once per (input, DRMs, method) into a CUDA graph and replayed by later
# calls with the same objects (streaming updates, repeated sketches of one operand).  Sparse summands are excluded:
# their passes size workspaces from the data.  TTSK_GRAPHS=0 (or `use_graphs(False)`) turns this off.
import os as _os
from collections import OrderedDict as _OrderedDict

_USE_GRAPHS = _os.environ.get("TTSK_GRAPHS", "1") != "0"
_SEQ_GRAPHS: "_OrderedDict" = _OrderedDict()
_SEQ_GRAPHS_MAX = 8
graph_stats = {"captured": 0, "replayed": 0, "eager": 0}


def use_graphs(flag: bool) -> None:
    global _USE_GRAPHS
    _USE_GRAPHS = bool(flag)
    if not flag:
        _SEQ_GRAPHS.clear()


def _device_ids(x) -> tuple:
    if isinstance(x, dict):
        return tuple(_device_ids(v) for v in x.values())
    if isinstance(x, (list, tuple)):
        return tuple(_device_ids(v) for v in x)
    return (id(x),)


def _drm_key(drm: Optional[DRM]):
    if drm is None:
        return None
    cores = getattr(drm, "cores", None)
    mats = getattr(drm, "sketching_mats", None)
    # by content, not by object: a rank slice made for one call (blocked sketches) shares its parent's core arrays
    return (type(drm).__name__, bool(drm.transpose), int(drm.seed), tuple(drm.shape), tuple(drm.rank_min),
            tuple(drm.rank_max), tuple(drm.true_rank),
            tuple(id(c) for c in cores) if cores is not None else None,
            tuple(id(m) for m in mats) if mats is not None else None)


def _graphed(kind: str, tensor: Tensor, left_drm: Optional[DRM], right_drm: DRM, fn: Callable):
    """`fn()` (device tensors out) through a cached CUDA graph, or None when the call is not graphable (then the caller
    runs `fn` eagerly).  A chain is captured the SECOND time its key -- the device arrays of the input, the DRMs, the
    method -- is seen, so a one-off sketch never pays for a capture; an input that is refreshed in place
    (`invalidate_device()` + same shapes re-uploads into the same device arrays) keeps its graph."""
    import torch

    parts = _summands(tensor)
    if not _USE_GRAPHS or not parts or any(isinstance(X, SparseTensor) or not _is_stock(type(X)) for X in parts):
        return None
    dev = [X.device() for X in parts]  # uploads (if any) happen here, outside the capture
    gen = int(be.lib().ttsk_workspace_generation(be.ctx()))
    key = (kind, be.device_index(), _device_ids(dev), _drm_key(left_drm), _drm_key(right_drm))
    ent = _SEQ_GRAPHS.get(key)
    if ent is None or ent["gen"] != gen:
        # first sighting (or the workspace arena moved): run eagerly -- this is also the warm-up a capture needs (DRM
        # cores uploaded, arena at its size, kernel attributes set)
        _SEQ_GRAPHS[key] = {"gen": None, "graph": None, "seen": 1, "keep": (tensor, left_drm, right_drm, dev)}
        while len(_SEQ_GRAPHS) > _SEQ_GRAPHS_MAX:
            _SEQ_GRAPHS.popitem(last=False)
        out = fn()
        _SEQ_GRAPHS[key]["gen"] = int(be.lib().ttsk_workspace_generation(be.ctx()))
        graph_stats["eager"] += 1
        return out
    _SEQ_GRAPHS.move_to_end(key)
    if ent["graph"] is None:
        if ent.get("failed"):
            return None
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        before = be.launch_count()
        try:
            with torch.cuda.graph(graph):
                out = fn()
        except Exception as exc:  # not capturable on this input: remember that and fall back
            if _os.environ.get("TTSK_DEBUG"):
                import sys as _sys
                import traceback as _tb
                print("[ttsk] CUDA-graph capture failed:", repr(exc), file=_sys.stderr)
                _tb.print_exc()
            torch.cuda.synchronize()
            ent["failed"] = True
            return None
        if int(be.lib().ttsk_workspace_generation(be.ctx())) != gen:
            ent["failed"] = True
            return None  # the arena moved during the capture: the graph is unusable
        ent.update(graph=graph, out=out, launches=be.launch_count() - before, fresh=True)
        graph_stats["captured"] += 1
    ent["graph"].replay()
    if ent.pop("fresh", False):
        pass  # the capture already went through the launch counter once without executing: it stands for this replay
    else:
        be.check(be.lib().ttsk_note_replayed_launches(be.ctx(), ent["launches"]))
    graph_stats["replayed"] += 1
    return ent["out"]


def _sequential_sketch(tensor: Tensor, left_drm: Optional[DRM], right_drm: DRM, method: SketchMethod):
    run = partial(_sequential_device, tensor, left_drm, right_drm, method)
    out = _graphed(method.value, tensor, left_drm, right_drm, run)
    if out is None:
        out = run()
    Psi, Omega = out
    return SketchContainer([be.to_host(p) for p in Psi], [be.to_host(o) for o in Omega])


def streaming_sketch(tensor: Tensor, left_drm: DRM, right_drm: DRM):
    """`streaming_sketch_device` with the launch-bound summands (TT / CP / Tucker / dense: one fixed chain of small
    launches each) replayed as ONE CUDA graph and the sparse summands accumulated into its output buffer afterwards
    (C5: 100 TT summands next to a sparse term).  The returned packed buffer is overwritten by the next call with the
    same operands: consume it (copy to the host, all-reduce, unpack) before sketching again."""
    parts = _summands(tensor)
    fixed = [X for X in parts if not isinstance(X, SparseTensor) and _is_stock(type(X))]
    if not fixed or not _USE_GRAPHS:
        return streaming_sketch_device(tensor, left_drm, right_drm)
    sub = fixed[0] if len(fixed) == 1 else TensorSum(fixed, shape=tuple(tensor.shape))
    run = partial(streaming_sketch_device, sub, left_drm, right_drm)
    out = _graphed("streaming", sub, left_drm, right_drm, run)
    packed, meta = out if out is not None else run()
    rest = [X for X in parts if not any(X is F for F in fixed)]
    if rest:
        streaming_sketch_device(rest[0] if len(rest) == 1 else TensorSum(rest, shape=tuple(tensor.shape)), left_drm, right_drm,
                                packed=packed)
    return packed, meta


def general_sketch(tensor: Tensor, left_drm: Optional[DRM], right_drm: DRM, method: SketchMethod) -> SketchContainer:
    """Sketch `tensor` with the given DRMs; returns host arrays in a SketchContainer."""
    if method != SketchMethod.hmt and left_drm is None:
        raise ValueError(f"left_drm must be provided for method '{method}'")
    if method == SketchMethod.streaming:
        packed, (shape, rL, rR) = streaming_sketch(tensor, left_drm, right_drm)
        return SketchContainer.unpack(be.to_host_pinned(packed), shape, rL, rR, copy=False)
    return _sequential_sketch(tensor, left_drm, right_drm, method)
