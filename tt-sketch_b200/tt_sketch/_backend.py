"""ctypes binding of libttsk.so (include/ttsk.h) plus the device-memory plumbing.

PyTorch is used for exactly three things: device allocations (its caching allocator), the
current CUDA stream, and `torch.distributed` -- never for arithmetic.  All arithmetic goes
through the `ttsk_*` entry points below.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_int, c_int32, c_int64,
                    c_uint64, c_void_p)

import numpy as np

MAX_ORDER = 16
_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("TTSK_LIB", os.path.join(_PKG_ROOT, "libttsk.so"))

DRM_GAUSS, DRM_TT = 1, 2


class TtskDrm(Structure):
    """`ttsk_drm` of include/ttsk.h."""
    _fields_ = [
        ("kind", c_int32),
        ("right", c_int32),
        ("seed", c_uint64),
        ("rank_min", c_int32 * MAX_ORDER),
        ("rank_max", c_int32 * MAX_ORDER),
        ("d_cores", c_void_p * MAX_ORDER),
        ("core_r0", c_int32 * MAX_ORDER),
        ("core_r1", c_int32 * MAX_ORDER),
    ]


# name -> (restype, argtypes); every symbol include/ttsk.h declares
SIGNATURES = {
    "ttsk_version": (c_int, []),
    "ttsk_last_error": (c_char_p, []),
    "ttsk_device_count": (c_int, [POINTER(c_int)]),
    "ttsk_create": (c_int, [c_int, POINTER(c_void_p)]),
    "ttsk_destroy": (c_int, [c_void_p]),
    "ttsk_malloc": (c_int, [c_void_p, c_int64, POINTER(c_void_p)]),
    "ttsk_free": (c_int, [c_void_p, c_void_p]),
    "ttsk_malloc_host": (c_int, [c_void_p, c_int64, POINTER(c_void_p)]),
    "ttsk_free_host": (c_int, [c_void_p, c_void_p]),
    "ttsk_memcpy_h2d": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "ttsk_memcpy_d2h": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "ttsk_memset_zero": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "ttsk_sync": (c_int, [c_void_p, c_void_p]),
    "ttsk_launch_count": (c_int64, [c_void_p]),
    "ttsk_note_replayed_launches": (c_int, [c_void_p, c_int64]),
    "ttsk_last_kernel_ms": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double)]),
    "ttsk_last_pass_ms": (c_int, [c_void_p, POINTER(c_double), c_int, POINTER(c_int)]),
    "ttsk_sg_pass_count": (c_int64, [c_void_p]),
    "ttsk_set_table_cache_cap": (c_int, [c_void_p, c_int64]),
    "ttsk_table_cache_bytes": (c_int64, [c_void_p]),
    "ttsk_set_stage_nnz": (c_int, [c_void_p, c_int64]),
    "ttsk_workspace_generation": (c_int64, [c_void_p]),
    "ttsk_trim": (c_int, [c_void_p]),
    "ttsk_lazy_gaussian": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, POINTER(c_int64), c_int, c_int,
                                   c_uint64, c_void_p, c_void_p]),
    "ttsk_lazy_sparse_sign": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, POINTER(c_int64), c_int, c_int, c_int, c_int,
                                      c_uint64, c_void_p, c_void_p]),
    "ttsk_selftest_div": (c_int, [c_void_p, c_int64, c_uint64, POINTER(c_uint64)]),
    "ttsk_selftest_sqrt": (c_int, [c_void_p, c_int64, c_uint64, POINTER(c_uint64)]),
    "ttsk_sketch_size": (c_int64, [c_int, POINTER(c_int64), POINTER(c_int32), POINTER(c_int32)]),
    "ttsk_sparse_sketch": (c_int, [c_void_p, c_int, POINTER(c_int64), c_int64, c_void_p, c_int64, c_void_p,
                                   POINTER(TtskDrm), POINTER(TtskDrm), c_void_p, c_int, c_void_p]),
    "ttsk_tt_sketch": (c_int, [c_void_p, c_int, POINTER(c_int64), POINTER(c_int32), POINTER(c_void_p), POINTER(TtskDrm),
                               POINTER(TtskDrm), c_void_p, c_void_p]),
    "ttsk_dense_sketch": (c_int, [c_void_p, c_int, POINTER(c_int64), c_void_p, POINTER(TtskDrm), POINTER(TtskDrm), c_void_p,
                                  c_void_p]),
    "ttsk_sparse_sketch_host": (c_int, [c_void_p, c_int, POINTER(c_int64), c_int64, c_void_p, c_int64, c_void_p,
                                        POINTER(TtskDrm), POINTER(TtskDrm), c_void_p, c_int]),
    "ttsk_sparse_sketch_stream": (c_int, [c_void_p, c_int, POINTER(c_int64), c_int64, c_void_p, c_int64, c_void_p,
                                          POINTER(TtskDrm), POINTER(TtskDrm), c_void_p, c_int]),
    "ttsk_sparse_omega": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_int,
                                  c_int64, c_int64, c_void_p, c_void_p]),
    "ttsk_sparse_psi": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p]),
    "ttsk_ttdrm_sparse_step": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int,
                                       c_void_p, c_void_p]),
    "ttsk_gemm": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_double, c_void_p, c_int64, c_int64, c_void_p,
                          c_int64, c_int64, c_double, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                          c_void_p]),
    "ttsk_khatri_rao": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                c_void_p]),
    "ttsk_pinv": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_void_p, c_void_p]),
    "ttsk_qr_q": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "ttsk_tns_count": (c_int64, [c_char_p, c_int64, c_void_p]),
    "ttsk_tns_parse": (c_int, [c_char_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "ttsk_svd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
}

_lib = None
_ctx = {}


class TtskError(RuntimeError):
    pass


def lib():
    """Load libttsk.so; raises ImportError loudly when the CUDA extension was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build the CUDA extension first "
                "(python tt-sketch_b200/build.py or __graft_entry__.build()); there is no CPU fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        msg = lib().ttsk_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"ttsk: {msg}")
        raise TtskError(f"ttsk error {rc}: {msg}")


def _torch():
    import torch

    return torch


def device_index() -> int:
    torch = _torch()
    if not torch.cuda.is_available():
        raise TtskError("no CUDA device: tt_sketch (B200 build) has no CPU fallback")
    return torch.cuda.current_device()


def ctx():
    """Per-device library context (created lazily)."""
    dev = device_index()
    if dev not in _ctx:
        h = c_void_p()
        check(lib().ttsk_create(dev, byref(h)))
        _ctx[dev] = h
    return _ctx[dev]


def stream():
    torch = _torch()
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(lib().ttsk_launch_count(ctx()))


# ---------------------------------------------------------------- device arrays (torch plumbing)
def empty(shape, dtype=None):
    torch = _torch()
    return torch.empty(shape, dtype=dtype or torch.float64, device=f"cuda:{device_index()}")


def zeros(shape, dtype=None):
    torch = _torch()
    return torch.zeros(shape, dtype=dtype or torch.float64, device=f"cuda:{device_index()}")


def to_device(arr, dtype=None):
    """NumPy (or torch) -> contiguous device tensor."""
    torch = _torch()
    if isinstance(arr, torch.Tensor):
        t = arr.to(f"cuda:{device_index()}")
        return t if dtype is None else t.to(dtype)
    a = np.ascontiguousarray(arr, dtype=dtype)
    return torch.from_numpy(a).to(f"cuda:{device_index()}")


def to_host(t) -> np.ndarray:
    return t.detach().cpu().numpy()


def to_host_pinned(t) -> np.ndarray:
    """Device -> host through a pinned buffer from torch's caching host allocator (full PCIe rate
    instead of the driver's pageable staging).  The returned array owns the pinned block through
    its base tensor, so the block is only recycled once the caller drops the array."""
    torch = _torch()
    if t.numel() * t.element_size() < (1 << 20):
        return to_host(t)
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def is_device(x) -> bool:
    torch = _torch()
    return isinstance(x, torch.Tensor)


def as_i64(seq):
    return (c_int64 * len(seq))(*[int(x) for x in seq])


def as_i32(seq):
    return (c_int32 * len(seq))(*[int(x) for x in seq])


# ---------------------------------------------------------------- thin op wrappers
def gemm(A, B, out=None, beta: float = 0.0, alpha: float = 1.0):
    """out = alpha * A @ B + beta * out for 2-D device tensors with arbitrary strides."""
    M, K = A.shape
    K2, N = B.shape
    if K != K2:
        raise ValueError(f"gemm: inner dimensions differ ({A.shape} @ {B.shape})")
    if out is None:
        out = empty((M, N))
        beta = 0.0
    if tuple(out.shape) != (M, N):
        raise ValueError("gemm: bad output shape")
    check(lib().ttsk_gemm(ctx(), M, N, K, alpha, ptr(A), A.stride(0), A.stride(1), ptr(B), B.stride(0), B.stride(1),
                          beta, ptr(out), out.stride(0), out.stride(1), 1, 0, 0, 0, stream()))
    return out


def gemm_batched(A, B, out, beta: float = 0.0, alpha: float = 1.0):
    """3-D batched variant: out[b] = alpha * A[b] @ B[b] + beta * out[b] (strided views)."""
    nb, M, K = A.shape
    _, _, N = B.shape
    check(lib().ttsk_gemm(ctx(), M, N, K, alpha, ptr(A), A.stride(1), A.stride(2), ptr(B), B.stride(1), B.stride(2),
                          beta, ptr(out), out.stride(1), out.stride(2), nb, A.stride(0), B.stride(0), out.stride(0),
                          stream()))
    return out


def khatri_rao(A, Rm):
    """out[j, k, m] = A[k, j] * Rm[j, m]; A (n, R) contiguous, Rm (R, r) row-strided."""
    n, R = A.shape
    r = Rm.shape[1]
    if not A.is_contiguous():
        A = A.contiguous()
    if Rm.stride(1) != 1:
        Rm = Rm.contiguous()
    out = empty((R, n, r))
    check(lib().ttsk_khatri_rao(ctx(), n, R, r, ptr(A), ptr(Rm), Rm.stride(0), ptr(out), stream()))
    return out


def pinv(A, rcond: float = -1.0):
    """Pseudo-inverse (n, m) of a small device matrix A (m, n)."""
    if not A.is_contiguous():
        A = A.contiguous()
    m, n = A.shape
    out = empty((n, m))
    check(lib().ttsk_pinv(ctx(), ptr(A), m, n, rcond, ptr(out), stream()))
    return out


def svd(A, u_times_s: bool = False):
    """Thin SVD of a device matrix (min(m, n) <= 256): (U or U*S, S descending, Vt) as device tensors."""
    if not A.is_contiguous():
        A = A.contiguous()
    m, n = A.shape
    k = min(m, n)
    U, S, Vt = empty((m, k)), empty((k,)), empty((k, n))
    check(lib().ttsk_svd(ctx(), ptr(A), m, n, ptr(U), ptr(S), ptr(Vt), 1 if u_times_s else 0, stream()))
    return U, S, Vt


def qr_q_inplace(A):
    m, n = A.shape
    if not A.is_contiguous():
        raise ValueError("qr_q_inplace needs a contiguous matrix")
    check(lib().ttsk_qr_q(ctx(), ptr(A), m, n, stream()))
    return A


def lazy_gaussian(d_idx, k: int, nnz: int, shape, rank_min: int, rank_max: int, seed: int):
    """Device (nnz, rank) Gaussian rows for the first k index rows of d_idx (k x nnz int64)."""
    out = empty((nnz, rank_max - rank_min))
    check(lib().ttsk_lazy_gaussian(ctx(), ptr(d_idx), d_idx.stride(0), k, nnz, as_i64(shape[:k]), int(rank_min),
                                   int(rank_max), int(seed) % 2**63, ptr(out), stream()))
    return out


def lazy_sparse_sign(d_idx, k: int, nnz: int, shape, rank: int, rank_min: int, rank_max: int, nnz_row: int, seed: int):
    """Device (nnz, rank_max - rank_min) sparse-sign rows for the first k index rows of d_idx (k x nnz int64)."""
    out = empty((nnz, rank_max - rank_min))
    check(lib().ttsk_lazy_sparse_sign(ctx(), ptr(d_idx), d_idx.stride(0), k, nnz, as_i64(shape[:k]), int(rank), int(rank_min),
                                      int(rank_max), int(nnz_row), int(seed) % 2**63, ptr(out), stream()))
    return out


def ttdrm_sparse_step(d_idx_mu, v_in, core, r_out_expected=None):
    """v_out = v_in @ core[:, idx, :] per nonzero; core (r_in, n, r_out) contiguous device."""
    r_in, n, r_out = core.shape
    nnz = d_idx_mu.shape[0]
    out = empty((nnz, r_out))
    if not core.is_contiguous():
        core = core.contiguous()
    if v_in is not None and not v_in.is_contiguous():
        v_in = v_in.contiguous()
    check(lib().ttsk_ttdrm_sparse_step(ctx(), nnz, ptr(d_idx_mu), ptr(v_in), r_in, ptr(core), n, r_out, ptr(out),
                                       stream()))
    return out


def sparse_omega(d_val, L, R, out):
    """out (rL, rR) += (L * val) @ R.T with L, R given as (r, nnz) device views."""
    nnz = d_val.shape[0]
    check(lib().ttsk_sparse_omega(ctx(), nnz, ptr(d_val), ptr(L), L.shape[0], L.stride(1), L.stride(0), ptr(R),
                                  R.shape[0], R.stride(1), R.stride(0), ptr(out), stream()))
    return out


def sparse_psi(d_idx_mu, n_mu, d_val, L, R, out):
    """out (r1, n_mu, r2) += segment sums; L/R are (r, nnz) device views or None."""
    nnz = d_val.shape[0]
    check(lib().ttsk_sparse_psi(
        ctx(), nnz, ptr(d_idx_mu), int(n_mu), ptr(d_val),
        ptr(L), L.shape[0] if L is not None else 1, L.stride(1) if L is not None else 0,
        L.stride(0) if L is not None else 0,
        ptr(R), R.shape[0] if R is not None else 1, R.stride(1) if R is not None else 0,
        R.stride(0) if R is not None else 0, ptr(out), stream()))
    return out


def sketch_size(shape, rL, rR) -> int:
    return int(lib().ttsk_sketch_size(len(shape), as_i64(shape), as_i32(list(rL) + [0]), as_i32(list(rR) + [0])))
