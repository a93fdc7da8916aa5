"""Size-independent properties of the sketch on the GPU (the reference's own test strategy,
tests/test_sketching_matrix.py): linearity, split == unsplit, blocked == unblocked, rank
increase keeps the old block, same seed => same result, exact recovery of low-rank tensors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sparse(shape, nnz, seed):
    from tt_sketch.tensor import SparseTensor
    rng = np.random.default_rng(seed)
    idx = np.stack([rng.integers(0, n, nnz) for n in shape]).astype(np.int64)
    return SparseTensor(shape, idx, rng.standard_normal(nnz))


def _close(a, b, tol=1e-10):
    den = max(np.max(np.abs(b)), 1e-300)
    return np.max(np.abs(a - b)) / den < tol


def _drms(shape, lr, rr, ltype, rtype):
    return (ltype(lr, shape=shape, transpose=False, seed=3), rtype(rr, shape=shape, transpose=True, seed=4))


def test_linearity_and_split_full_size_stream():
    """sketch(X1) + sketch(X2) == sketch(X1 + X2) and split(16) == unsplit at 2e6 nonzeros of the
    C4 shape (a size the CPU oracle cannot finish in seconds)."""
    from tt_sketch.drm import SparseGaussianDRM
    from tt_sketch.sketch import stream_sketch

    shape = (10000, 10000, 10000, 500)
    lr, rr = (20,) * 3, (40,) * 3
    X = _sparse(shape, 2_000_000, 1)
    L, R = _drms(shape, lr, rr, SparseGaussianDRM, SparseGaussianDRM)
    whole = stream_sketch(X, lr, rr, left_drm=L, right_drm=R)
    parts = stream_sketch(X.split(16), lr, rr, left_drm=L, right_drm=R)
    again = stream_sketch(X, lr, rr, left_drm=L, right_drm=R)
    for a, b, c in zip(whole.Psi_cores + whole.Omega_mats, parts.Psi_cores + parts.Omega_mats,
                       again.Psi_cores + again.Omega_mats):
        assert _close(a, b) and _close(a, c, 1e-12)
    halves = X.split(2).tensors
    s1 = stream_sketch(halves[0], lr, rr, left_drm=L, right_drm=R)
    s2 = stream_sketch(halves[1], lr, rr, left_drm=L, right_drm=R)
    both = s1.sketch_ + s2.sketch_
    for a, b in zip(whole.Psi_cores + whole.Omega_mats, both.Psi_cores + both.Omega_mats):
        assert _close(a, b)
    upd = s1 + halves[1]  # streaming update with the stored DRMs
    for a, b in zip(whole.Psi_cores, upd.Psi_cores):
        assert _close(a, b)


def test_host_streaming_call_equals_device_call():
    """ttsk_sparse_sketch_host (graduated chunks through pinned staging: every chunk buckets, passes and flushes its
    own segments) == ttsk_sparse_sketch on device-resident COO: two different tilings of the same sums, including
    the segment-GEMM and the unbucketed last-mode forms (shape (3000, 3000, 40, 50), 3e6 nonzeros -> chunks of 1e6 and 2e6)."""
    import torch
    from ctypes import byref
    from tt_sketch import _backend as be
    from tt_sketch.drm import SparseGaussianDRM
    from tt_sketch.sketch_container import SketchContainer
    from tt_sketch.sketch_dispatch import drm_descriptor

    shape = (3000, 3000, 40, 50)
    lr, rr = (20,) * 3, (40,) * 3
    nnz = 3_000_000
    X = _sparse(shape, nnz, 11)
    L, R = _drms(shape, lr, rr, SparseGaussianDRM, SparseGaussianDRM)
    ld, _k1 = drm_descriptor(L)
    rd, _k2 = drm_descriptor(R)
    _, total = SketchContainer.layout(shape, lr, rr)
    lib, ctx = be.lib(), be.ctx()
    d_idx, d_val = torch.from_numpy(X.indices).cuda(), torch.from_numpy(X.entries).cuda()
    out = torch.empty(total, dtype=torch.float64, device="cuda")
    sg0 = lib.ttsk_sg_pass_count(ctx)
    be.check(lib.ttsk_sparse_sketch(ctx, 4, be.as_i64(shape), nnz, be.ptr(d_idx), d_idx.stride(0), be.ptr(d_val),
                                    byref(ld), byref(rd), be.ptr(out), 0, be.stream()))
    torch.cuda.synchronize()
    assert lib.ttsk_sg_pass_count(ctx) >= sg0 + 2, "segment-GEMM / unbucketed forms were not taken"
    h_idx, h_val = torch.from_numpy(X.indices).pin_memory(), torch.from_numpy(X.entries).pin_memory()
    h_out = torch.empty(total, dtype=torch.float64).pin_memory()
    be.check(lib.ttsk_sparse_sketch_host(ctx, 4, be.as_i64(shape), nnz, h_idx.data_ptr(), h_idx.stride(0),
                                         h_val.data_ptr(), byref(ld), byref(rd), h_out.data_ptr(), 0))
    assert _close(h_out.numpy(), out.cpu().numpy(), tol=1e-11)


@pytest.mark.parametrize("kind", ["gauss", "tt"])
def test_blocked_equals_unblocked_and_rank_increase(kind):
    from tt_sketch.drm import SparseGaussianDRM, TensorTrainDRM
    from tt_sketch.sketch import blocked_stream_sketch, stream_sketch

    shape = (9, 10, 11, 8)
    X = _sparse(shape, 700, 2)
    T = SparseGaussianDRM if kind == "gauss" else TensorTrainDRM
    lr, rr = (4, 5, 6), (6, 7, 8)
    L, R = _drms(shape, lr, rr, T, T)
    full = stream_sketch(X, lr, rr, left_drm=L, right_drm=R)
    blk = blocked_stream_sketch(X, L, R, [(0, 0, 0), (1, 2, 3), (3, 3, 4), lr], [(0, 0, 0), (2, 4, 5), rr])
    for a, b in zip(full.Psi_cores + full.Omega_mats, blk.Psi_cores + blk.Omega_mats):
        assert _close(a, b)
    if kind == "gauss":
        nl, nr = (6, 7, 8), (9, 10, 11)
        inc = full.increase_rank(X, nl, nr)
        direct = stream_sketch(X, nl, nr, left_drm=L.increase_rank(nl), right_drm=R.increase_rank(nr))
        for a, b in zip(inc.Psi_cores + inc.Omega_mats, direct.Psi_cores + direct.Omega_mats):
            assert a.shape == b.shape and _close(a, b)
        lp, rp = (1,) + lr, rr + (1,)
        for i, (a, b) in enumerate(zip(full.Psi_cores, inc.Psi_cores)):
            assert _close(a, b[: lp[i], :, : rp[i]])


@pytest.mark.parametrize("method", ["stream", "orth", "hmt"])
@pytest.mark.parametrize("fmt", ["tt", "cp", "sparse", "dense", "sum"])
def test_exact_recovery(fmt, method):
    """A tensor of exact TT rank <= sketch rank is reproduced (reference :208-254, error < 1e-8)."""
    from tt_sketch.sketch import hmt_sketch, orthogonal_sketch, stream_sketch
    from tt_sketch.tensor import CPTensor, DenseTensor, TensorTrain

    shape = (5, 6, 7, 4)
    base = TensorTrain.random(shape, 3, seed=11)
    dense = base.to_numpy()
    rank = 3
    if fmt == "tt":
        X = base
    elif fmt == "cp":
        X = CPTensor.random(shape, 3, seed=12)
        dense = X.to_numpy()
    elif fmt == "sparse":
        X = DenseTensor(dense).to_sparse()
    elif fmt == "dense":
        X = DenseTensor(dense)
    else:
        other = CPTensor.random(shape, 2, seed=13)
        X = base + other + DenseTensor(dense).to_sparse()
        dense = 2 * dense + other.to_numpy()
        rank = 5
    # left rank == exact TT rank (Omega has full row rank, like the reference's tests :269-306);
    # an over-sized left rank makes Omega numerically singular and Omega^+ amplifies rounding.
    lr, rr = (rank,) * 3, (2 * rank,) * 3
    if method == "stream":
        tt = stream_sketch(X, lr, rr, seed=5).to_tt()
    elif method == "orth":
        tt = orthogonal_sketch(X, lr, rr, seed=5)
    else:
        tt = hmt_sketch(X, (rank + 2,) * 3, seed=5)
    err = np.linalg.norm(tt.to_numpy() - dense) / np.linalg.norm(dense)
    assert err < 1e-8, err
    if method == "orth":  # cores are left-orthogonal
        for c in tt.cores[:-1]:
            m = c.reshape(-1, c.shape[2])
            assert np.allclose(m.T @ m, np.eye(m.shape[1]), atol=1e-10)


def test_errors_raise_like_reference():
    from tt_sketch.drm import SparseGaussianDRM, TensorTrainDRM
    from tt_sketch.sketch import stream_sketch
    from tt_sketch.tensor import TensorTrain

    X = _sparse((5, 6, 7), 50, 1)
    L = SparseGaussianDRM((2, 3), shape=(5, 6, 8), transpose=False, seed=1)
    R = SparseGaussianDRM((3, 4), shape=(5, 6, 7), transpose=True, seed=1)
    with pytest.raises(ValueError):
        stream_sketch(X, (2, 3), (3, 4), left_drm=L, right_drm=R)
    L2 = SparseGaussianDRM((2, 3), shape=(5, 6, 7), transpose=False, seed=1)
    with pytest.raises(ValueError):
        stream_sketch(X, (2, 2), (3, 4), left_drm=L2, right_drm=R)  # DRM rank != requested
    tt = TensorTrain.random((5, 6, 7), 2, seed=1)
    with pytest.raises(AttributeError):  # Gaussian DRM cannot sketch a TT (capability missing)
        stream_sketch(tt, (2, 3), (3, 4), left_drm=L2, right_drm=R)
