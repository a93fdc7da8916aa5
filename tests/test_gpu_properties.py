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
    the segment-GEMM form (shape (3000, 3000, 40, 50), 3e6 nonzeros -> chunks of 1e6 and 2e6)."""
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
    assert lib.ttsk_sg_pass_count(ctx) >= sg0 + 1, "the segment-GEMM form was not taken"
    h_idx, h_val = torch.from_numpy(X.indices).pin_memory(), torch.from_numpy(X.entries).pin_memory()
    h_out = torch.empty(total, dtype=torch.float64).pin_memory()
    be.check(lib.ttsk_sparse_sketch_host(ctx, 4, be.as_i64(shape), nnz, h_idx.data_ptr(), h_idx.stride(0),
                                         h_val.data_ptr(), byref(ld), byref(rd), h_out.data_ptr(), 0))
    assert _close(h_out.numpy(), out.cpu().numpy(), tol=1e-11)


def test_unbucketed_last_mode_vs_oracle(oracle_lib):
    """Narrow left rank (4 columns): the last mode takes the unbucketed form (T[i_mu, :] in shared memory, CAS adds),
    wide ranks take the sorted register-accumulating form; both against the CPU oracle at 1e5 nonzeros."""
    from tt_sketch import _backend as be
    from tt_sketch.drm import SparseGaussianDRM
    from tt_sketch.sketch import stream_sketch

    shape = (300, 300, 40, 50)
    nnz = 100_000
    X = _sparse(shape, nnz, 21)
    for lr, rr, flat in [((4, 4, 4), (6, 6, 6), True), ((12, 12, 12), (16, 16, 16), False)]:
        L, R = _drms(shape, lr, rr, SparseGaussianDRM, SparseGaussianDRM)
        sg0 = be.lib().ttsk_sg_pass_count(be.ctx())
        stt = stream_sketch(X, lr, rr, left_drm=L, right_drm=R)
        taken = be.lib().ttsk_sg_pass_count(be.ctx()) - sg0
        oL = oracle_lib.Drm("gauss", False, shape, (0,) * 3, lr, int(L.seed))
        oR = oracle_lib.Drm("gauss", True, shape, (0,) * 3, rr, int(R.seed))
        Psi, Om = oracle_lib.general_sketch(("sparse", shape, X.indices, X.entries), oL, oR, "streaming", fast_sparse=True)
        for a, b in zip(stt.Psi_cores + stt.Omega_mats, Psi + Om):
            assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
        assert (taken >= 1) if flat else True


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


def test_tt_gather_and_sampled_error_on_device():
    """TensorTrain.gather (reference tensor.py:414-440) through the per-nonzero chain kernel == entries of the dense
    reconstruction; SparseTensor.dot / error(fast=True) against a TT (tensor.py:250-255, 52-87) == dense arithmetic;
    and at a size that cannot be densified (C4 shape, 2e6 nonzeros) the gather is linear in the TT."""
    from tt_sketch.tensor import CPTensor, SparseTensor, TensorTrain

    rng = np.random.default_rng(0)
    shape = (6, 7, 8, 5)
    tt = TensorTrain.random(shape, (3, 4, 2), seed=5)
    idx = np.stack([rng.integers(0, n, 500) for n in shape]).astype(np.int64)
    dense = tt.to_numpy()
    assert np.max(np.abs(tt.gather(idx) - dense[tuple(idx)])) <= 1e-14 * np.max(np.abs(dense))
    assert np.max(np.abs(tt.gather(tuple(idx)) - dense[tuple(idx)])) <= 1e-14 * np.max(np.abs(dense))
    flat = rng.choice(int(np.prod(shape)), 300, replace=False)
    uidx = np.stack(np.unravel_index(flat, shape)).astype(np.int64)
    sp = SparseTensor(shape, uidx, rng.standard_normal(300))
    assert abs(sp.dot(tt) - float(np.sum(sp.to_numpy() * dense))) <= 1e-12 * max(1.0, abs(sp.dot(tt)))
    exact = float(np.linalg.norm(sp.to_numpy() - dense))
    assert abs(sp.error(tt, fast=True) - exact) <= 1e-7 * exact
    assert abs(sp.error(tt) - exact) <= 1e-12 * exact
    cp = CPTensor.random(shape, 3, seed=8)
    assert abs(sp.dot(cp) - float(np.sum(sp.to_numpy() * cp.to_numpy()))) <= 1e-12
    big = (10000, 10000, 10000, 500)
    X = _sparse(big, 2_000_000, 31)
    a, b = TensorTrain.random(big, 5, seed=1), TensorTrain.random(big, 3, seed=2)
    ga, gb = a.gather(X.indices), b.gather(X.indices)
    # the TT of the sum (block-diagonal cores) gathers to the sum of the gathers
    cat = [np.concatenate([a.cores[0], b.cores[0]], axis=2)]
    for ca, cb in zip(a.cores[1:-1], b.cores[1:-1]):
        blk = np.zeros((ca.shape[0] + cb.shape[0], ca.shape[1], ca.shape[2] + cb.shape[2]))
        blk[:ca.shape[0], :, :ca.shape[2]] = ca
        blk[ca.shape[0]:, :, ca.shape[2]:] = cb
        cat.append(blk)
    cat.append(np.concatenate([a.cores[-1], b.cores[-1]], axis=0))
    gs = TensorTrain(cat).gather(X.indices)
    assert np.max(np.abs(gs - (ga + gb))) <= 1e-12 * np.max(np.abs(ga + gb))


def test_public_operator_registry_is_the_plug_in_point():
    """Replacing an entry of OMEGA_METHODS / PSI_METHODS (the reference's plug-in point, sketch_dispatch.py:59-82)
    changes what stream_sketch / orthogonal_sketch run; restoring it restores the device path; a new tensor class
    registered with NumPy-level operators and a host-only DRM method is sketched too."""
    from tt_sketch import sketch_dispatch as sd
    from tt_sketch.drm import TensorTrainDRM
    from tt_sketch.sketch import orthogonal_sketch, stream_sketch
    from tt_sketch.tensor import DenseTensor

    shape = (5, 6, 7)
    X = DenseTensor(np.random.default_rng(3).standard_normal(shape))
    lr, rr = (3, 4), (4, 5)
    L = TensorTrainDRM(lr, shape=shape, transpose=False, seed=1)
    R = TensorTrainDRM(rr, shape=shape, transpose=True, seed=2)
    base = stream_sketch(X, lr, rr, left_drm=L, right_drm=R)
    calls = {"omega": 0, "psi": 0}
    stock_o, stock_p = sd.OMEGA_METHODS[DenseTensor], sd.PSI_METHODS[DenseTensor]

    def my_omega(left, right, **kw):
        calls["omega"] += 1
        return 2.0 * stock_o(left, right, **kw)

    def my_psi(left, right, **kw):
        calls["psi"] += 1
        return 2.0 * stock_p(left, right, **kw)

    sd.OMEGA_METHODS[DenseTensor], sd.PSI_METHODS[DenseTensor] = my_omega, my_psi
    try:
        doubled = stream_sketch(X, lr, rr, left_drm=L, right_drm=R)
        orthogonal_sketch(X, lr, rr, left_drm=L, right_drm=R)
    finally:
        sd.OMEGA_METHODS[DenseTensor], sd.PSI_METHODS[DenseTensor] = stock_o, stock_p
    assert calls == {"omega": 4, "psi": 6}
    for a, b in zip(doubled.Psi_cores + doubled.Omega_mats, base.Psi_cores + base.Omega_mats):
        assert _close(a, 2.0 * b, tol=1e-12)
    again = stream_sketch(X, lr, rr, left_drm=L, right_drm=R)
    for a, b in zip(again.Psi_cores + again.Omega_mats, base.Psi_cores + base.Omega_mats):
        assert np.array_equal(a, b)

    class Shifted(DenseTensor):  # a user-defined format: X + c, sketched through NumPy-level operators
        pass

    class MyDRM(TensorTrainDRM):
        def sketch_shifted(self, tensor):  # host-only contraction, no *_device twin
            return self.sketch_dense(DenseTensor(tensor.data))

    sd.DRM_SKETCH_METHOD_DISPATCH[Shifted] = "sketch_shifted"
    sd.OMEGA_METHODS[Shifted] = lambda l, r, *, tensor, **kw: stock_o(l, r, tensor=DenseTensor(tensor.data), **kw)
    sd.PSI_METHODS[Shifted] = lambda l, r, *, tensor, **kw: stock_p(l, r, tensor=DenseTensor(tensor.data), **kw)
    try:
        ML = MyDRM(lr, shape=shape, transpose=False, seed=1)
        MR = MyDRM(rr, shape=shape, transpose=True, seed=2)
        got = stream_sketch(Shifted(X.data), lr, rr, left_drm=ML, right_drm=MR)
    finally:
        for reg in (sd.DRM_SKETCH_METHOD_DISPATCH, sd.OMEGA_METHODS, sd.PSI_METHODS):
            reg.pop(Shifted, None)
    for a, b in zip(got.Psi_cores + got.Omega_mats, base.Psi_cores + base.Omega_mats):
        assert _close(a, b, tol=1e-12)


def test_cuda_graph_replay_of_launch_bound_sketches():
    """Sketches of TT / CP / dense input are fixed chains of small launches: the second call with the same objects
    captures the chain into a CUDA graph, later calls replay it.  Replays must equal the eager result bit for bit,
    count their kernel launches, follow an in-place edit of the input (`invalidate_device()` re-uploads into the same
    device arrays) and a replaced core; a one-off call never captures."""
    from tt_sketch import _backend as be
    from tt_sketch import sketch_dispatch as sd
    from tt_sketch.drm import TensorTrainDRM
    from tt_sketch.sketch import hmt_sketch, orthogonal_sketch, stream_sketch
    from tt_sketch.tensor import CPTensor, TensorTrain

    shape = (7, 8, 9, 10)
    lr, rr = (3, 4, 5), (5, 6, 7)
    L = TensorTrainDRM(lr, shape=shape, transpose=False, seed=1)
    R = TensorTrainDRM(rr, shape=shape, transpose=True, seed=2)
    for make, call in [(lambda: TensorTrain.random(shape, (4, 5, 3), seed=6),
                        lambda X: orthogonal_sketch(X, lr, rr, left_drm=L, right_drm=R).cores),
                       (lambda: CPTensor.random(shape, 6, seed=7),
                        lambda X: (lambda s: s.Psi_cores + s.Omega_mats)(stream_sketch(X, lr, rr, left_drm=L, right_drm=R))),
                       (lambda: TensorTrain.random(shape, (4, 5, 3), seed=8) + CPTensor.random(shape, 3, seed=9),
                        lambda X: hmt_sketch(X, rr, drm=R).cores)]:
        X = make()
        sd.use_graphs(False)
        want = call(X)
        sd.use_graphs(True)
        c0 = dict(sd.graph_stats)
        first = call(X)                       # first sighting: eager
        assert sd.graph_stats["captured"] == c0["captured"]
        l0 = be.launch_count()
        second = call(X)                      # second: capture + replay
        l1 = be.launch_count()
        third = call(X)                       # replay only
        l2 = be.launch_count()
        assert sd.graph_stats["captured"] == c0["captured"] + 1 and sd.graph_stats["replayed"] >= c0["replayed"] + 2
        assert l1 - l0 > 0 and l2 - l1 == l1 - l0
        for got in (first, second, third):
            for a, b in zip(got, want):
                assert np.array_equal(a, b)
        # in-place edit of the host arrays + invalidate: same device arrays, same graph, new numbers
        parts = X.tensors if hasattr(X, "tensors") else [X]
        parts[0].cores[1][...] *= 2.0
        parts[0].invalidate_device()
        edited = call(X)
        assert sd.graph_stats["captured"] == c0["captured"] + 1
        sd.use_graphs(False)
        want2 = call(X)
        sd.use_graphs(True)
        for a, b in zip(edited, want2):
            assert np.array_equal(a, b)
        assert not all(np.array_equal(a, b) for a, b in zip(edited, want))
        # a replaced core of another shape: new device arrays, eager again (no stale replay)
        if isinstance(parts[0], TensorTrain):
            r0 = parts[0].cores[0].shape[2]
            parts[0][0] = np.concatenate([parts[0].cores[0], np.zeros((1, shape[0], 1))], axis=2)
            parts[0][1] = np.concatenate([parts[0].cores[1], np.zeros((1,) + parts[0].cores[1].shape[1:])], axis=0)
            assert parts[0].cores[0].shape[2] == r0 + 1
            again = call(X)
            for a, b in zip(again, want2):
                assert _close(a, b, tol=1e-9)
