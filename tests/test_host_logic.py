"""CPU-only tests: the C ABI loads and exports every declared symbol, host-side bookkeeping
(ranks, orientation, seeds, packing, TT-DRM core generation) matches the reference's golden
values.  No kernel is launched here."""
import os
import re

import numpy as np
import pytest

from _golden import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_loads_and_exports_every_declared_symbol():
    from tt_sketch import _backend as be

    lib = be.lib()
    header = open(os.path.join(ROOT, "include", "ttsk.h")).read()
    declared = set(re.findall(r"\b(ttsk_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ttsk.h but not exported"
    assert declared == set(be.SIGNATURES), declared ^ set(be.SIGNATURES)
    assert lib.ttsk_version() == 100


def test_no_device_fails_loudly():
    import torch

    from tt_sketch import _backend as be

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(be.TtskError):
        be.ctx()
    from ctypes import byref, c_void_p
    h = c_void_p()
    assert be.lib().ttsk_create(0, byref(h)) == -4  # TTSK_E_NODEVICE, no CPU fallback


def test_process_and_trim_ranks():
    from tt_sketch.utils import process_tt_rank, trim_ranks

    assert process_tt_rank(3, (4, 5, 6), trim=False) == (3, 3)
    assert process_tt_rank((2, 9), (4, 5, 6), trim=False) == (2, 9)
    with pytest.raises(ValueError):
        process_tt_rank((2, 3, 4), (4, 5, 6), trim=False)
    assert trim_ranks((2, 3, 4, 2), (100, 100, 100)) == (2, 6, 2)
    assert trim_ranks((10, 10, 10), (5, 100)) == (5, 10)
    assert trim_ranks((3, 3, 3, 3), (3, 100, 2)) == (3, 6, 2)


def test_drm_bookkeeping_matches_reference_orientation():
    from tt_sketch.drm import SparseGaussianDRM

    shape = (7, 8, 9, 10)
    left = SparseGaussianDRM((3, 4, 5), shape=shape, transpose=False, seed=2**40 + 5)
    right = SparseGaussianDRM((5, 6, 7), shape=shape, transpose=True, seed=23)
    assert left.seed == (2**40 + 5) % (2**32 - 1) and isinstance(left.seed, int)
    assert left.rank == (3, 4, 5) and right.rank == (7, 6, 5) and right.bond_rank == (5, 6, 7)
    sl = right.slice((1, 2, 3), (5, 6, 7))
    assert sl.rank_min == (3, 2, 1) and sl.rank_max == (7, 6, 5) and sl.bond_rank == (4, 4, 4)
    assert sl.true_rank == right.true_rank and sl.seed == right.seed
    t = right.T
    assert t.transpose is False and t.rank == (5, 6, 7)
    bigger = left.increase_rank((5, 6, 7))
    assert bigger.rank == (5, 6, 7) and bigger.seed == left.seed
    z = load("sketches.npz")
    assert tuple(z["sparse_gauss_R_rank_max"]) == right.rank_max  # reference's internal orientation


def test_ttdrm_cores_match_reference():
    import multiprocessing

    from tt_sketch.drm import TensorTrainDRM

    z = load("ttdrm_cores.npz")
    if int(z["cpu_count"]) != multiprocessing.cpu_count():
        pytest.skip("TT-DRM core values depend on cpu_count() (reference quirk, SURVEY App. B-3)")
    for side, tr in (("l", False), ("r", True)):
        drm = TensorTrainDRM(tuple(int(x) for x in z[side + "_rank"]), shape=tuple(int(x) for x in z[side + "_shape"]),
                             transpose=tr, seed=int(z[side + "_seed"]))
        for i, c in enumerate(drm.cores):
            assert np.array_equal(c, z[f"{side}_core{i}"])


def test_random_normal_thread_layout():
    from oracle.sketch_oracle import multithreaded_normal
    from tt_sketch.utils import MultithreadedRNG

    a = MultithreadedRNG((13, 7), seed=99, threads=4).values
    assert np.array_equal(a, multithreaded_normal((13, 7), 99, threads=4))
    assert not np.array_equal(a, MultithreadedRNG((13, 7), seed=99, threads=3).values)


def test_container_pack_unpack_transpose_scale():
    from tt_sketch.sketch_container import SketchContainer

    shape, rl, rr = (4, 5, 6), (2, 3), (3, 4)
    items, total = SketchContainer.layout(shape, rl, rr)
    assert total == 1 * 4 * 3 + 2 * 5 * 4 + 3 * 6 * 1 + 2 * 3 + 3 * 4
    flat = np.arange(total, dtype=float)
    sk = SketchContainer.unpack(flat, shape, rl, rr)
    assert [p.shape for p in sk.Psi_cores] == [(1, 4, 3), (2, 5, 4), (3, 6, 1)]
    assert [o.shape for o in sk.Omega_mats] == [(2, 3), (3, 4)]
    assert np.array_equal(sk.pack(), flat)
    assert sk.left_rank == rl and sk.right_rank == rr and sk.shape == shape
    two = sk * 2.0
    assert np.array_equal(two.pack(), 2 * flat)  # works here; the reference raises (App. B-7)
    assert np.array_equal((sk + sk).pack(), 2 * flat)
    t = sk.T
    assert t.Psi_cores[0].shape == (1, 6, 3) and t.Omega_mats[0].shape == (4, 3)
    z = SketchContainer.zero(shape, rl, rr)
    assert all(not p.any() for p in z.Psi_cores)


def test_tensor_containers():
    from tt_sketch.tensor import CPTensor, DenseTensor, SparseTensor, TensorSum, TensorTrain

    rng = np.random.default_rng(0)
    tt = TensorTrain.random((3, 4, 5), (2, 3), seed=1)
    assert tt.shape == (3, 4, 5) and tt.rank == (2, 3)
    assert np.allclose(tt.T.to_numpy(), tt.to_numpy().transpose(2, 1, 0))
    cp = CPTensor.random((3, 4, 5), 4, seed=2)
    assert np.allclose(cp.T.to_numpy(), cp.to_numpy().transpose(2, 1, 0))
    dn = DenseTensor(rng.standard_normal((3, 4, 5)))
    sp = dn.to_sparse()
    assert sp.nnz == 60 and np.allclose(sp.to_numpy(), dn.data)
    parts = sp.split(7)
    assert isinstance(parts, TensorSum) and parts.num_summands == 7
    assert sum(p.nnz for p in parts.tensors) == 60 and np.allclose(parts.to_numpy(), dn.data)
    s = tt + cp + dn
    assert isinstance(s, TensorSum) and s.num_summands == 3
    assert np.allclose(s.to_numpy(), tt.to_numpy() + cp.to_numpy() + dn.data)
    assert np.allclose((2 * tt).to_numpy(), 2 * tt.to_numpy())
    assert np.allclose((s * 0.5).to_numpy(), 0.5 * s.to_numpy())
    bad = SparseTensor((3, 3), np.array([[0, 3], [1, 1]]), np.ones(2))
    with pytest.raises(ValueError):
        bad.check_indices()


def test_entry_point_validation_without_gpu():
    from tt_sketch.sketch import blocked_stream_sketch, orthogonal_sketch, stream_sketch
    from tt_sketch.tensor import TensorTrain

    tt = TensorTrain.random((3, 4, 5), 2, seed=1)
    with pytest.raises(ValueError):
        stream_sketch(tt, (2, 3), (3, 2))
    with pytest.raises(ValueError):
        orthogonal_sketch(tt, (3, 3), (2, 4))
    with pytest.raises(ValueError):
        blocked_stream_sketch(tt, object(), object(), [], [])


def test_fused_path_selection_is_host_logic_only():
    """Which summands take the single-call device paths (ttsk_sparse_sketch / ttsk_tt_sketch) is decided from the
    tensor and DRM types alone, without touching the device."""
    from tt_sketch.drm import SparseGaussianDRM, TensorTrainDRM
    from tt_sketch.sketch_dispatch import _fusable, _fusable_tt
    from tt_sketch.tensor import CPTensor, SparseTensor, TensorTrain

    shape = (6, 5, 4)
    tt = TensorTrain.random(shape, 2, seed=1)
    cp = CPTensor.random(shape, 2, seed=1)
    sp = SparseTensor(shape, np.zeros((3, 1), dtype=np.int64), np.ones(1))
    tl = TensorTrainDRM((2, 2), shape=shape, transpose=False, seed=1)
    tr = TensorTrainDRM((3, 3), shape=shape, transpose=True, seed=2)
    gl = SparseGaussianDRM((2, 2), shape=shape, transpose=False, seed=1)
    gr = SparseGaussianDRM((3, 3), shape=shape, transpose=True, seed=2)
    assert _fusable_tt(tt, tl, tr)
    assert not _fusable_tt(cp, tl, tr) and not _fusable_tt(sp, tl, tr)
    assert not _fusable_tt(tt, gl, gr)  # a Gaussian DRM has no TensorTrain contraction (reference: sketch_sparse only)
    assert _fusable(sp, gl, gr) and _fusable(sp, tl, tr) and _fusable(sp, gl, tr)
    assert not _fusable(tt, tl, tr)
    big = TensorTrainDRM((70, 70), shape=(80, 80, 80), transpose=False, seed=1)
    assert not _fusable(SparseTensor((80, 80, 80), np.zeros((3, 1), dtype=np.int64), np.ones(1)), big, big)  # rank > 64


def test_tensor_dot_norm_error_semantics_on_host():
    """Tensor.dot / norm / error(fast) and the gathers that need no device (reference tensor.py:52-131, 250-291,
    542-560, 670-671, 726-732) against dense NumPy arithmetic."""
    from tt_sketch.tensor import CPTensor, DenseTensor, SparseTensor, TensorTrain

    rng = np.random.default_rng(4)
    shape = (5, 6, 7)
    tt, tt2 = TensorTrain.random(shape, 3, seed=1), TensorTrain.random(shape, 2, seed=2)
    cp = CPTensor.random(shape, 4, seed=3)
    dn = DenseTensor(rng.standard_normal(shape))
    assert abs(dn.dot(cp) - float(np.sum(dn.data * cp.to_numpy()))) < 1e-12
    assert abs((cp + tt).dot(dn) - (cp.dot(dn) + tt.dot(dn))) < 1e-12   # (TT . TT runs on the device: GPU tests)
    tsum = tt.add(tt2)
    assert tsum.rank == (5, 5) and np.allclose(tsum.to_numpy(), tt.to_numpy() + tt2.to_numpy(), atol=1e-13)
    flat = rng.choice(int(np.prod(shape)), 40, replace=False)
    idx = np.stack(np.unravel_index(flat, shape)).astype(np.int64)
    sp = SparseTensor(shape, idx, rng.standard_normal(40))
    assert np.allclose(cp.gather(idx), cp.to_numpy()[tuple(idx)])
    probe = np.stack([rng.integers(0, n, 30) for n in shape])
    assert np.allclose(sp.gather(probe), sp.to_numpy()[tuple(probe)])
    assert abs(sp.dot(cp) - float(np.sum(sp.to_numpy() * cp.to_numpy()))) < 1e-13
    exact = float(np.linalg.norm(cp.to_numpy() - dn.data))
    assert abs(cp.error(dn) - exact) < 1e-12 and abs(cp.error(dn, fast=True) - exact) < 1e-7 * exact
    assert abs(cp.error(dn, relative=True) - exact / float(np.linalg.norm(dn.data))) < 1e-12


def test_dense_gaussian_drm_matrices_bit_identical_to_reference():
    """DenseGaussianDRM draws its matrices on the host (legacy MT19937 stream, reference dense_gaussian_drm.py:36-56):
    bit-identical to the reference's, for the full DRM, a slice and a rank increase; the global NumPy generator is
    left alone."""
    from _golden import load, stored_list
    from tt_sketch.drm import DenseGaussianDRM

    z = load("tucker_dense_gauss.npz")
    shape = (5, 6, 7, 4)
    lrank, rrank = tuple(int(x) for x in z["dg_lrank"]), tuple(int(x) for x in z["dg_rrank"])
    np.random.seed(1234)
    probe = np.random.uniform()
    np.random.seed(1234)
    L = DenseGaussianDRM(lrank, shape=shape, transpose=False, seed=11)
    R = DenseGaussianDRM(rrank, shape=shape, transpose=True, seed=23)
    assert np.random.uniform() == probe
    for drm, key in ((L, "dg_L_mat"), (R, "dg_R_mat"), (L.slice((1, 1, 0), (3, 3, 2)), "dg_Lslice_mat"),
                     (L.increase_rank((4, 5, 4)), "dg_Linc_mat")):
        want = stored_list(z, key)
        assert len(drm.sketching_mats) == len(want)
        for a, b in zip(drm.sketching_mats, want):
            assert a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))
    assert L.rank == lrank and R.rank == rrank[::-1]


def test_tucker_tensor_container_semantics():
    """TuckerTensor (reference tensor.py:746-816): shape / rank / size, transposition, scaling, dense reconstruction and
    the seeded random constructor (orthonormal factor rows)."""
    from tt_sketch.tensor import TuckerTensor

    X = TuckerTensor.random((6, 7, 8), (2, 9, 3), seed=5)
    assert X.shape == (6, 7, 8) and X.rank == (2, 7, 3) and X.size == 2 * 7 * 3 + 2 * 6 + 7 * 7 + 3 * 8
    for U in X.factors:
        assert np.allclose(U @ U.T, np.eye(U.shape[0]), atol=1e-12)
    dense = np.einsum("abc,ai,bj,ck->ijk", X.core, *X.factors)
    assert np.allclose(X.to_numpy(), dense, atol=1e-13)
    assert np.allclose(X.T.to_numpy(), dense.transpose(2, 1, 0), atol=1e-13)
    assert np.allclose((2.5 * X).to_numpy(), 2.5 * dense, atol=1e-13)
    Y = TuckerTensor.random((6, 7, 8), (2, 9, 3), seed=5)
    assert np.array_equal(X.core, Y.core) and all(np.array_equal(a, b) for a, b in zip(X.factors, Y.factors))


def test_frostt_tns_ingestion(tmp_path):
    """FROSTT .tns(.gz) text -> COO through the native parser (reference scripts/frostt.py:51-89): indices 0-based,
    values parsed like Python's float(), comments / blank lines / CRLF / trailing blanks tolerated, the .npz cache,
    and the error cases."""
    import gzip

    from tt_sketch.frostt import get_frostt_tensor, parse_tns, process_frostt_tensor

    rng = np.random.default_rng(7)
    shape = (37, 5, 1200, 9)
    n = 50_000
    idx = np.stack([rng.integers(1, s + 1, n) for s in shape])
    vals = rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n)
    texts = ["%.17g" % v if i % 3 else ("%e" % v if i % 2 else repr(float(v))) for i, v in enumerate(vals)]
    lines = ["# a comment", ""]
    for i in range(n):
        sep = " " if i % 5 else "\t"
        lines.append(sep.join(str(int(x)) for x in idx[:, i]) + sep + texts[i] + ("  " if i % 7 == 0 else "") + ("\r" if i % 11 == 0 else ""))
    body = ("\n".join(lines) + "\n").encode()
    want_vals = np.array([float(t) for t in texts])
    got_idx, got_val, mx = parse_tns(body)
    assert got_idx.shape == (4, n) and np.array_equal(got_idx, idx - 1)
    assert np.array_equal(got_val.view(np.uint64), want_vals.view(np.uint64))
    assert np.array_equal(mx, idx.max(axis=1) - 1)
    assert np.array_equal(parse_tns(body[:-1])[1], got_val)       # no newline at the end of the file
    for name, opener in (("t.tns.gz", gzip.open), ("u.tns", open)):
        path = tmp_path / name
        with opener(path, "wb") as f:
            f.write(body)
        X = process_frostt_tensor(str(path), nnz=n, shape=shape)
        assert X.shape == shape and X.nnz == n and np.array_equal(X.indices, idx - 1) and np.array_equal(X.entries, want_vals)
    assert process_frostt_tensor(str(tmp_path / "u.tns")).shape == tuple(int(m) for m in idx.max(axis=1))
    Y = get_frostt_tensor("https://example.org/frostt/t.tns.gz", n, shape, data_dir=str(tmp_path))
    assert (tmp_path / "t.tns.npz").exists() and np.array_equal(Y.entries, want_vals)
    (tmp_path / "t.tns.gz").unlink()                               # second call: served from the cache
    Z = get_frostt_tensor("t.tns.gz", n, shape, data_dir=str(tmp_path))
    assert Z.shape == shape and np.array_equal(Z.indices, idx - 1)
    with pytest.raises(FileNotFoundError):
        get_frostt_tensor("missing.tns.gz", data_dir=str(tmp_path))
    with pytest.raises(ValueError):
        parse_tns(body, nnz=n + 1)
    for bad in (b"1 2 3\n1 2 x\n", b"0 1 2.0\n", b"1 2 3.0\n1 2\n", b"# only a comment\n"):
        with pytest.raises(ValueError):
            parse_tns(bad)
    with pytest.raises(ValueError):
        process_frostt_tensor(str(tmp_path / "u.tns"), shape=(3, 3, 3, 3))
