"""SURVEY section 8(f) rank 2 on the GPU: TuckerTensor input under TT-DRMs and the DenseGaussianDRM, against golden
vectors written by the UNMODIFIED reference (tests/golden/tucker_dense_gauss.npz, make_golden.py::gen_tucker_dense_gauss).
Tolerances as in test_gpu_parity.py: Psi / Omega and DRM contractions relative max-norm <= 1e-10; assembled /
orthogonalised TTs compared as reconstructed tensors (relative Frobenius <= 1e-9)."""
import numpy as np
import pytest

from _golden import load, rel_err, stored_list, tensor_desc
from _product import make_tensor

pytestmark = pytest.mark.gpu
TOL = 1e-10
TT_TOL = 1e-9
Z = load("tucker_dense_gauss.npz")


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no CPU fallback exists)"


def _dense(cores):
    from oracle.sketch_oracle import to_dense
    return to_dense(("tt", list(cores)))


def _tt_close(got, want, tol=TT_TOL):
    a, b = _dense(got), _dense(want)
    assert a.shape == b.shape
    assert np.linalg.norm(a - b) <= tol * np.linalg.norm(b), np.linalg.norm(a - b) / np.linalg.norm(b)


def _ttdrm(z, prefix, rank, shape, right):
    from tt_sketch.drm import TensorTrainDRM

    cores = [np.ascontiguousarray(c) for c in stored_list(z, prefix + "_core")]
    return TensorTrainDRM(rank, shape=shape, transpose=right, seed=int(z[prefix + "_seed"]), cores=cores)


@pytest.mark.parametrize("name", [str(n) for n in Z["tucker_names"]])
def test_tucker_sketches_vs_reference_golden(name):
    """stream / orthogonal / HMT sketch of a TuckerTensor and the operator-level sketch_tucker contractions
    (reference tucker_sketch.py:9-46, tensor_train_drm.py:124-145)."""
    from tt_sketch.sketch import hmt_sketch, orthogonal_sketch, stream_sketch
    from tt_sketch.sketch_dispatch import OMEGA_METHODS, PSI_METHODS, get_sketch_method
    from tt_sketch.tensor import TuckerTensor

    z = Z
    X = TuckerTensor([np.ascontiguousarray(u) for u in stored_list(z, name + "_U")], np.ascontiguousarray(z[name + "_core"]))
    assert rel_err(X.to_numpy(), z[name + "_dense"]) < 1e-13
    assert rel_err(X.T.to_numpy(), np.transpose(z[name + "_dense"])) < 1e-13
    lrank = tuple(int(x) for x in z[name + "_lrank"]); rrank = tuple(int(x) for x in z[name + "_rrank"])
    L, R = _ttdrm(z, name + "_L", lrank, X.shape, False), _ttdrm(z, name + "_R", rrank, X.shape, True)
    stt = stream_sketch(X, lrank, rrank, left_drm=L, right_drm=R)
    for a, b in zip(stt.Psi_cores, stored_list(z, name + "_stream_Psi")):
        assert rel_err(a, b) < TOL
    for a, b in zip(stt.Omega_mats, stored_list(z, name + "_stream_Omega")):
        assert rel_err(a, b) < TOL
    _tt_close(stt.C_cores(), stored_list(z, name + "_stream_C"))
    Lc, Rc = list(get_sketch_method(X, L)(X)), list(get_sketch_method(X, R)(X))
    for got, want in ((Lc, stored_list(z, name + "_Lc")), (Rc, stored_list(z, name + "_Rc"))):
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert rel_err(a, b) < TOL
    # the NumPy-in / NumPy-out registry entries on the reference's own contractions
    d = len(X.shape)
    wl, wr = stored_list(z, name + "_Lc"), stored_list(z, name + "_Rc")
    for mu in range(d - 1):
        assert rel_err(OMEGA_METHODS[TuckerTensor](wl[mu], wr[mu], tensor=X, mu=mu), z[f"{name}_stream_Omega{mu}"]) < TOL
    for mu in range(d):
        got = PSI_METHODS[TuckerTensor](wl[mu - 1] if mu > 0 else None, wr[mu] if mu < d - 1 else None, tensor=X, mu=mu)
        assert rel_err(got, z[f"{name}_stream_Psi{mu}"]) < TOL
    _tt_close(orthogonal_sketch(X, lrank, rrank, left_drm=L, right_drm=R).cores, stored_list(z, name + "_orth_C"))
    _tt_close(hmt_sketch(X, rrank, drm=R).cores, stored_list(z, name + "_hmt_C"))


@pytest.mark.parametrize("n_dims,rank,method", [(2, 2, "stream"), (3, 3, "stream"), (3, 2, "orth"), (3, 3, "hmt")])
def test_tucker_exact_recovery(n_dims, rank, method):
    """reference tests/test_sketching_matrix.py:523-545: sketch ranks above the TT rank of a Tucker tensor recover it."""
    from tt_sketch.drm import TensorTrainDRM
    from tt_sketch.sketch import hmt_sketch, orthogonal_sketch, stream_sketch
    from tt_sketch.tensor import TuckerTensor

    shape = tuple(range(10, 10 + n_dims))
    X = TuckerTensor.random(shape, rank, seed=180)
    lrank = tuple(range(rank, rank + n_dims - 1)); rrank = tuple(range(rank + 1, rank + n_dims))
    kw = dict(left_drm_type=TensorTrainDRM, right_drm_type=TensorTrainDRM, seed=180)
    if method == "stream":
        tt = stream_sketch(X, lrank, rrank, **kw).to_tt()
    elif method == "orth":
        tt = orthogonal_sketch(X, lrank, rrank, **kw)
    else:
        tt = hmt_sketch(X, rrank, seed=180, drm_type=TensorTrainDRM)
    dense = X.to_numpy()
    assert np.linalg.norm(tt.to_numpy() - dense) <= 1e-8 * np.linalg.norm(dense)


def test_tucker_in_a_tensor_sum_and_sliced_drm_refused():
    from tt_sketch.drm import TensorTrainDRM
    from tt_sketch.sketch import stream_sketch
    from tt_sketch.tensor import TensorTrain, TuckerTensor

    shape = (7, 8, 9, 10)
    X, Y = TuckerTensor.random(shape, 3, seed=1), TensorTrain.random(shape, 2, seed=2)
    L = TensorTrainDRM((3, 4, 5), shape=shape, transpose=False, seed=3)
    R = TensorTrainDRM((5, 6, 7), shape=shape, transpose=True, seed=4)
    both = stream_sketch(X + Y, (3, 4, 5), (5, 6, 7), left_drm=L, right_drm=R)
    one, two = (stream_sketch(T, (3, 4, 5), (5, 6, 7), left_drm=L, right_drm=R) for T in (X, Y))
    for a, b, c in zip(both.Psi_cores + both.Omega_mats, one.Psi_cores + one.Omega_mats, two.Psi_cores + two.Omega_mats):
        assert rel_err(a, b + c) < 1e-12
    with pytest.raises(ValueError):
        list(L.slice((1, 1, 1), (3, 4, 5)).sketch_tucker(X))


def test_dense_gaussian_drm_vs_reference_golden():
    """DenseGaussianDRM (reference dense_gaussian_drm.py:17-80) on sparse / TT / dense inputs: DRM contractions,
    stream + orthogonal sketches and a blocked sketch against the reference's outputs."""
    from tt_sketch.drm import DenseGaussianDRM
    from tt_sketch.sketch import blocked_stream_sketch, orthogonal_sketch, stream_sketch
    from tt_sketch.sketch_dispatch import get_sketch_method

    z = Z
    lrank = tuple(int(x) for x in z["dg_lrank"]); rrank = tuple(int(x) for x in z["dg_rrank"])
    shape = (5, 6, 7, 4)
    L = DenseGaussianDRM(lrank, shape=shape, transpose=False, seed=11)
    R = DenseGaussianDRM(rrank, shape=shape, transpose=True, seed=23)
    for key in ("dg_sparse", "dg_tt", "dg_dense"):
        X = make_tensor(tensor_desc(z, key + "_T"))
        for drm, side in ((L, "L"), (R, "R")):
            got, want = list(get_sketch_method(X, drm)(X)), stored_list(z, f"{key}_{side}c")
            assert len(got) == len(want)
            for a, b in zip(got, want):
                if key == "dg_tt":
                    assert rel_err(a, b) < TOL
                else:  # gathers / the matrices themselves: exact
                    assert np.array_equal(np.ascontiguousarray(a), b)
        stt = stream_sketch(X, lrank, rrank, left_drm=L, right_drm=R)
        for a, b in zip(stt.Psi_cores, stored_list(z, key + "_stream_Psi")):
            assert rel_err(a, b) < TOL
        for a, b in zip(stt.Omega_mats, stored_list(z, key + "_stream_Omega")):
            assert rel_err(a, b) < TOL
        _tt_close(orthogonal_sketch(X, lrank, rrank, left_drm=L, right_drm=R).cores, stored_list(z, key + "_orth_C"))
    X = make_tensor(tensor_desc(z, "dg_sparse_T"))
    ls = [tuple(int(x) for x in r) for r in z["dg_lslices"]]; rs = [tuple(int(x) for x in r) for r in z["dg_rslices"]]
    sk = blocked_stream_sketch(X, L, R, ls, rs)
    for a, b in zip(sk.Psi_cores, stored_list(z, "dg_sparse_blocked_Psi")):
        assert rel_err(a, b) < TOL
    for a, b in zip(sk.Omega_mats, stored_list(z, "dg_sparse_blocked_Omega")):
        assert rel_err(a, b) < TOL


def test_tt_algebra_after_the_sketch_vs_reference_golden():
    """SURVEY 8(f) rank 3 on the device: TensorTrain.orthogonalize / round / svdvals / dot / norm / error / gather
    (reference tensor.py:414-609; goldens tests/golden/tt_algebra.npz written by the unmodified reference).  Cores of an
    orthogonalised / rounded TT are unique only up to signs and rotations, so tensors, ranks, singular values and
    orthogonality are compared."""
    from tt_sketch import _backend as be

    z = load("tt_algebra.npz")
    a, b, s = (make_tensor(tensor_desc(z, k + "_T")) for k in "abs")
    dense = s.to_numpy()
    # ---- orthogonalize: same tensor, orthonormal columns, and (LAPACK signs) the reference's cores themselves
    o = s.orthogonalize()
    assert o.rank == s.rank
    assert np.linalg.norm(o.to_numpy() - dense) <= 1e-13 * np.linalg.norm(dense)
    for i, c in enumerate(o.cores[:-1]):
        m = c.reshape(-1, c.shape[2])
        assert np.max(np.abs(m.T @ m - np.eye(m.shape[1]))) < 1e-13
        assert rel_err(c, z[f"orth_C{i}"]) < 1e-9
    assert rel_err(o.cores[-1], z[f"orth_C{len(o.cores) - 1}"]) < 1e-9
    # ---- round: by tolerance, by rank, and at machine precision
    for key, kw in (("round_eps", dict(eps=1e-3)), ("round_rank", dict(max_rank=(3, 4, 2))), ("round_exact", dict(eps=1e-12))):
        r = s.round(**kw)
        assert r.rank == tuple(int(x) for x in z[key + "_rank"]), (key, r.rank)
        want = z[key + "_dense"]
        assert np.linalg.norm(r.to_numpy() - want) <= 1e-10 * np.linalg.norm(want), key
        for c in r.cores[1:]:  # left in right-orthogonal form
            m = c.reshape(c.shape[0], -1)
            assert np.max(np.abs(m @ m.T - np.eye(m.shape[0]))) < 1e-12
    # ---- singular values of the unfoldings
    sv = s.svdvals()
    assert len(sv) == 4
    for i, v in enumerate(sv):
        want = z[f"svdvals{i}"]
        assert v.shape == want.shape and np.max(np.abs(v - want)) <= 1e-12 * want[0]
    # ---- scalars
    assert abs(a.dot(b) - float(z["dot_ab"])) <= 1e-13 * max(1.0, abs(float(z["dot_ab"])))
    assert abs(s.norm() - float(z["norm_s"])) <= 1e-13 * float(z["norm_s"])
    assert abs(s.error(a) - float(z["err_sa"])) <= 1e-9 * float(z["err_sa"])
    assert abs(s.error(a, relative=True) - float(z["err_sa_rel"])) <= 1e-9 * float(z["err_sa_rel"])
    assert rel_err(s.gather(z["gather_idx"]), z["gather_s"]) < 1e-13
    # ---- ttsk_svd itself, tall and wide, against NumPy
    rng = np.random.default_rng(0)
    for m, n in ((40, 7), (6, 90), (33, 33)):
        A = rng.standard_normal((m, n))
        U, S, Vt = (be.to_host(t) for t in be.svd(be.to_device(A)))
        assert np.max(np.abs(S - np.linalg.svd(A, compute_uv=False))) < 1e-13 * S[0]
        assert np.max(np.abs((U * S) @ Vt - A)) < 1e-13 * S[0]
        assert np.max(np.abs(U.T @ U - np.eye(len(S)))) < 1e-13 and np.max(np.abs(Vt @ Vt.T - np.eye(len(S)))) < 1e-13
        US = be.to_host(be.svd(be.to_device(A), u_times_s=True)[0])
        assert np.max(np.abs(US - U * S)) < 1e-13 * S[0]
