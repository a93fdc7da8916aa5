"""GPU parity tests: the CUDA path (through the C ABI, via the tt_sketch host mirror) against
  (1) golden vectors written by the UNMODIFIED reference (tests/golden/), and
  (2) the CPU oracle (oracle/) on the same seeded inputs at sizes it finishes in seconds.
Tolerances: DRM entries bit-exact (uint64 view); Psi / Omega relative max-norm <= 1e-10
(summation order differs); assembled / orthogonalised TTs compared as reconstructed tensors,
relative Frobenius error <= 1e-9 (pseudo-inverse of a nearly singular Omega amplifies rounding).
"""
import numpy as np
import pytest

from _golden import drm_desc, load, rel_err, stored_list, tensor_desc
from _product import make_drm, make_tensor

pytestmark = pytest.mark.gpu
TOL = 1e-10
TT_TOL = 1e-9


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no CPU fallback exists)"


def test_lazy_gaussian_bit_exact_vs_golden():
    from tt_sketch import _backend as be

    z = load("lazy_gaussian.npz")
    for n, row in enumerate(z["cases"]):
        d = int(row[0]); shape = tuple(int(x) for x in row[1:1 + d]); k, rmin, rmax, seed = (int(x) for x in row[5:9])
        idx, want = z[f"c{n}_idx"], z[f"c{n}_out"]
        got = be.to_host(be.lazy_gaussian(be.to_device(idx, np.int64), k, idx.shape[1], shape, rmin, rmax, seed))
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), n


def test_lazy_gaussian_bit_exact_vs_oracle_large(oracle_lib):
    """2e7 variates incl. the int32-wrapping shape and the C4 shape."""
    from tt_sketch import _backend as be

    rng = np.random.default_rng(7)
    for shape, nnz, rmin, rmax, seed in [((10000, 10000, 10000, 500), 250000, 0, 40, 2),
                                         ((70000, 70000, 3), 200000, 3, 23, 179),
                                         ((10, 12, 14, 7), 300000, 5, 25, 5)]:
        idx = np.stack([rng.integers(0, n, nnz) for n in shape]).astype(np.int64)
        for k in (1, len(shape)):
            want = oracle_lib.inds_to_normal(idx[:k], shape[:k], rmin, rmax, seed)
            got = be.to_host(be.lazy_gaussian(be.to_device(idx, np.int64), k, nnz, shape, rmin, rmax, seed))
            bad = int((got.view(np.uint64) != want.view(np.uint64)).sum())
            assert bad == 0, (shape, k, bad)


def test_straight_line_division_is_ieee():
    """div_rn_safe (the branch-free division inside ndtri) == __ddiv_rn on 2e8 operand pairs."""
    from ctypes import byref, c_uint64

    from tt_sketch import _backend as be

    bad = c_uint64(123)
    be.check(be.lib().ttsk_selftest_div(be.ctx(), 200_000_000, 20240917, byref(bad)))
    assert bad.value == 0


def test_straight_line_sqrt_is_ieee():
    """sqrt_rn_safe (the branch-free square root of the tail branch) == __dsqrt_rn on 2e8 operands."""
    from ctypes import byref, c_uint64

    from tt_sketch import _backend as be

    bad = c_uint64(123)
    be.check(be.lib().ttsk_selftest_sqrt(be.ctx(), 200_000_000, 20240918, byref(bad)))
    assert bad.value == 0


def test_ndtri_tail_and_edge_uniforms(oracle_lib):
    """The deferred tail branch is exercised by every sketch; here the Gaussian rows of a sliced
    DRM must equal the same columns of the unsliced one bit-for-bit (reference
    tests/test_fast_lazy_gaussian.py:169-200)."""
    from tt_sketch import _backend as be

    rng = np.random.default_rng(3)
    shape = (10, 12, 14, 7)
    idx = np.stack([rng.integers(0, n, 5000) for n in shape]).astype(np.int64)
    di = be.to_device(idx, np.int64)
    full = be.to_host(be.lazy_gaussian(di, 4, 5000, shape, 0, 30, 11))
    part = be.to_host(be.lazy_gaussian(di, 4, 5000, shape, 12, 25, 11))
    assert np.array_equal(full[:, 12:25].view(np.uint64), part.view(np.uint64))
    perm = rng.permutation(5000)
    shuf = be.to_host(be.lazy_gaussian(be.to_device(idx[:, perm], np.int64), 4, 5000, shape, 0, 30, 11))
    assert np.array_equal(shuf.view(np.uint64), full[perm].view(np.uint64))
    assert abs(full.mean()) < 0.02 and abs(full.std() - 1) < 0.02


SKETCH = load("sketches.npz")
NAMES = [str(n) for n in SKETCH["names"]]


def _dense_of(cores):
    from oracle.sketch_oracle import to_dense
    return to_dense(("tt", list(cores)))


def _tt_close(got_cores, want_cores, tol=TT_TOL):
    a, b = _dense_of(got_cores), _dense_of(want_cores)
    assert a.shape == b.shape
    assert np.linalg.norm(a - b) <= tol * np.linalg.norm(b), np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("name", NAMES)
def test_sketch_vs_reference_golden(oracle_lib, name):
    from tt_sketch.sketch import blocked_stream_sketch, hmt_sketch, orthogonal_sketch, stream_sketch
    from tt_sketch.sketch_dispatch import get_sketch_method

    z = SKETCH
    desc = tensor_desc(z, name + "_T")
    shape = oracle_lib.tshape(desc)
    oL, oR = drm_desc(z, name, "L", shape), drm_desc(z, name, "R", shape)
    X, L, R = make_tensor(desc), make_drm(oL), make_drm(oR)
    lrank = tuple(int(x) for x in z[name + "_lrank"]); rrank = tuple(int(x) for x in z[name + "_rrank"])
    methods = str(z[name + "_methods"]).split(",")
    if "stream" in methods:
        stt = stream_sketch(X, lrank, rrank, left_drm=L, right_drm=R)
        for a, b in zip(stt.Psi_cores, stored_list(z, name + "_stream_Psi")):
            assert rel_err(a, b) < TOL
        for a, b in zip(stt.Omega_mats, stored_list(z, name + "_stream_Omega")):
            assert rel_err(a, b) < TOL
        _tt_close(stt.C_cores(), stored_list(z, name + "_stream_C"))
        _tt_close(stt.C_cores(direction="right"), stt.C_cores(direction="left"), 1e-7)
        if desc[0] != "sum":  # operator-level DRM contractions
            for side, drm, odrm in (("L", L, oL), ("R", R, oR)):
                got = list(get_sketch_method(X, drm)(X))
                want = stored_list(z, f"{name}_{side}c")
                assert len(got) == len(want)
                for a, b in zip(got, want):
                    if odrm.kind == "gauss":
                        assert np.array_equal(np.ascontiguousarray(a).view(np.uint64), b.view(np.uint64))
                    else:
                        assert rel_err(a, b) < TOL
    if "orth" in methods:
        tt = orthogonal_sketch(X, lrank, rrank, left_drm=L, right_drm=R)
        _tt_close(tt.cores, stored_list(z, name + "_orth_C"))
    if "hmt" in methods:
        tt = hmt_sketch(X, rrank, drm=R)
        _tt_close(tt.cores, stored_list(z, name + "_hmt_C"))
    if "blocked" in methods:
        ls = [tuple(int(x) for x in r) for r in z[name + "_lslices"]]
        rs = [tuple(int(x) for x in r) for r in z[name + "_rslices"]]
        sk = blocked_stream_sketch(X, L, R, ls, rs)
        for a, b in zip(sk.Psi_cores, stored_list(z, name + "_blocked_Psi")):
            assert rel_err(a, b) < TOL
        for a, b in zip(sk.Omega_mats, stored_list(z, name + "_blocked_Omega")):
            assert rel_err(a, b) < TOL


@pytest.mark.parametrize("name", [n for n in NAMES if not n.startswith("sum")])
def test_operator_level_plugins_vs_oracle(oracle_lib, name):
    """The NumPy-in/NumPy-out registry entries (OMEGA_METHODS / PSI_METHODS) on the golden inputs."""
    from tt_sketch.sketch_dispatch import OMEGA_METHODS, PSI_METHODS

    z = SKETCH
    desc = tensor_desc(z, name + "_T")
    shape = oracle_lib.tshape(desc)
    d = len(shape)
    oL, oR = drm_desc(z, name, "L", shape), drm_desc(z, name, "R", shape)
    X = make_tensor(desc)
    Lc, Rc = oracle_lib.drm_contractions(oL, desc), oracle_lib.drm_contractions(oR, desc)
    rL, rR = oL.rank, oR.rank
    for mu in range(d - 1):
        want = oracle_lib.omega(desc, Lc[mu], Rc[mu], mu)
        got = OMEGA_METHODS[type(X)](np.ascontiguousarray(Lc[mu]), np.ascontiguousarray(Rc[mu]), tensor=X, mu=mu,
                                     omega_shape=want.shape)
        assert rel_err(got, want) < TOL
    for mu in range(d):
        Lm = Lc[mu - 1] if mu > 0 else None
        Rm = Rc[mu] if mu < d - 1 else None
        pshape = (Lm.shape[0] if (Lm is not None and desc[0] in ("sparse", "dense")) else (Lm.shape[1] if Lm is not None else 1),
                  shape[mu],
                  Rm.shape[0] if (Rm is not None and desc[0] in ("sparse", "dense")) else (Rm.shape[1] if Rm is not None else 1))
        want = oracle_lib.psi(desc, Lm, Rm, mu, pshape)
        got = PSI_METHODS[type(X)](None if Lm is None else np.ascontiguousarray(Lm),
                                   None if Rm is None else np.ascontiguousarray(Rm), tensor=X, mu=mu, psi_shape=pshape)
        assert rel_err(got, want) < TOL


def _c4_like(nnz, seed=0, shape=(10000, 10000, 10000, 500)):
    idx = np.stack([np.random.default_rng(100 + k + seed).integers(0, n, nnz) for k, n in enumerate(shape)]).astype(np.int64)
    val = np.random.default_rng(99 + seed).standard_normal(nnz)
    return shape, idx, val


@pytest.mark.parametrize("kinds,nnz", [(("gauss", "gauss"), 30000), (("tt", "tt"), 30000), (("gauss", "tt"), 30000),
                                       (("gauss", "gauss"), 150000)])
def test_sparse_c4_shape_vs_oracle(oracle_lib, kinds, nnz):
    """BASELINE config 4's shape and ranks (rL=20, rR=40): fused kernel vs oracle.  nnz=3e4 takes the
    global-atomic bucket scatter, nnz=1.5e5 the two-level shared-memory scatter."""
    from oracle.sketch_oracle import Drm
    from tt_sketch.sketch import stream_sketch

    shape, idx, val = _c4_like(nnz)
    d = len(shape)
    rl, rr = (20,) * 3, (40,) * 3
    ocores_l = oracle_lib.tt_drm_cores(shape, rl, 1, False) if kinds[0] == "tt" else []
    ocores_r = oracle_lib.tt_drm_cores(shape, rr, 2, True) if kinds[1] == "tt" else []
    oL = Drm(kinds[0], False, shape, (0,) * 3, rl, 1, ocores_l)
    oR = Drm(kinds[1], True, shape, (0,) * 3, rr, 2, ocores_r)
    desc = ("sparse", shape, idx, val)
    Psi, Om = oracle_lib.general_sketch(desc, oL, oR, "streaming", fast_sparse=True)
    stt = stream_sketch(make_tensor(desc), rl, rr, left_drm=make_drm(oL), right_drm=make_drm(oR))
    for a, b in zip(stt.Psi_cores, Psi):
        assert rel_err(a, b) < TOL
    for a, b in zip(stt.Omega_mats, Om):
        assert rel_err(a, b) < TOL
    assert [p.shape for p in stt.Psi_cores] == [(1, 10000, 40), (20, 10000, 40), (20, 10000, 40), (20, 500, 1)]


@pytest.mark.parametrize("rl,rr", [((20, 20, 20), (40, 40, 40)), ((5, 7, 3), (6, 9, 4))])
def test_sparse_segment_gemm_form_vs_oracle(oracle_lib, rl, rr):
    """Small trailing modes and long segments: mode 2 runs in the segment-GEMM form (T_j^T R tables once per
    segment instead of per-nonzero gathers + MMAs); same oracle, same tolerance."""
    from oracle.sketch_oracle import Drm
    from tt_sketch import _backend as be
    from tt_sketch.sketch import stream_sketch

    shape, nnz = (3000, 3000, 16, 30), 120000
    _, idx, val = _c4_like(nnz, seed=7, shape=shape)
    idx[2, :5000] = 3   # one long segment next to ordinary ones
    oL = Drm("gauss", False, shape, (0,) * 3, rl, 11)
    oR = Drm("gauss", True, shape, (0,) * 3, rr, 12)
    desc = ("sparse", shape, idx, val)
    Psi, Om = oracle_lib.general_sketch(desc, oL, oR, "streaming", fast_sparse=True)
    before = be.lib().ttsk_sg_pass_count(be.ctx())
    stt = stream_sketch(make_tensor(desc), rl, rr, left_drm=make_drm(oL), right_drm=make_drm(oR))
    assert be.lib().ttsk_sg_pass_count(be.ctx()) > before, "segment-GEMM form was not taken"
    for a, b in zip(stt.Psi_cores, Psi):
        assert rel_err(a, b) < TOL
    for a, b in zip(stt.Omega_mats, Om):
        assert rel_err(a, b) < TOL


def test_sparse_duplicates_empty_slices_and_ragged_segments(oracle_lib):
    """Edge cases: duplicate coordinates (summed), slices with no nonzero, one huge segment next
    to singletons, nnz not a multiple of the tile, a single nonzero."""
    from oracle.sketch_oracle import Drm
    from tt_sketch.sketch import stream_sketch

    rng = np.random.default_rng(5)
    shape = (6, 50, 7, 3)
    for nnz in (1, 5, 33, 4099):
        idx = np.stack([rng.integers(0, n, nnz) for n in shape]).astype(np.int64)
        idx[1, : nnz // 2] = 17            # one long segment in mode 1, most other slices empty
        if nnz > 4:
            idx[:, 1] = idx[:, 0]          # duplicate coordinate
        val = rng.standard_normal(nnz)
        desc = ("sparse", shape, idx, val)
        oL = Drm("gauss", False, shape, (0, 0, 0), (2, 5, 9), 31)
        oR = Drm("gauss", True, shape, (0, 0, 0), (4, 11, 3), 32)  # not ordered: use general_sketch directly
        from tt_sketch.sketch_dispatch import SketchMethod, general_sketch
        sk = general_sketch(make_tensor(desc), make_drm(oL), make_drm(oR), SketchMethod.streaming)
        Psi, Om = oracle_lib.general_sketch(desc, oL, oR, "streaming", fast_sparse=True)
        for a, b in zip(sk.Psi_cores, Psi):
            assert rel_err(a, b) < TOL, nnz
        for a, b in zip(sk.Omega_mats, Om):
            assert rel_err(a, b) < TOL, nnz


def test_sparse_sign_drm_bit_exact_and_sketch_vs_golden():
    """SparseSignDRM: ttsk_lazy_sparse_sign == the reference's inds_to_sparse_sign entry for entry
    (integers: exact), and stream_sketch under SparseSignDRMs == the reference's Psi / Omega."""
    from tt_sketch import _backend as be
    from tt_sketch.drm import SparseSignDRM
    from tt_sketch.sketch import stream_sketch

    z = load("sparse_sign.npz")
    for n, c in enumerate(z["cases"]):
        d = int(c[0]); shape = tuple(int(x) for x in c[1:1 + d]); k, rank, rmin, rmax, nzr, seed = (int(x) for x in c[5:])
        idx = z[f"c{n}_idx"]
        got = be.to_host(be.lazy_sparse_sign(be.to_device(idx, np.int64), k, idx.shape[1], shape, rank, rmin, rmax, nzr, seed))
        assert np.array_equal(got, z[f"c{n}_out"].astype(np.float64)), n
    t = tensor_desc(z, "sk_T")
    X = make_tensor(t)
    lr, rr = tuple(int(x) for x in z["lrank"]), tuple(int(x) for x in z["rrank"])
    left = SparseSignDRM(lr, shape=t[1], transpose=False, seed=11)
    right = SparseSignDRM(rr, shape=t[1], transpose=True, seed=23, num_non_zero_per_row=tuple(int(x) for x in z["sk_right_nnz"]))
    stt = stream_sketch(X, lr, rr, left_drm=left, right_drm=right)
    for i, a in enumerate(stt.Psi_cores):
        assert rel_err(a, z[f"sk_Psi{i}"]) < TOL
    for i, a in enumerate(stt.Omega_mats):
        assert rel_err(a, z[f"sk_Omega{i}"]) < TOL
    # slices (blocked sketches) pick columns of the same matrix
    sl = left.slice((1, 1, 2), (3, 4, 5))
    full = [be.to_host(m) for m in left.sketch_sparse_device(X)]
    part = [be.to_host(m) for m in sl.sketch_sparse_device(X)]
    for mu, (lo, hi) in enumerate(zip((1, 1, 2), (3, 4, 5))):
        assert np.array_equal(part[mu], full[mu][lo:hi])


@pytest.mark.parametrize("shape,lr,rr", [((300, 200, 50), (6, 8), (10, 4)),
                                         ((40, 30, 20, 10, 8), (4, 6, 8, 2), (8, 12, 6, 4)),
                                         ((5000, 3000, 700, 9), (20, 20, 20), (40, 40, 40))])
def test_gather_pass_on_payload_partition_vs_oracle(oracle_lib, shape, lr, rr):
    """Modes whose two factors are both prefix tables take the payload partition + bulk-copy gather pass
    (csrc/ttsk_sparse_gather.cu): orders 3, 4, 5, one very long segment, duplicates -- against the oracle."""
    from oracle.sketch_oracle import Drm
    from tt_sketch import _backend as be
    from tt_sketch.sketch_dispatch import SketchMethod, general_sketch

    nnz = 150_000
    _, idx, val = _c4_like(nnz, seed=len(shape), shape=shape)
    idx[1, :40_000] = 7          # one long segment in the gather mode
    idx[:, 1000:1100] = idx[:, :100]  # duplicates
    d = len(shape)
    oL = Drm("gauss", False, shape, (0,) * (d - 1), lr, 41)
    oR = Drm("gauss", True, shape, (0,) * (d - 1), rr, 42)
    desc = ("sparse", shape, idx, val)
    Psi, Om = oracle_lib.general_sketch(desc, oL, oR, "streaming", fast_sparse=True)
    before = be.lib().ttsk_sg_pass_count(be.ctx())
    sk = general_sketch(make_tensor(desc), make_drm(oL), make_drm(oR), SketchMethod.streaming)
    assert be.lib().ttsk_sg_pass_count(be.ctx()) > before
    for a, b in zip(sk.Psi_cores, Psi):
        assert rel_err(a, b) < TOL
    for a, b in zip(sk.Omega_mats, Om):
        assert rel_err(a, b) < TOL
