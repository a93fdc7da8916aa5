"""Pins the CPU oracle (oracle/) against the golden vectors the UNMODIFIED reference
produced (tests/golden/, see make_golden.py) and, when available, against the compiled
reference module in oracle/_ref and SciPy's ndtri.  CPU only."""
import numpy as np
import pytest

from _golden import drm_desc, load, rel_err, stored_list, tensor_desc

TOL = 1e-12  # oracle vs reference: same algorithm, same BLAS; only einsum paths may differ


def test_lazy_gaussian_bit_exact_vs_golden(oracle_lib):
    z = load("lazy_gaussian.npz")
    cases = z["cases"]
    for n, row in enumerate(cases):
        d = int(row[0]); shape = tuple(int(x) for x in row[1:1 + d]); k, rmin, rmax, seed = (int(x) for x in row[5:9])
        idx, want = z[f"c{n}_idx"], z[f"c{n}_out"]
        for restated in (False, True):
            got = oracle_lib.inds_to_normal(idx, shape[:k], rmin, rmax, seed, restated=restated)
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (n, restated)


def test_ndtri_restated_matches_scipy_including_tails(oracle_lib):
    import scipy.special
    rng = np.random.default_rng(0)
    u = rng.random(200000)
    samples = np.concatenate([u, u * 1e-6, u * 1e-12, 1 - u * 1e-6, u * 2.0 ** -40,
                              np.array([2.0 ** -52, 1 - 2.0 ** -52, 0.5, 0.13533528323661269189,
                                        1 - 0.13533528323661269189, np.nextafter(0.13533528323661269189, 1)])])
    want = scipy.special.ndtri(samples)
    for restated in (False, True):
        got = oracle_lib.ndtri(samples, restated=restated)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), restated


def test_log_restated_matches_libm(oracle_lib):
    import ctypes
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.random(100000) * 0.1353 + 2.0 ** -52, 2.0 + rng.random(100000) * 6.5,
                        2.0 ** -rng.integers(3, 52, 50000) * (1 + rng.random(50000))])
    a, b = np.empty_like(x), np.empty_like(x)
    oracle_lib._lib().ora_log_array(x.ctypes.data, a.ctypes.data, b.ctypes.data, x.size)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    assert np.array_equal(a.view(np.uint64), np.log(x).view(np.uint64)) or np.allclose(a, np.log(x), rtol=1e-15)


def test_oracle_vs_compiled_reference_module(oracle_lib):
    from oracle.ref_import import load_ref_extension
    try:
        ext = load_ref_extension()
    except ImportError:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(2)
    for shape in [(10, 12, 14, 7), (70000, 70000, 3), (10000, 10000, 10000, 500)]:
        k = len(shape)
        idx = np.stack([rng.integers(0, n, 20000) for n in shape]).astype(np.int64)
        want = np.asarray(ext.inds_to_normal(idx, shape[:k], 3, 29, 77))
        got = oracle_lib.inds_to_normal(idx, shape[:k], 3, 29, 77, restated=True)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_ttdrm_core_generation(oracle_lib):
    z = load("ttdrm_cores.npz")
    threads = int(z["cpu_count"])
    for side, right in (("l", False), ("r", True)):
        shape = tuple(int(x) for x in z[side + "_shape"]); rank = tuple(int(x) for x in z[side + "_rank"])
        cores = oracle_lib.tt_drm_cores(shape, rank, int(z[side + "_seed"]), right, threads=threads)
        for i, c in enumerate(cores):
            assert np.array_equal(c, z[f"{side}_core{i}"])


SKETCH = load("sketches.npz")
NAMES = [str(n) for n in SKETCH["names"]]


@pytest.mark.parametrize("name", NAMES)
def test_sketch_oracle_vs_golden(oracle_lib, name):
    z = SKETCH
    t = tensor_desc(z, name + "_T")
    shape = oracle_lib.tshape(t)
    L, R = drm_desc(z, name, "L", shape), drm_desc(z, name, "R", shape)
    methods = str(z[name + "_methods"]).split(",")
    if "stream" in methods:
        Psi, Om = oracle_lib.general_sketch(t, L, R, "streaming")
        for a, b in zip(Psi, stored_list(z, name + "_stream_Psi")):
            assert rel_err(a, b) < TOL
        for a, b in zip(Om, stored_list(z, name + "_stream_Omega")):
            assert rel_err(a, b) < TOL
        # scatter formulation == mask loop
        Psi2, _ = oracle_lib.general_sketch(t, L, R, "streaming", fast_sparse=True)
        for a, b in zip(Psi2, Psi):
            assert rel_err(a, b) < TOL
        C = oracle_lib.assemble(Psi, Om)
        want = stored_list(z, name + "_stream_C")
        assert oracle_lib.tt_error(C, want) < 1e-9
        if t[0] != "sum":
            for side, drm in (("L", L), ("R", R)):
                got = oracle_lib.drm_contractions(drm, t)
                for a, b in zip(got, stored_list(z, f"{name}_{side}c")):
                    if drm.kind == "gauss":
                        assert np.array_equal(np.ascontiguousarray(a).view(np.uint64), b.view(np.uint64))
                    else:
                        assert rel_err(a, b) < TOL
    if "orth" in methods:
        Psi, _ = oracle_lib.general_sketch(t, L, R, "orthogonal")
        assert oracle_lib.tt_error(Psi, stored_list(z, name + "_orth_C")) < 1e-9
    if "hmt" in methods:
        Psi, _ = oracle_lib.general_sketch(t, None, R, "hmt")
        assert oracle_lib.tt_error(Psi, stored_list(z, name + "_hmt_C")) < 1e-9
    if "blocked" in methods:
        ls = [tuple(int(x) for x in r) for r in z[name + "_lslices"]]
        rs = [tuple(int(x) for x in r) for r in z[name + "_rslices"]]
        Psi, Om = oracle_lib.blocked_sketch(t, L, R, ls, rs)
        for a, b in zip(Psi, stored_list(z, name + "_blocked_Psi")):
            assert rel_err(a, b) < TOL
        for a, b in zip(Om, stored_list(z, name + "_blocked_Omega")):
            assert rel_err(a, b) < TOL


def test_oracle_matches_reference_stt_ops_golden():
    """The oracle reproduces the reference's SketchedTensorTrain `+` and `increase_rank` outputs
    (tests/golden/stt_ops.npz): a sketch update is the sum of sketches, a rank increase is the blocked sketch
    over the slices [0, old, new] (reference sketch.py:292-349)."""
    from _golden import load, rel_err, stored_list, tensor_desc
    from oracle import sketch_oracle as orc

    z = load("stt_ops.npz")
    shape = (7, 8, 9, 10)
    lrank = tuple(int(x) for x in z["lrank"]); rrank = tuple(int(x) for x in z["rrank"])
    A, B = tensor_desc(z, "a_T"), tensor_desc(z, "b_T")
    oL = orc.Drm("gauss", False, shape, (0,) * 3, lrank, 11)
    oR = orc.Drm("gauss", True, shape, (0,) * 3, rrank, 23)
    Pa, Oa = orc.general_sketch(A, oL, oR, "streaming")
    Pb, Ob = orc.general_sketch(B, oL, oR, "streaming")
    for a, b, w in zip(Pa, Pb, stored_list(z, "gauss_add_Psi")):
        assert rel_err(a + b, w) < 1e-12
    for a, b, w in zip(Oa, Ob, stored_list(z, "gauss_add_Omega")):
        assert rel_err(a + b, w) < 1e-12
    new_l = tuple(int(x) for x in z["gauss_inc_lrank"]); new_r = tuple(int(x) for x in z["gauss_inc_rrank"])
    oL2 = orc.Drm("gauss", False, shape, (0,) * 3, new_l, 11)
    oR2 = orc.Drm("gauss", True, shape, (0,) * 3, new_r, 23)
    P, O = orc.blocked_sketch(A, oL2, oR2, [(0,) * 3, lrank, new_l], [(0,) * 3, rrank, new_r])
    for a, w in zip(P, stored_list(z, "gauss_inc_Psi")):
        assert rel_err(a, w) < 1e-12
    for a, w in zip(O, stored_list(z, "gauss_inc_Omega")):
        assert rel_err(a, w) < 1e-12


def test_sparse_sign_oracle_matches_reference_goldens(oracle_lib):
    """inds_to_sparse_sign restatement == the reference's int16 output (fast_lazy_gaussian.pyx:121-180) on 110 fixed
    cases: rank slices, fewer non-zeros than columns, the int32 stride wrap; and a whole stream_sketch under
    SparseSignDRMs (sparse_sign_drm.py:34-51)."""
    from _golden import load, rel_err, tensor_desc
    from oracle import sketch_oracle as orc

    z = load("sparse_sign.npz")
    for n, c in enumerate(z["cases"]):
        d = int(c[0]); shape = tuple(int(x) for x in c[1:1 + d]); k, rank, rmin, rmax, nzr, seed = (int(x) for x in c[5:])
        got = orc.inds_to_sparse_sign(z[f"c{n}_idx"], shape[:k], rank, rmin, rmax, nzr, seed)
        assert np.array_equal(got, z[f"c{n}_out"]), n
    t = tensor_desc(z, "sk_T")
    lr, rr = tuple(int(x) for x in z["lrank"]), tuple(int(x) for x in z["rrank"])
    L = orc.Drm("sign", False, t[1], (0,) * 3, lr, 11)
    R = orc.Drm("sign", True, t[1], (0,) * 3, rr, 23, nnz_row=tuple(int(x) for x in z["sk_right_nnz"])[::-1])
    Psi, Om = orc.general_sketch(t, L, R, "streaming")
    for i, a in enumerate(Psi):
        assert rel_err(a, z[f"sk_Psi{i}"]) < 1e-12
    for i, a in enumerate(Om):
        assert rel_err(a, z[f"sk_Omega{i}"]) < 1e-12
