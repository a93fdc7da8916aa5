#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the
authoring container.  Run from the repo root:

    oracle/build_ref.sh && python tests/golden/make_golden.py

The reference cannot travel to the GPU box, the fixtures can.  Every fixture stores the
INPUTS (tensor data, DRM seeds, TT-DRM cores -- the cores depend on the host's cpu_count(),
SURVEY.md App. B-3, so they are data, not something to regenerate) and the reference's
OUTPUTS (DRM entries, Psi/Omega, assembled TT cores).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402

ref = import_reference()
from tt_sketch.drm import SparseGaussianDRM, TensorTrainDRM  # noqa: E402
from tt_sketch.drm.fast_lazy_gaussian import inds_to_normal  # noqa: E402
from tt_sketch.sketch import (blocked_stream_sketch, hmt_sketch, orthogonal_sketch,  # noqa: E402
                              stream_sketch)
from tt_sketch.tensor import CPTensor, DenseTensor, SparseTensor, TensorSum, TensorTrain  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def gen_lazy_gaussian():
    rng = np.random.default_rng(12345)
    out = {}
    cases = []
    shapes = [(10, 12, 14, 7), (70000, 70000, 3), (10000, 10000, 10000, 500)]
    n = 0
    for shape in shapes:
        for k in range(1, len(shape) + 1):
            for (rmin, rmax) in [(0, 17), (5, 17), (12, 17)]:
                for seed in (5, 179):
                    nnz = 48
                    idx = np.stack([rng.integers(0, m, nnz) for m in shape[:k]]).astype(np.int64)
                    g = np.asarray(inds_to_normal(idx, shape[:k], rmin, rmax, seed))
                    out[f"c{n}_idx"] = idx
                    out[f"c{n}_out"] = g
                    cases.append((len(shape),) + tuple(shape) + (0,) * (4 - len(shape)) + (k, rmin, rmax, seed))
                    n += 1
    # deep tails / edge uniforms are covered by test_oracle_pin via scipy.special.ndtri
    out["cases"] = np.array(cases, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "lazy_gaussian.npz"), **out)
    print("lazy_gaussian.npz:", n, "cases")


def drm_pack(prefix, drm, store):
    store[prefix + "_seed"] = np.int64(int(drm.seed))
    store[prefix + "_rank_min"] = np.array(drm.rank_min, dtype=np.int64)  # internal orientation
    store[prefix + "_rank_max"] = np.array(drm.rank_max, dtype=np.int64)
    store[prefix + "_true_rank"] = np.array(drm.true_rank, dtype=np.int64)
    if hasattr(drm, "cores"):
        for i, c in enumerate(drm.cores):
            store[prefix + f"_core{i}"] = c


def sketch_pack(prefix, Psi, Omega, store):
    for i, p in enumerate(Psi):
        store[prefix + f"_Psi{i}"] = p
    for i, o in enumerate(Omega):
        store[prefix + f"_Omega{i}"] = o


def tensor_pack(prefix, t, store):
    if isinstance(t, SparseTensor):
        store[prefix + "_kind"] = np.array("sparse")
        store[prefix + "_shape"] = np.array(t.shape, dtype=np.int64)
        store[prefix + "_indices"] = np.asarray(t.indices, dtype=np.int64)
        store[prefix + "_entries"] = t.entries
    elif isinstance(t, DenseTensor):
        store[prefix + "_kind"] = np.array("dense")
        store[prefix + "_data"] = t.data
    elif isinstance(t, TensorTrain):
        store[prefix + "_kind"] = np.array("tt")
        for i, c in enumerate(t.cores):
            store[prefix + f"_c{i}"] = c
    elif isinstance(t, CPTensor):
        store[prefix + "_kind"] = np.array("cp")
        for i, c in enumerate(t.cores):
            store[prefix + f"_c{i}"] = c
    elif isinstance(t, TensorSum):
        store[prefix + "_kind"] = np.array("sum")
        store[prefix + "_n"] = np.int64(len(t.tensors))
        for i, s in enumerate(t.tensors):
            tensor_pack(prefix + f"_s{i}", s, store)


def make_sparse(shape, nnz, seed):
    rng = np.random.default_rng(seed)
    idx = np.stack([rng.integers(0, n, nnz) for n in shape]).astype(np.int64)
    return SparseTensor(shape, idx, rng.standard_normal(nnz))


def gen_sketches():
    store = {}
    names = []

    def run(name, tensor, lrank, rrank, ldrm_t, rdrm_t, methods=("stream",), lslices=None, rslices=None):
        shape = tensor.shape
        left = ldrm_t(lrank, shape=shape, transpose=False, seed=11)
        right = rdrm_t(rrank, shape=shape, transpose=True, seed=23)
        tensor_pack(name + "_T", tensor, store)
        drm_pack(name + "_L", left, store)
        drm_pack(name + "_R", right, store)
        store[name + "_Lkind"] = np.array(ldrm_t.__name__)
        store[name + "_Rkind"] = np.array(rdrm_t.__name__)
        store[name + "_lrank"] = np.array(lrank, dtype=np.int64)
        store[name + "_rrank"] = np.array(rrank, dtype=np.int64)
        store[name + "_methods"] = np.array(",".join(methods))
        if "stream" in methods:
            stt = stream_sketch(tensor, lrank, rrank, left_drm=left, right_drm=right)
            sketch_pack(name + "_stream", stt.Psi_cores, stt.Omega_mats, store)
            for i, c in enumerate(stt.C_cores()):
                store[name + f"_stream_C{i}"] = c
            # per-bond DRM contractions (operator-level parity for non-sum tensors)
            if not isinstance(tensor, TensorSum):
                from tt_sketch.sketch_dispatch import get_sketch_method
                for i, m in enumerate(get_sketch_method(tensor, left)(tensor)):
                    store[name + f"_Lc{i}"] = np.ascontiguousarray(m)
                for i, m in enumerate(get_sketch_method(tensor, right)(tensor)):
                    store[name + f"_Rc{i}"] = np.ascontiguousarray(m)
        if "orth" in methods:
            tt = orthogonal_sketch(tensor, lrank, rrank, left_drm=left, right_drm=right)
            for i, c in enumerate(tt.cores):
                store[name + f"_orth_C{i}"] = c
        if "hmt" in methods:
            tt = hmt_sketch(tensor, rrank, drm=right)
            for i, c in enumerate(tt.cores):
                store[name + f"_hmt_C{i}"] = c
        if "blocked" in methods:
            sk = blocked_stream_sketch(tensor, left, right, lslices, rslices)
            sketch_pack(name + "_blocked", sk.Psi_cores, sk.Omega_mats, store)
            store[name + "_lslices"] = np.array(lslices, dtype=np.int64)
            store[name + "_rslices"] = np.array(rslices, dtype=np.int64)
        names.append(name)

    sp = make_sparse((7, 8, 9, 10), 300, 1)
    run("sparse_gauss", sp, (3, 4, 5), (5, 6, 7), SparseGaussianDRM, SparseGaussianDRM,
        methods=("stream", "blocked"),
        lslices=[(0, 0, 0), (2, 2, 3), (3, 4, 5)], rslices=[(0, 0, 0), (3, 3, 3), (5, 6, 7)])
    run("sparse_ttdrm", sp, (3, 4, 5), (5, 6, 7), TensorTrainDRM, TensorTrainDRM,
        methods=("stream", "orth", "hmt", "blocked"),
        lslices=[(0, 0, 0), (2, 2, 3), (3, 4, 5)], rslices=[(0, 0, 0), (3, 3, 3), (5, 6, 7)])
    run("sparse_mixed", sp, (6, 7, 8), (3, 4, 5), TensorTrainDRM, SparseGaussianDRM, methods=("stream",))
    sp2 = make_sparse((9, 10), 40, 2)
    run("sparse_d2", sp2, (3,), (5,), SparseGaussianDRM, SparseGaussianDRM, methods=("stream",))
    sp3 = make_sparse((9, 10, 11), 120, 3)
    run("sparse_d3", sp3, (3, 4), (5, 6), SparseGaussianDRM, TensorTrainDRM, methods=("stream", "orth"))
    sp5 = make_sparse((5, 6, 7, 8, 4), 400, 4)
    run("sparse_d5", sp5, (3, 4, 5, 4), (5, 6, 7, 5), SparseGaussianDRM, SparseGaussianDRM, methods=("stream",))

    rng = np.random.default_rng(5)
    dn = DenseTensor(rng.standard_normal((5, 6, 7, 8)))
    run("dense", dn, (3, 4, 5), (5, 6, 7), TensorTrainDRM, TensorTrainDRM, methods=("stream", "orth", "hmt"))
    dn2 = DenseTensor(rng.standard_normal((6, 6, 6)))
    run("dense_lbig", dn2, (5, 6), (3, 4), TensorTrainDRM, TensorTrainDRM, methods=("stream",))

    tt = TensorTrain.random((7, 8, 9, 10), (4, 5, 3), seed=6)
    run("tt", tt, (3, 4, 5), (5, 6, 7), TensorTrainDRM, TensorTrainDRM,
        methods=("stream", "orth", "hmt", "blocked"),
        lslices=[(0, 0, 0), (1, 2, 2), (3, 4, 5)], rslices=[(0, 0, 0), (2, 3, 4), (5, 6, 7)])
    cp = CPTensor.random((7, 8, 9, 10), 6, seed=7)
    run("cp", cp, (3, 4, 5), (5, 6, 7), TensorTrainDRM, TensorTrainDRM, methods=("stream", "orth", "hmt"))

    tsum = TensorSum([tt, sp, cp, 0.5 * TensorTrain.random((7, 8, 9, 10), 2, seed=8)])
    run("sum", tsum, (3, 4, 5), (5, 6, 7), TensorTrainDRM, TensorTrainDRM,
        methods=("stream", "orth", "blocked"),
        lslices=[(0, 0, 0), (2, 2, 2), (3, 4, 5)], rslices=[(0, 0, 0), (4, 4, 4), (5, 6, 7)])
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "sketches.npz"), **store)
    print("sketches.npz:", names)


def gen_ttdrm_cores():
    """TT-DRM core generation (tensor_train_drm.py:46-56) with the thread count pinned so
    the fixture is reproducible anywhere: values depend on cpu_count() (App. B-3)."""
    import multiprocessing
    store = {"cpu_count": np.int64(multiprocessing.cpu_count())}
    for name, shape, rank, transpose, seed in [("l", (7, 8, 9, 10), (3, 4, 5), False, 11),
                                               ("r", (7, 8, 9, 10), (5, 6, 7), True, 23)]:
        drm = TensorTrainDRM(rank, shape=shape, transpose=transpose, seed=seed)
        store[name + "_shape"] = np.array(shape)
        store[name + "_rank"] = np.array(rank)
        store[name + "_seed"] = np.int64(seed)
        for i, c in enumerate(drm.cores):
            store[name + f"_core{i}"] = c
    np.savez_compressed(os.path.join(OUT, "ttdrm_cores.npz"), **store)
    print("ttdrm_cores.npz")


def gen_stt_ops():
    """SketchedTensorTrain streaming operations (sketch.py:272-361 of the reference): `+` (sketch update with
    the stored DRMs), `increase_rank` (block (0, 0) reused), `.T`, `to_tt`."""
    store = {}
    shape = (7, 8, 9, 10)
    sp = make_sparse(shape, 300, 1)
    sp_b = make_sparse(shape, 150, 9)
    tt_b = TensorTrain.random(shape, (2, 3, 2), seed=10)
    lrank, rrank = (3, 4, 5), (5, 6, 7)
    tensor_pack("a_T", sp, store)
    tensor_pack("b_T", sp_b, store)
    tensor_pack("c_T", tt_b, store)
    # ---- SparseGaussianDRM: add, increase_rank, transpose
    left = SparseGaussianDRM(lrank, shape=shape, transpose=False, seed=11)
    right = SparseGaussianDRM(rrank, shape=shape, transpose=True, seed=23)
    stt = stream_sketch(sp, lrank, rrank, left_drm=left, right_drm=right)
    added = stt + sp_b
    sketch_pack("gauss_add", added.Psi_cores, added.Omega_mats, store)
    new_l, new_r = (4, 6, 6), (6, 8, 9)
    inc = stt.increase_rank(sp, new_l, new_r)
    sketch_pack("gauss_inc", inc.Psi_cores, inc.Omega_mats, store)
    store["gauss_inc_lrank"] = np.array(new_l, dtype=np.int64)
    store["gauss_inc_rrank"] = np.array(new_r, dtype=np.int64)
    tr = stt.T
    sketch_pack("gauss_T", tr.Psi_cores, tr.Omega_mats, store)
    for i, c in enumerate(tr.C_cores()):
        store[f"gauss_T_C{i}"] = c
    for i, c in enumerate(inc.to_tt().cores):
        store[f"gauss_inc_C{i}"] = c
    # ---- TensorTrainDRM: add a TensorTrain to the sketch of a sparse tensor
    left = TensorTrainDRM(lrank, shape=shape, transpose=False, seed=11)
    right = TensorTrainDRM(rrank, shape=shape, transpose=True, seed=23)
    drm_pack("tt_L", left, store)
    drm_pack("tt_R", right, store)
    stt = stream_sketch(sp, lrank, rrank, left_drm=left, right_drm=right)
    added = stt + tt_b
    sketch_pack("tt_add", added.Psi_cores, added.Omega_mats, store)
    for i, c in enumerate(added.to_tt().cores):
        store[f"tt_add_C{i}"] = c
    store["lrank"] = np.array(lrank, dtype=np.int64)
    store["rrank"] = np.array(rrank, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "stt_ops.npz"), **store)
    print("stt_ops.npz")


def gen_sparse_sign():
    """SparseSignDRM (sparse_sign_drm.py:12-51, fast_lazy_gaussian.pyx:121-180 of the reference): raw
    inds_to_sparse_sign outputs (int16: 0 / -1 / +1) incl. rank slices, fewer non-zeros than columns and the int32
    stride wrap, and one stream_sketch of a sparse tensor under SparseSignDRMs."""
    from tt_sketch.drm import SparseSignDRM
    from tt_sketch.drm.fast_lazy_gaussian import inds_to_sparse_sign

    rng = np.random.default_rng(777)
    store, cases, n = {}, [], 0
    for shape in [(10, 12, 14, 7), (70000, 70000, 3), (10000, 10000, 10000, 500)]:
        for k in range(1, len(shape) + 1):
            for (rank, rmin, rmax, nzr) in [(17, 0, 17, 17), (17, 5, 17, 17), (17, 0, 9, 6), (40, 12, 33, 40), (8, 0, 8, 1)]:
                for seed in (5, 179):
                    nnz = 40
                    idx = np.stack([rng.integers(0, m, nnz) for m in shape[:k]]).astype(np.int64)
                    g = np.asarray(inds_to_sparse_sign(idx, shape[:k], rank, rmin, rmax, nzr, seed))
                    store[f"c{n}_idx"] = idx
                    store[f"c{n}_out"] = g.astype(np.int16)
                    cases.append((len(shape),) + tuple(shape) + (0,) * (4 - len(shape)) + (k, rank, rmin, rmax, nzr, seed))
                    n += 1
    store["cases"] = np.array(cases, dtype=np.int64)
    shape = (7, 8, 9, 10)
    sp = make_sparse(shape, 300, 1)
    lrank, rrank = (3, 4, 5), (5, 6, 7)
    tensor_pack("sk_T", sp, store)
    left = SparseSignDRM(lrank, shape=shape, transpose=False, seed=11)
    right = SparseSignDRM(rrank, shape=shape, transpose=True, seed=23, num_non_zero_per_row=(3, 4, 2))
    stt = stream_sketch(sp, lrank, rrank, left_drm=left, right_drm=right)
    sketch_pack("sk", stt.Psi_cores, stt.Omega_mats, store)
    store["sk_right_nnz"] = np.array((3, 4, 2), dtype=np.int64)
    store["lrank"] = np.array(lrank, dtype=np.int64)
    store["rrank"] = np.array(rrank, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "sparse_sign.npz"), **store)
    print("sparse_sign.npz:", n, "cases")


def gen_tucker_dense_gauss():
    """Section 8(f) rank 2: TuckerTensor input under TT-DRMs (tucker_sketch.py:9-46, tensor_train_drm.py:124-145 of
    the reference) and the DenseGaussianDRM (dense_gaussian_drm.py:17-80) on sparse / TT / dense inputs, incl. a
    blocked sketch and a rank increase.  Stored like sketches.npz plus the DenseGaussianDRM matrices themselves."""
    from tt_sketch.drm import DenseGaussianDRM
    from tt_sketch.sketch_dispatch import get_sketch_method
    from tt_sketch.tensor import TuckerTensor

    store, names = {}, []
    # ---- Tucker, d = 2, 3, 4 (the reference tests d = 2, 3 through exact recovery only)
    for name, shape, trank, lrank, rrank in [("tucker_d2", (10, 11), 3, (3,), (4,)),
                                             ("tucker_d3", (10, 11, 12), (3, 2, 4), (3, 4), (4, 5)),
                                             ("tucker_d4", (7, 8, 9, 10), (3, 4, 2, 3), (3, 4, 5), (5, 6, 7))]:
        X = TuckerTensor.random(shape, trank, seed=180)
        left = TensorTrainDRM(lrank, shape=shape, transpose=False, seed=11)
        right = TensorTrainDRM(rrank, shape=shape, transpose=True, seed=23)
        store[name + "_core"] = X.core
        for i, U in enumerate(X.factors):
            store[name + f"_U{i}"] = U
        drm_pack(name + "_L", left, store)
        drm_pack(name + "_R", right, store)
        store[name + "_lrank"] = np.array(lrank, dtype=np.int64)
        store[name + "_rrank"] = np.array(rrank, dtype=np.int64)
        stt = stream_sketch(X, lrank, rrank, left_drm=left, right_drm=right)
        sketch_pack(name + "_stream", stt.Psi_cores, stt.Omega_mats, store)
        for i, c in enumerate(stt.C_cores()):
            store[name + f"_stream_C{i}"] = c
        for i, m in enumerate(get_sketch_method(X, left)(X)):
            store[name + f"_Lc{i}"] = np.ascontiguousarray(m)
        for i, m in enumerate(get_sketch_method(X, right)(X)):
            store[name + f"_Rc{i}"] = np.ascontiguousarray(m)
        for i, c in enumerate(orthogonal_sketch(X, lrank, rrank, left_drm=left, right_drm=right).cores):
            store[name + f"_orth_C{i}"] = c
        for i, c in enumerate(hmt_sketch(X, rrank, drm=right).cores):
            store[name + f"_hmt_C{i}"] = c
        store[name + "_dense"] = X.to_numpy()
        names.append(name)
    store["tucker_names"] = np.array(names)
    # ---- DenseGaussianDRM
    shape = (5, 6, 7, 4)
    lrank, rrank = (3, 4, 3), (4, 6, 4)
    sp = make_sparse(shape, 200, 12)
    tt = TensorTrain.random(shape, (3, 4, 2), seed=13)
    dn = DenseTensor(np.random.default_rng(14).standard_normal(shape))
    tensor_pack("dg_sparse_T", sp, store)
    tensor_pack("dg_tt_T", tt, store)
    tensor_pack("dg_dense_T", dn, store)
    left = DenseGaussianDRM(lrank, shape=shape, transpose=False, seed=11)
    right = DenseGaussianDRM(rrank, shape=shape, transpose=True, seed=23)
    for side, drm in (("L", left), ("R", right)):
        for i, m in enumerate(drm.sketching_mats):
            store[f"dg_{side}_mat{i}"] = m
    sl = DenseGaussianDRM(lrank, shape=shape, transpose=False, seed=11).slice((1, 1, 0), (3, 3, 2))
    for i, m in enumerate(sl.sketching_mats):
        store[f"dg_Lslice_mat{i}"] = m
    inc = left.increase_rank((4, 5, 4))
    for i, m in enumerate(inc.sketching_mats):
        store[f"dg_Linc_mat{i}"] = m
    for key, X in (("dg_sparse", sp), ("dg_tt", tt), ("dg_dense", dn)):
        stt = stream_sketch(X, lrank, rrank, left_drm=left, right_drm=right)
        sketch_pack(key + "_stream", stt.Psi_cores, stt.Omega_mats, store)
        for i, m in enumerate(get_sketch_method(X, left)(X)):
            store[key + f"_Lc{i}"] = np.ascontiguousarray(m)
        for i, m in enumerate(get_sketch_method(X, right)(X)):
            store[key + f"_Rc{i}"] = np.ascontiguousarray(m)
        for i, c in enumerate(orthogonal_sketch(X, lrank, rrank, left_drm=left, right_drm=right).cores):
            store[key + f"_orth_C{i}"] = c
    lsl, rsl = [(0, 0, 0), (2, 2, 1), (3, 4, 3)], [(0, 0, 0), (2, 3, 2), (4, 6, 4)]
    sk = blocked_stream_sketch(sp, left, right, lsl, rsl)
    sketch_pack("dg_sparse_blocked", sk.Psi_cores, sk.Omega_mats, store)
    store["dg_lslices"] = np.array(lsl, dtype=np.int64)
    store["dg_rslices"] = np.array(rsl, dtype=np.int64)
    store["dg_lrank"] = np.array(lrank, dtype=np.int64)
    store["dg_rrank"] = np.array(rrank, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "tucker_dense_gauss.npz"), **store)
    print("tucker_dense_gauss.npz:", names)


def gen_tt_algebra():
    """Section 8(f) rank 3: TensorTrain.orthogonalize / round / svdvals / dot / norm / error / gather of the reference
    (tensor.py:414-609) on a TT whose ranks are inflated by a direct sum (so rounding has something to cut)."""
    store = {}
    shape = (7, 8, 9, 10)
    a = TensorTrain.random(shape, (4, 5, 3), seed=21)
    b = TensorTrain.random(shape, (2, 3, 2), seed=22)
    s = a.add(b * 1e-4)                       # ranks (6, 8, 5): the small summand is what eps = 1e-3 removes
    tensor_pack("a_T", a, store)
    tensor_pack("b_T", b, store)
    tensor_pack("s_T", s, store)
    for i, c in enumerate(s.orthogonalize().cores):
        store[f"orth_C{i}"] = c
    for key, kw in (("round_eps", dict(eps=1e-3)), ("round_rank", dict(max_rank=(3, 4, 2))), ("round_exact", dict(eps=1e-12))):
        r = s.round(**kw)
        store[key + "_rank"] = np.array(r.rank, dtype=np.int64)
        store[key + "_dense"] = r.to_numpy()
    for i, v in enumerate(s.svdvals()):
        store[f"svdvals{i}"] = v
    store["dot_ab"] = np.float64(a.dot(b))
    store["norm_s"] = np.float64(s.norm())
    store["err_sa"] = np.float64(s.error(a))
    store["err_sa_rel"] = np.float64(s.error(a, relative=True))
    rng = np.random.default_rng(23)
    idx = np.stack([rng.integers(0, n, 64) for n in shape]).astype(np.int64)
    store["gather_idx"] = idx
    store["gather_s"] = s.gather(idx)
    np.savez_compressed(os.path.join(OUT, "tt_algebra.npz"), **store)
    print("tt_algebra.npz")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "stt_ops":  # later additions leave the earlier fixtures untouched
        gen_stt_ops()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sparse_sign":
        gen_sparse_sign()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "tt_algebra":
        gen_tt_algebra()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "tucker_dense_gauss":
        gen_tucker_dense_gauss()
        sys.exit(0)
    gen_lazy_gaussian()
    gen_sketches()
    gen_ttdrm_cores()
    gen_stt_ops()
    gen_sparse_sign()
    gen_tucker_dense_gauss()
    gen_tt_algebra()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
