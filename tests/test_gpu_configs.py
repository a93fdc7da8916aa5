"""GPU parity at the sizes BASELINE.json names (C1, C2, C3 in full; C4 / C5 at their shape and ranks with as
many nonzeros as the CPU oracle sketches in seconds), against the CPU oracle on the same seeded inputs.

Tolerances: Psi / Omega relative max-norm <= 1e-10; orthogonalised cores (Householder QR with LAPACK's
sign convention, so cores are comparable element-wise) <= 1e-9.
"""
import numpy as np
import pytest

from _golden import rel_err
from _product import make_drm, make_tensor

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _odrm(orc, d):
    return orc.Drm("tt", d.transpose, d.shape, d.bond_rank_min, d.bond_rank_max, d.seed, list(d.cores))


def _check(sk, Psi, Om, tol=TOL):
    assert len(sk.Psi_cores) == len(Psi) and len(sk.Omega_mats) == len(Om)
    for a, b in zip(sk.Psi_cores, Psi):
        assert rel_err(a, b) < tol
    for a, b in zip(sk.Omega_mats, Om):
        assert rel_err(a, b) < tol


def test_c1_dense_20pow5_stream_sketch_full_size(oracle_lib):
    """BASELINE configs[0]: stream_sketch of a dense 20^5 tensor, TensorTrainDRM rL=10 / rR=15."""
    from tt_sketch.drm import TensorTrainDRM
    from tt_sketch.sketch import stream_sketch
    from tt_sketch.tensor import DenseTensor

    shape = (20,) * 5
    X = DenseTensor(np.random.default_rng(0).standard_normal(shape))
    lr, rr = (10,) * 4, (15,) * 4
    L = TensorTrainDRM(lr, shape=shape, transpose=False, seed=1)
    R = TensorTrainDRM(rr, shape=shape, transpose=True, seed=2)
    stt = stream_sketch(X, lr, rr, left_drm=L, right_drm=R)
    Psi, Om = oracle_lib.general_sketch(("dense", X.data), _odrm(oracle_lib, L), _odrm(oracle_lib, R), "streaming")
    _check(stt, Psi, Om)
    assert [p.shape for p in stt.Psi_cores] == [(1, 20, 15)] + [(10, 20, 15)] * 3 + [(10, 20, 1)]


def test_c2_tt_orthogonal_sketch_full_size(oracle_lib):
    """BASELINE configs[1]: orthogonal_sketch of an order-10 TT (dim 50, rank 100) to rank 20 (right rank 40);
    also the streaming sketch of the same tensor."""
    from tt_sketch.drm import TensorTrainDRM
    from tt_sketch.sketch import orthogonal_sketch, stream_sketch
    from tt_sketch.tensor import TensorTrain

    shape = (50,) * 10
    T = TensorTrain.random(shape, 100, seed=2)
    lr, rr = (20,) * 9, (40,) * 9
    L = TensorTrainDRM(lr, shape=shape, transpose=False, seed=1)
    R = TensorTrainDRM(rr, shape=shape, transpose=True, seed=2)
    tt = orthogonal_sketch(T, lr, rr, left_drm=L, right_drm=R)
    Psi, _ = oracle_lib.general_sketch(("tt", T.cores), _odrm(oracle_lib, L), _odrm(oracle_lib, R), "orthogonal")
    assert tt.rank == lr
    for a, b in zip(tt.cores, Psi):
        assert rel_err(a, b) < 1e-9
    stt = stream_sketch(T, lr, rr, left_drm=L, right_drm=R)
    Psi, Om = oracle_lib.general_sketch(("tt", T.cores), _odrm(oracle_lib, L), _odrm(oracle_lib, R), "streaming")
    _check(stt, Psi, Om)


def test_c3_cp_stream_sketch_full_size(oracle_lib):
    """BASELINE configs[2]: stream_sketch of an order-8 CP tensor (dim 100, CP rank 200) to TT rank 30 (right 60)."""
    from tt_sketch.drm import TensorTrainDRM
    from tt_sketch.sketch import stream_sketch
    from tt_sketch.tensor import CPTensor

    shape = (100,) * 8
    C = CPTensor.random(shape, 200, seed=3)
    lr, rr = (30,) * 7, (60,) * 7
    L = TensorTrainDRM(lr, shape=shape, transpose=False, seed=1)
    R = TensorTrainDRM(rr, shape=shape, transpose=True, seed=2)
    stt = stream_sketch(C, lr, rr, left_drm=L, right_drm=R)
    Psi, Om = oracle_lib.general_sketch(("cp", C.cores), _odrm(oracle_lib, L), _odrm(oracle_lib, R), "streaming")
    _check(stt, Psi, Om)


# ---------------------------------------------------------------- C4 / C5 shape
C4_SHAPE = (10000, 10000, 10000, 500)


def _c4_coo(nnz, seed=0):
    idx = np.stack([np.random.default_rng(100 + k + seed).integers(0, n, nnz) for k, n in enumerate(C4_SHAPE)]).astype(np.int64)
    val = np.random.default_rng(99 + seed).standard_normal(nnz)
    return idx, val


def _oracle_sparse_chunked(orc, idx, val, oL, oR, blocked=None, chunk=25000):
    """Oracle sketch of a sparse tensor as the sum over nonzero chunks (the sketch is linear in the
    nonzeros; the reference's TT-DRM path materialises an (r, nnz, r) array per level, 12.8 KB per nonzero)."""
    Psi = Om = None
    for c0 in range(0, idx.shape[1], chunk):
        desc = ("sparse", C4_SHAPE, idx[:, c0:c0 + chunk], val[c0:c0 + chunk])
        if blocked is None:
            P, O = orc.general_sketch(desc, oL, oR, "streaming", fast_sparse=True)
        else:
            P, O = orc.blocked_sketch(desc, oL, oR, blocked[0], blocked[1], fast_sparse=True)
        Psi = P if Psi is None else [a + b for a, b in zip(Psi, P)]
        Om = O if Om is None else [a + b for a, b in zip(Om, O)]
    return Psi, Om


def _c4_drms(orc, kinds):
    rl, rr = (20,) * 3, (40,) * 3
    cl = orc.tt_drm_cores(C4_SHAPE, rl, 1, False) if kinds[0] == "tt" else []
    cr = orc.tt_drm_cores(C4_SHAPE, rr, 2, True) if kinds[1] == "tt" else []
    return (orc.Drm(kinds[0], False, C4_SHAPE, (0,) * 3, rl, 1, cl), orc.Drm(kinds[1], True, C4_SHAPE, (0,) * 3, rr, 2, cr))


@pytest.mark.parametrize("kinds", [("tt", "tt"), ("tt", "gauss")])
def test_c5_sparse_term_ttdrm_bucketed_chain_vs_oracle(oracle_lib, kinds):
    """TT DRMs on a sparse tensor at the C4 / C5 shape with 2e5 nonzeros: the bucketed DMMA chain kernel with
    the two-level scatter (it is taken from 1.5e5 nonzeros on), against the oracle."""
    from tt_sketch.sketch import stream_sketch

    idx, val = _c4_coo(200_000, seed=3)
    oL, oR = _c4_drms(oracle_lib, kinds)
    Psi, Om = _oracle_sparse_chunked(oracle_lib, idx, val, oL, oR)
    stt = stream_sketch(make_tensor(("sparse", C4_SHAPE, idx, val)), (20,) * 3, (40,) * 3, left_drm=make_drm(oL),
                        right_drm=make_drm(oR))
    _check(stt, Psi, Om)


@pytest.mark.parametrize("kinds", [("gauss", "gauss"), ("tt", "tt")])
def test_host_streaming_in_several_chunks_vs_oracle(oracle_lib, kinds):
    """ttsk_sparse_sketch_host with the staging size forced down so the input is streamed in >= 3 chunks
    (graduated chunk lengths, every chunk flushes its own segments): same oracle, same tolerance."""
    from ctypes import byref

    import torch

    from tt_sketch import _backend as be
    from tt_sketch.sketch_container import SketchContainer
    from tt_sketch.sketch_dispatch import drm_descriptor

    nnz = 180_000
    idx, val = _c4_coo(nnz, seed=5)
    oL, oR = _c4_drms(oracle_lib, kinds)
    Psi, Om = _oracle_sparse_chunked(oracle_lib, idx, val, oL, oR)
    L, R = make_drm(oL), make_drm(oR)
    ld, lk = drm_descriptor(L)
    rd, rk = drm_descriptor(R)
    _, total = SketchContainer.layout(C4_SHAPE, (20,) * 3, (40,) * 3)
    h_idx = torch.from_numpy(idx).pin_memory()
    h_val = torch.from_numpy(val).pin_memory()
    h_out = torch.empty(total, dtype=torch.float64).pin_memory()
    lib, ctx = be.lib(), be.ctx()
    be.check(lib.ttsk_set_stage_nnz(ctx, 50_000))  # chunks of 6250, 18750, 50000, 50000, 50000, 5000
    try:
        be.check(lib.ttsk_sparse_sketch_host(ctx, 4, be.as_i64(C4_SHAPE), nnz, h_idx.data_ptr(), h_idx.stride(0),
                                             h_val.data_ptr(), byref(ld), byref(rd), h_out.data_ptr(), 0))
    finally:
        be.check(lib.ttsk_set_stage_nnz(ctx, 1 << 24))
    sk = SketchContainer.unpack(h_out.numpy(), C4_SHAPE, (20,) * 3, (40,) * 3)
    _check(sk, Psi, Om)
    del lk, rk


def test_c5_shaped_tensorsum_blocked_2x2_vs_oracle(oracle_lib):
    """BASELINE configs[4] scaled to what the oracle sketches in seconds: TensorSum(3 TT rank 10 + sparse
    6e4 nnz) at the C4 shape, TensorTrainDRMs 20 / 40, blocked_stream_sketch with 2 x 2 rank blocks, against
    the oracle's blocked sketch (NOT against the unblocked GPU result)."""
    from tt_sketch.sketch import blocked_stream_sketch
    from tt_sketch.tensor import SparseTensor, TensorSum, TensorTrain

    idx, val = _c4_coo(60_000, seed=9)
    tts = [TensorTrain.random(C4_SHAPE, 10, seed=1000 + k) for k in range(3)]
    oL, oR = _c4_drms(oracle_lib, ("tt", "tt"))
    ls, rs = [(0,) * 3, (10,) * 3, (20,) * 3], [(0,) * 3, (20,) * 3, (40,) * 3]
    Psi, Om = _oracle_sparse_chunked(oracle_lib, idx, val, oL, oR, blocked=(ls, rs))
    for t in tts:
        P, O = oracle_lib.blocked_sketch(("tt", t.cores), oL, oR, ls, rs)
        Psi = [a + b for a, b in zip(Psi, P)]
        Om = [a + b for a, b in zip(Om, O)]
    S = TensorSum(tts + [SparseTensor(C4_SHAPE, idx, val)])
    sk = blocked_stream_sketch(S, make_drm(oL), make_drm(oR), ls, rs)
    _check(sk, Psi, Om)


def test_table_cache_cap_never_evicts_tables_in_use(oracle_lib):
    """With the prefix-table cache capped below what one sketch needs, bonds whose table does not fit are
    generated on the fly and tables of the running call are never dropped: results stay equal to the oracle."""
    from tt_sketch import _backend as be
    from tt_sketch.sketch import stream_sketch

    shape = (300, 200, 150, 40)
    nnz = 400_000
    rng = np.random.default_rng(11)
    idx = np.stack([rng.integers(0, n, nnz) for n in shape]).astype(np.int64)
    val = rng.standard_normal(nnz)
    oL = oracle_lib.Drm("gauss", False, shape, (0,) * 3, (6, 7, 8), 21)
    oR = oracle_lib.Drm("gauss", True, shape, (0,) * 3, (9, 10, 11), 22)
    desc = ("sparse", shape, idx, val)
    Psi, Om = oracle_lib.general_sketch(desc, oL, oR, "streaming", fast_sparse=True)
    lib, ctx = be.lib(), be.ctx()
    be.check(lib.ttsk_trim(ctx))
    try:
        for cap in (1 << 20, 300_000, 0):  # L_1 (60000 x 7) and R_1 (6000 x 10) tables fit or not
            be.check(lib.ttsk_set_table_cache_cap(ctx, cap))
            for _ in range(2):  # second call: cache hits / re-generation
                stt = stream_sketch(make_tensor(desc), (6, 7, 8), (9, 10, 11), left_drm=make_drm(oL), right_drm=make_drm(oR))
                _check(stt, Psi, Om)
    finally:
        be.check(lib.ttsk_set_table_cache_cap(ctx, 6 << 30))
        be.check(lib.ttsk_trim(ctx))


def test_device_copy_follows_replaced_host_arrays(oracle_lib):
    """Replacing a container's arrays after a sketch must be seen by the next sketch (the device copy is keyed
    by the identity of the host arrays); `invalidate_device()` covers in-place edits."""
    from tt_sketch.sketch import stream_sketch
    from tt_sketch.tensor import SparseTensor, TensorTrain

    shape = (12, 9, 10)
    rng = np.random.default_rng(4)
    idx = np.stack([rng.integers(0, n, 500) for n in shape]).astype(np.int64)
    X = SparseTensor(shape, idx, rng.standard_normal(500))
    oL = oracle_lib.Drm("gauss", False, shape, (0, 0), (3, 4), 5)
    oR = oracle_lib.Drm("gauss", True, shape, (0, 0), (5, 6), 6)
    L, R = make_drm(oL), make_drm(oR)
    stream_sketch(X, (3, 4), (5, 6), left_drm=L, right_drm=R)
    X.entries = rng.standard_normal(500)                      # replaced array
    stt = stream_sketch(X, (3, 4), (5, 6), left_drm=L, right_drm=R)
    _check(stt, *oracle_lib.general_sketch(("sparse", shape, idx, X.entries), oL, oR, "streaming", fast_sparse=True))
    X.entries[:50] = 0.0                                      # in-place edit
    X.invalidate_device()
    stt = stream_sketch(X, (3, 4), (5, 6), left_drm=L, right_drm=R)
    _check(stt, *oracle_lib.general_sketch(("sparse", shape, idx, X.entries), oL, oR, "streaming", fast_sparse=True))
    T = TensorTrain.random(shape, 4, seed=8)
    cl = oracle_lib.tt_drm_cores(shape, (3, 4), 1, False)
    cr = oracle_lib.tt_drm_cores(shape, (5, 6), 2, True)
    oLt = oracle_lib.Drm("tt", False, shape, (0, 0), (3, 4), 1, cl)
    oRt = oracle_lib.Drm("tt", True, shape, (0, 0), (5, 6), 2, cr)
    Lt, Rt = make_drm(oLt), make_drm(oRt)
    stream_sketch(T, (3, 4), (5, 6), left_drm=Lt, right_drm=Rt)
    T[1] = rng.standard_normal(T[1].shape)                    # replaced core
    stt = stream_sketch(T, (3, 4), (5, 6), left_drm=Lt, right_drm=Rt)
    _check(stt, *oracle_lib.general_sketch(("tt", T.cores), oLt, oRt, "streaming"))


# ---------------------------------------------------------------- SketchedTensorTrain operations vs reference goldens
def test_sketched_tensor_train_ops_vs_reference_golden(oracle_lib):
    """`+` (streaming update), `increase_rank`, `.T`, `to_tt` against outputs written by the unmodified
    reference (tests/golden/stt_ops.npz, reference sketch.py:272-361)."""
    from _golden import load, stored_list, tensor_desc
    from oracle.sketch_oracle import Drm, to_dense
    from tt_sketch.sketch import stream_sketch

    z = load("stt_ops.npz")
    shape = (7, 8, 9, 10)
    lrank = tuple(int(x) for x in z["lrank"]); rrank = tuple(int(x) for x in z["rrank"])
    A, B, C = (make_tensor(tensor_desc(z, p)) for p in ("a_T", "b_T", "c_T"))

    def same(sk, prefix):
        for a, b in zip(sk.Psi_cores, stored_list(z, prefix + "_Psi")):
            assert rel_err(a, b) < TOL
        for a, b in zip(sk.Omega_mats, stored_list(z, prefix + "_Omega")):
            assert rel_err(a, b) < TOL

    def tt_close(cores, want, tol=1e-9):
        a, b = to_dense(("tt", list(cores))), to_dense(("tt", list(want)))
        assert np.linalg.norm(a - b) <= tol * np.linalg.norm(b)

    L = make_drm(Drm("gauss", False, shape, (0,) * 3, lrank, 11))
    R = make_drm(Drm("gauss", True, shape, (0,) * 3, rrank, 23))
    stt = stream_sketch(A, lrank, rrank, left_drm=L, right_drm=R)
    same(stt + B, "gauss_add")
    new_l = tuple(int(x) for x in z["gauss_inc_lrank"]); new_r = tuple(int(x) for x in z["gauss_inc_rrank"])
    inc = stt.increase_rank(A, new_l, new_r)
    same(inc, "gauss_inc")
    assert inc.left_rank == new_l and inc.right_rank == new_r
    tt_close(inc.to_tt().cores, stored_list(z, "gauss_inc_C"))
    tr = stt.T
    same(tr, "gauss_T")
    assert tr.shape == shape[::-1]
    tt_close(tr.C_cores(), stored_list(z, "gauss_T_C"))

    def tt_drm(side, right, rank):
        cores = stored_list(z, f"tt_{side}_core")
        return make_drm(Drm("tt", right, shape, (0,) * 3, rank, int(z[f"tt_{side}_seed"]), cores))

    Lt, Rt = tt_drm("L", False, lrank), tt_drm("R", True, rrank)
    stt = stream_sketch(A, lrank, rrank, left_drm=Lt, right_drm=Rt)
    added = stt + C
    same(added, "tt_add")
    tt_close(added.to_tt().cores, stored_list(z, "tt_add_C"))
