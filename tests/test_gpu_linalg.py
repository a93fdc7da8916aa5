"""Dense building blocks against NumPy/SciPy float64: strided split-K GEMM, Jacobi pseudo-inverse
(gelsd cut-off), Householder QR with LAPACK's sign convention.  Tolerance 1e-12 relative."""
import numpy as np
import pytest
import scipy.linalg

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (20, 40, 10000), (10, 15, 160000), (3000, 20, 40), (65, 33, 31), (200, 1, 100)])
def test_gemm_strided(M, N, K):
    from tt_sketch import _backend as be

    rng = np.random.default_rng(0)
    A, B = rng.standard_normal((M, K)), rng.standard_normal((K, N))
    dA, dB = be.to_device(A), be.to_device(B)
    assert _rel(be.to_host(be.gemm(dA, dB)), A @ B) < 1e-12
    dAt, dBt = be.to_device(np.ascontiguousarray(A.T)), be.to_device(np.ascontiguousarray(B.T))
    assert _rel(be.to_host(be.gemm(dAt.T, dBt.T)), A @ B) < 1e-12
    C0 = rng.standard_normal((M, N))
    out = be.to_device(C0.copy())
    be.gemm(dA, dB, out=out, beta=1.0)
    assert _rel(be.to_host(out), C0 + A @ B) < 1e-12
    if N > 4:
        assert _rel(be.to_host(be.gemm(dA, dB[:, 1:4])), A @ B[:, 1:4]) < 1e-12


@pytest.mark.parametrize("m,n", [(20, 40), (40, 20), (1, 5), (7, 7), (100, 200), (3, 1)])
def test_pinv(m, n):
    from tt_sketch import _backend as be

    rng = np.random.default_rng(1)
    A = rng.standard_normal((m, n))
    assert _rel(be.to_host(be.pinv(be.to_device(A))), np.linalg.pinv(A)) < 1e-11
    if min(m, n) >= 5:  # rank deficient, explicit cut-off (same rule as lstsq(cond=rcond): s <= rcond*s_max dropped)
        A[:, -1] = A[:, 0] * 2 - A[:, 1]
        if m <= n:
            A[-1] = A[0] + A[1]
        B = rng.standard_normal((m, 6))
        want = scipy.linalg.lstsq(A, B, cond=1e-10)[0]
        got = be.to_host(be.gemm(be.pinv(be.to_device(A), 1e-10), be.to_device(B)))
        assert _rel(got, want) < 1e-9
        # a singular value well above eps*s_max is kept by the default cut-off, like gelsd
        U, _, Vt = np.linalg.svd(rng.standard_normal((m, n)), full_matrices=False)
        sv = np.logspace(0, -12, min(m, n))
        A2 = (U * sv) @ Vt
        got2 = be.to_host(be.pinv(be.to_device(A2)))
        assert _rel(got2 @ A2 @ got2, got2) < 1e-3 and np.linalg.norm(got2, 2) > 1e11


@pytest.mark.parametrize("m,n", [(1000, 20), (50, 50), (37, 5), (5, 1), (20000, 40), (200000, 40), (9001, 64), (8192, 1), (30000, 33)])
def test_qr_matches_lapack(m, n):
    from tt_sketch import _backend as be

    rng = np.random.default_rng(2)
    A = rng.standard_normal((m, n))
    q_want, _ = scipy.linalg.qr(A, mode="economic")
    q = be.to_host(be.qr_q_inplace(be.to_device(A.copy())))   # m >= 8192: the grid-cooperative kernel (one row block per SM)
    assert _rel(q, q_want) < 1e-11
    assert np.allclose(q.T @ q, np.eye(n), atol=1e-12)


def test_qr_grid_kernel_handles_dependent_columns():
    """A zero column (tau = 0) and an exactly dependent column in a tall panel: same Q as LAPACK where it is determined."""
    from tt_sketch import _backend as be

    rng = np.random.default_rng(5)
    A = rng.standard_normal((10000, 6))
    A[:, 2] = 0.0
    q_want, _ = scipy.linalg.qr(A, mode="economic")
    q = be.to_host(be.qr_q_inplace(be.to_device(A.copy())))
    assert _rel(q[:, :2], q_want[:, :2]) < 1e-11
    assert np.allclose(q.T @ q, np.eye(6), atol=1e-10)


def test_orth_step_matches_reference_formula():
    from tt_sketch.sketch_dispatch import orth_step

    rng = np.random.default_rng(3)
    Psi, Om = rng.standard_normal((4, 30, 12)), rng.standard_normal((6, 12))
    m = scipy.linalg.lstsq(Om.T, Psi.reshape(120, 12).T)[0].T
    q_want, _ = scipy.linalg.qr(m, mode="economic")
    assert _rel(orth_step(Psi, Om), q_want.reshape(4, 30, 6)) < 1e-10
    q2, _ = scipy.linalg.qr(Psi.reshape(120, 12), mode="economic")
    assert _rel(orth_step(Psi, None), q2.reshape(4, 30, 12)) < 1e-10
