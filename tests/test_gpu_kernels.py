"""GPU parity tests, one kernel at a time, against numpy / the CPU oracle (pattern of the reference's
PBLAS testers: serial recompute of each routine, PBLAS/TESTING/pdblas3tst.f)."""
import ctypes as C

import numpy as np
import pytest

from tests.helpers import lu_err, first_mismatch

pytestmark = pytest.mark.gpu
I64 = C.c_int64


def _f(a):
    return np.asfortranarray(a)


@pytest.mark.parametrize("M,N,K,pad", [(1, 1, 1, 0), (7, 5, 3, 0), (128, 128, 16, 0), (129, 127, 17, 1), (300, 200, 64, 0),
                                       (64, 1000, 512, 0), (1000, 64, 2, 3), (513, 515, 100, 1), (2048, 1024, 512, 0)])
def test_dgemm_update(S, M, N, K, pad):
    rng = np.random.default_rng(M * 1000 + N)
    lda, ldb, ldc = M + pad, K + pad, M + 2 * pad
    A = _f(rng.uniform(-1, 1, (lda, K))); B = _f(rng.uniform(-1, 1, (ldb, N))); Cm = _f(rng.uniform(-1, 1, (ldc, N)))
    ref = Cm.copy(order="F")
    ref[:M, :] -= A[:M, :] @ B[:K, :]
    out = Cm.copy(order="F")
    S.lib().slb200_test_gemm(I64(M), I64(N), K, S.api._ptr(A), I64(lda), S.api._ptr(B), I64(ldb), S.api._ptr(out), I64(ldc), 0, 1)
    err = np.abs(out - ref).max()
    assert err <= 4 * K * 2.0 ** -53 * 4, (err, np.unravel_index(np.abs(out - ref).argmax(), out.shape))
    assert np.array_equal(out[M:, :], Cm[M:, :])          # rows beyond M untouched


@pytest.mark.parametrize("M,N,K", [(5, 3, 2), (64, 64, 16), (100, 130, 33), (256, 192, 256),
                                   (2048, 300, 256), (2500, 1031, 100), (4096, 129, 37), (3000, 64, 16)])   # M >= 2048: packed real kernel
def test_zgemm_update(S, M, N, K):
    rng = np.random.default_rng(M + N + K)
    cz = lambda r, c: _f(rng.uniform(-1, 1, (r, c)) + 1j * rng.uniform(-1, 1, (r, c)))
    A, B, Cm = cz(M, K), cz(K, N), cz(M, N)
    ref = Cm - A @ B
    out = Cm.copy(order="F")
    ldc = M + 3                                                  # guard rows below C
    outp = np.full((ldc, N), -9923.0 + 0j, order="F"); outp[:M, :] = Cm
    S.lib().slb200_test_gemm(I64(M), I64(N), K, S.api._ptr(A), I64(M), S.api._ptr(B), I64(K), S.api._ptr(outp), I64(ldc), 1, 1)
    assert np.abs(outp[:M, :] - ref).max() <= 64 * K * 2.0 ** -53
    assert np.all(outp[M:, :] == -9923.0)


@pytest.mark.parametrize("jb,n,cplx", [(2, 5, 0), (3, 1, 0), (64, 100, 0), (65, 33, 0), (200, 300, 0), (512, 1000, 0), (40, 50, 1), (256, 64, 1)])
def test_trsm_llnu(S, jb, n, cplx):
    rng = np.random.default_rng(jb + n)
    if cplx:
        L = _f(np.tril(rng.uniform(-1, 1, (jb, jb)) + 1j * rng.uniform(-1, 1, (jb, jb)), -1) * 0.7 + np.eye(jb))
        B = _f(rng.uniform(-1, 1, (jb, n)) + 1j * rng.uniform(-1, 1, (jb, n)))
    else:
        L = _f(np.tril(rng.uniform(-1, 1, (jb, jb)), -1) + np.eye(jb))
        B = _f(rng.uniform(-1, 1, (jb, n)))
    # garbage above / on the diagonal must be ignored (unit lower)
    Lg = L + np.triu(np.full((jb, jb), 7.0), 0)
    Lg = _f(Lg)
    import scipy.linalg as sla
    ref = sla.solve_triangular(L, B, lower=True, unit_diagonal=True)
    out = B.copy(order="F")
    S.lib().slb200_test_trsm(jb, I64(n), S.api._ptr(Lg), I64(jb), S.api._ptr(out), I64(jb), cplx)
    scale = np.abs(ref).max()
    assert np.abs(out - ref).max() / scale < 1e-9 * max(1, jb / 8)


@pytest.mark.parametrize("m,jb", [(2, 2), (4, 2), (10, 3), (17, 4), (33, 32), (100, 32), (100, 33), (1000, 64), (3000, 100),
                                  (5000, 512), (40000, 512), (70000, 256)])
def test_panel_real(S, O, m, jb):
    a0 = O.pdmatgen(m, jb, 100)
    ref = a0.copy(order="F")
    ipr, infr = O.getrf(ref, jb)
    out = a0.copy(order="F")
    ipiv = np.zeros(jb, np.int32); info = C.c_int(-7)
    ms = S.lib().slb200_test_panel(m, jb, S.api._ptr(out), I64(m), ipiv.ctypes.data_as(C.c_void_p), C.byref(info), 0)
    assert info.value == infr == 0
    assert np.array_equal(ipiv, ipr), ("first pivot mismatch at column", first_mismatch(ipiv, ipr), ipiv[:8], ipr[:8])
    assert lu_err(out, ref, a0) < 1.0
    print(f"panel {m}x{jb}: {ms:.3f} ms")


def test_panel_ties_and_zero_column(S, O):
    """Ties go to the first (lowest) row like idamax; an all-zero column gives INFO and IPIV(j)=j (pdamax_.c:486)."""
    a0 = np.asfortranarray(np.array([[1., 2, 3, 4], [-1, 2, 0, 1], [1, -2, 3, 4], [1, 2, 3, 5], [0.5, 1, 1, 1]]))
    ref = a0.copy(order="F"); ipr, infr = O.getrf(ref, 4)
    out = a0.copy(order="F"); ipiv = np.zeros(4, np.int32); info = C.c_int(0)
    S.lib().slb200_test_panel(5, 4, S.api._ptr(out), I64(5), ipiv.ctypes.data_as(C.c_void_p), C.byref(info), 0)
    assert np.array_equal(ipiv, ipr) and info.value == infr
    assert np.allclose(out, ref, rtol=0, atol=1e-14)
    a0 = O.pdmatgen(40, 8, 100); a0[:, 3] = 0.0
    ref = a0.copy(order="F"); ipr, infr = O.getrf(ref, 8)
    out = a0.copy(order="F"); ipiv = np.zeros(8, np.int32)
    S.lib().slb200_test_panel(40, 8, S.api._ptr(out), I64(40), ipiv.ctypes.data_as(C.c_void_p), C.byref(info), 0)
    assert infr == 4 and info.value == 4 and np.array_equal(ipiv, ipr) and ipiv[3] == 4


@pytest.mark.parametrize("m,jb", [(6, 3), (50, 16), (700, 64), (3000, 256)])
def test_panel_complex(S, O, m, jb):
    a0 = O.pzmatgen(m, jb, 100)
    ref = a0.copy(order="F"); ipr, infr = O.getrf(ref, jb)
    out = a0.copy(order="F"); ipiv = np.zeros(jb, np.int32); info = C.c_int(0)
    S.lib().slb200_test_panel(m, jb, S.api._ptr(out), I64(m), ipiv.ctypes.data_as(C.c_void_p), C.byref(info), 1)
    assert info.value == 0 and np.array_equal(ipiv, ipr), first_mismatch(ipiv, ipr)
    assert lu_err(out, ref, a0) < 1.0


@pytest.mark.parametrize("m,n,j0,jb", [(10, 7, 0, 4), (64, 33, 16, 16), (1000, 300, 128, 64), (5000, 129, 512, 512)])
def test_laswp_block(S, m, n, j0, jb):
    rng = np.random.default_rng(m + n)
    a = _f(rng.uniform(-1, 1, (m, n)))
    piv = np.array([rng.integers(j0 + t, m) for t in range(jb)], dtype=np.int32)
    piv[::5] = j0 + np.arange(jb)[::5]              # some "no swap" entries
    if jb > 3:
        piv[2] = piv[1]                               # a repeated outside pivot row
    ref = a.copy(order="F")
    for t in range(jb):                               # PDLASWP forward (pdlaswp.f:163-172)
        p = piv[t]
        if p != j0 + t:
            ref[[j0 + t, p], :] = ref[[p, j0 + t], :]
    out = a.copy(order="F")
    ip1 = (piv + 1).astype(np.int32)
    S.lib().slb200_test_laswp(m, I64(n), S.api._ptr(out), I64(m), j0, jb, ip1.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, ref)                   # pure data movement: bit exact


def test_microbench_peaks(S):
    L = S.lib()
    dmma = L.slb200_bench_dmma_tflops(20000)
    dfma = L.slb200_bench_dfma_tflops(20000)
    gbs = L.slb200_bench_copy_gbs(I64(2 << 30))
    print(f"PEAKS dmma={dmma:.2f} TFLOP/s dfma={dfma:.2f} TFLOP/s copy={gbs:.1f} GB/s")
    assert dmma > 5 and dfma > 5 and gbs > 1000
