"""The oracle's restatement of the SURVEY 8(f) routines (oracle/oracle_next.c) against an independent LAPACK (scipy):
every routine restated there has a LAPACK twin with the same algorithm, so agreement is to rounding."""
import os

import numpy as np
import pytest
from scipy.linalg import lapack


@pytest.fixture(autouse=True)
def lapack_estimator(O):
    """LAPACK's DLACON carries EST between its stages; the reference's PDLACON resets it on every call (pdlacon.f:188-189) and so always
    returns the alternating-sign value.  The oracle's default is the reference's behaviour (pinned by the executed source,
    tests/test_reference_fortran.py); the LAPACK comparisons of this file run it in LAPACK's mode, which pins the iteration itself."""
    O.lacon_keep_est(True)
    yield
    O.lacon_keep_est(False)


def rnd(n, m=None, seed=0, cond=None):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, (n, m or n))
    if cond:                                   # badly scaled rows / columns: equilibration has something to do
        a = (10.0 ** rng.uniform(-cond, cond, (n, 1))) * a * (10.0 ** rng.uniform(-cond, cond, (1, m or n)))
    return np.asfortranarray(a)


@pytest.mark.parametrize("shape", [(1, 1), (7, 5), (64, 64), (130, 257)])
def test_dlange(O, shape):
    a = rnd(*shape, seed=3)
    for nm in "M1OIFE":
        assert O.dlange(nm, a) == pytest.approx(lapack.dlange(nm, a), rel=1e-14)
    assert O.dlange("1", a[:0, :]) == 0.0


@pytest.mark.parametrize("n,cond", [(5, None), (40, 3), (97, 8)])
def test_dgeequ_dlaqge(O, n, cond):
    a = rnd(n, n + 3, seed=n, cond=cond)
    r, c, rowcnd, colcnd, amax, info = O.dgeequ(a)
    r2, c2, rowcnd2, colcnd2, amax2, info2 = lapack.dgeequ(a)
    assert info == info2 == 0
    np.testing.assert_allclose(r, r2, rtol=1e-15); np.testing.assert_allclose(c, c2, rtol=1e-15)
    assert (rowcnd, colcnd, amax) == pytest.approx((rowcnd2, colcnd2, amax2), rel=1e-15)
    b = a.copy(order="F")
    eq = O.dlaqge(b, r, c, rowcnd, colcnd, amax)
    want = {(False, False): "N", (True, False): "R", (False, True): "C", (True, True): "B"}[(rowcnd < 0.1, colcnd < 0.1)]
    assert eq == want
    exp = a * (r[:, None] if eq in "RB" else 1.0) * (c[None, :] if eq in "CB" else 1.0)
    np.testing.assert_allclose(b, exp, rtol=1e-15)
    z = a.copy(order="F"); z[3, :] = 0.0
    assert O.dgeequ(z)[5] == 4 == lapack.dgeequ(z)[5]
    z = a.copy(order="F"); z[:, 2] = 0.0
    assert O.dgeequ(z)[5] == n + 3 == lapack.dgeequ(z)[5]


@pytest.mark.parametrize("n", [2, 3, 10, 100, 333])
def test_dgecon(O, n):
    for seed in range(3):
        a = rnd(n, seed=seed + 10 * n, cond=2 if seed == 2 else None)
        lu = a.copy(order="F"); ip, info = O.getrf(lu, 16)
        assert info == 0
        for nm in "1I":
            anorm = O.dlange(nm, a)
            rc = O.dgecon(nm, lu, anorm)
            rc2, inf2 = lapack.dgecon(lu, anorm, norm=nm)
            assert rc == pytest.approx(rc2, rel=1e-10)
            true = 1.0 / (anorm * np.linalg.norm(np.linalg.inv(a), 1 if nm == "1" else np.inf))
            assert true * (1 - 1e-8) <= rc <= 10 * true          # the estimate bounds ||inv(A)|| from below
    assert O.dgecon("1", np.asfortranarray(np.eye(1)), 1.0) == 1.0 and O.dgecon("1", lu, 0.0) == 0.0


def test_dgecon_as_the_reference_source_behaves(O):
    """default mode: RCOND = 1 / (ANORM * 2 ||inv(A) x_alt||_1 / (3 N)), x_alt the alternating-sign vector of pdlacon.f:360-371"""
    O.lacon_keep_est(False)
    for n in (2, 5, 40, 200):
        a = rnd(n, seed=n)
        lu = a.copy(order="F"); ip, info = O.getrf(lu, 16)
        k = np.arange(1, n + 1)
        xalt = np.where(k % 2 == 0, -1.0, 1.0) * (1.0 + (k - 1) / (n - 1))
        for nm in "1I":
            anorm = O.dlange(nm, a)
            pa = (np.tril(lu, -1) + np.eye(n)) @ np.triu(lu)              # the solves use L and U without the interchanges: P A
            y = np.linalg.solve(pa if nm == "1" else pa.T, xalt)
            expect = 1.0 / (anorm * 2.0 * np.abs(y).sum() / (3 * n))
            rc = O.dgecon(nm, lu, anorm)
            assert rc == pytest.approx(expect, rel=1e-9)
            O.lacon_keep_est(True)
            assert O.dgecon(nm, lu, anorm) <= rc * (1 + 1e-12)        # LAPACK's estimate of ||inv(A)|| is never the smaller one
            O.lacon_keep_est(False)


@pytest.mark.parametrize("trans", ["N", "T"])
@pytest.mark.parametrize("n,nrhs", [(2, 1), (50, 3), (200, 2)])
def test_dgerfs(O, trans, n, nrhs):
    a = rnd(n, seed=n + nrhs, cond=2)
    b = rnd(n, nrhs, seed=5)
    lu = a.copy(order="F"); ip, info = O.getrf(lu, 8)
    x = b.copy(order="F"); O.getrs(lu, ip, x, trans)
    ferr, berr = O.dgerfs(trans, a, lu, ip, b, x)
    # scipy exposes no dgerfs; dgesvx with FACT = 'F' is dgetrs + dgerfs on the factors it is given
    res = lapack.dgesvx(a, b, fact="F", trans=trans, af=lu, ipiv=ip.copy(), equed="N")   # IPIV goes in 1-based (and comes back 0-based, in place)
    x2, ferr2, berr2, info2 = res[7], res[9], res[10], res[11]
    assert info2 == 0
    np.testing.assert_allclose(x, x2, rtol=1e-10, atol=1e-14 * np.abs(x2).max())
    np.testing.assert_allclose(berr, berr2, rtol=0.5, atol=2e-16)   # berr is O(eps): compare magnitudes
    np.testing.assert_allclose(ferr, ferr2, rtol=0.2)
    # a perturbed solution is repaired, and FERR bounds the true error
    xt = np.linalg.solve(a if trans == "N" else a.T, b)
    xp = np.asfortranarray(x * (1.0 + 1e-7 * np.sin(np.arange(n))[:, None]))
    ferr3, berr3 = O.dgerfs(trans, a, lu, ip, b, xp)
    assert np.all(berr3 < 1e-14)
    assert np.all(np.abs(xp - xt).max(axis=0) / np.abs(xt).max(axis=0) <= ferr3 * 1.0001 + 1e-300)


@pytest.mark.parametrize("fact,trans,cond", [("N", "N", None), ("E", "N", 6), ("E", "T", 6), ("N", "T", None), ("E", "N", None)])
def test_dgesvx(O, fact, trans, cond):
    n, nrhs = 120, 3
    a = rnd(n, seed=7, cond=cond); b = rnd(n, nrhs, seed=8)
    a1, b1 = a.copy(order="F"), b.copy(order="F")
    af, x = np.zeros((n, n), order="F"), np.zeros((n, nrhs), order="F")
    ipiv, r, c = np.zeros(n, np.int32), np.zeros(n), np.zeros(n)
    eq, rcond, ferr, berr, info = O.dgesvx(fact, trans, a1, af, ipiv, "N", r, c, b1, x, nb=16)
    res = lapack.dgesvx(a, b, fact=fact, trans=trans)
    as2, lu2, ipiv2, equed2, rs2, cs2, bs2, x2, rcond2, ferr2, berr2, info2 = res
    assert info == info2 == 0
    eq2 = equed2.decode() if isinstance(equed2, bytes) else equed2
    assert eq == eq2
    if cond:
        assert eq != "N"
    np.testing.assert_array_equal(ipiv, ipiv2 + 1)
    np.testing.assert_allclose(a1, as2, rtol=1e-15)             # the equilibrated matrix
    np.testing.assert_allclose(b1, bs2, rtol=1e-15)
    assert rcond == pytest.approx(rcond2, rel=1e-8)
    np.testing.assert_allclose(x, x2, rtol=1e-9, atol=1e-13 * np.abs(x2).max())
    np.testing.assert_allclose(ferr, ferr2, rtol=0.3)
    # FACT = 'F' with the factors and scalings just computed reproduces the solution
    x3 = np.zeros((n, nrhs), order="F"); b3 = b.copy(order="F")
    if eq in "RB" and trans == "N":
        pass
    eq3, rcond3, ferr3, berr3, info3 = O.dgesvx("F", trans, a1, af, ipiv, eq, r, c, b3, x3, nb=16)
    assert info3 == 0 and eq3 == eq and rcond3 == pytest.approx(rcond, rel=1e-12)
    np.testing.assert_allclose(x3, x, rtol=1e-9, atol=1e-13 * np.abs(x).max())


def test_dgesvx_singular(O):
    n = 20
    a = rnd(n, seed=1); a[:, 5] = a[:, 4]                       # exactly singular
    b = rnd(n, 1, seed=2)
    af, x = np.zeros((n, n), order="F"), np.zeros((n, 1), order="F")
    ipiv, r, c = np.zeros(n, np.int32), np.zeros(n), np.zeros(n)
    eq, rcond, ferr, berr, info = O.dgesvx("N", "N", a.copy(order="F"), af, ipiv, "N", r, c, b.copy(order="F"), x)
    assert info > 0 and (rcond == 0.0 or info == n + 1)


@pytest.mark.parametrize("uplo", ["L", "U"])
@pytest.mark.parametrize("n,nb", [(1, 4), (7, 3), (64, 8), (150, 40), (100, 100)])
def test_dpotrf_dpotrs(O, uplo, n, nb):
    m = rnd(n, seed=n)
    a = np.asfortranarray(m @ m.T + n * np.eye(n))
    f = a.copy(order="F")
    assert O.dpotrf(uplo, f, nb) == 0
    c, info = lapack.dpotrf(a, lower=(uplo == "L"))
    tri = np.tril if uplo == "L" else np.triu
    np.testing.assert_allclose(tri(f), tri(c), rtol=1e-12, atol=1e-13 * np.abs(c).max())
    other = (np.triu_indices(n, 1) if uplo == "L" else np.tril_indices(n, -1))
    assert np.array_equal(f[other], a[other])                   # the other triangle is not referenced
    b = rnd(n, 3, seed=5); x = b.copy(order="F")
    O.dpotrs(uplo, f, x)
    np.testing.assert_allclose(x, lapack.dpotrs(c, b, lower=(uplo == "L"))[0], rtol=1e-10, atol=1e-13)
    bad = a.copy(order="F"); k = n // 2; bad[k, k] = -1.0
    assert lapack.dpotrf(bad, lower=(uplo == "L"))[1] == k + 1 == O.dpotrf(uplo, bad, nb)


@pytest.mark.parametrize("n,nb", [(1, 4), (7, 3), (64, 8), (150, 40), (33, 64)])
def test_dgetri(O, n, nb):
    a = rnd(n, seed=3 * n)
    lu = a.copy(order="F"); ip, info = O.getrf(lu, nb)
    inv = lu.copy(order="F")
    assert O.dgetri(inv, ip, nb) == 0
    inv2, info2 = lapack.dgetri(lu, ip - 1)
    assert info2 == 0
    np.testing.assert_allclose(inv, inv2, rtol=1e-9, atol=1e-12 * np.abs(inv2).max())
    if n > 2:
        lu[2, 2] = 0.0
        assert O.dgetri(lu.copy(order="F"), ip, nb) == 3 == lapack.dgetri(lu, ip - 1)[1]


def test_pblas_definitions(O):
    a, b, c = rnd(5, 4, seed=1), rnd(4, 6, seed=2), rnd(5, 6, seed=3)
    np.testing.assert_allclose(O.dgemm("N", "N", 2.0, a, b, 0.5, c), 2.0 * a @ b + 0.5 * c)
    np.testing.assert_allclose(O.dgemm("T", "T", 1.0, a.T.copy(), b.T.copy(), 0.0, np.full_like(c, np.nan)), a @ b)
    t = rnd(5, seed=4) + 3 * np.eye(5); bb = rnd(5, 3, seed=5)
    for side in "LR":
        for uplo in "LU":
            for ta in "NT":
                for dg in "NU":
                    tri = np.tril(t) if uplo == "L" else np.triu(t)
                    if dg == "U":
                        tri = tri - np.diag(np.diag(tri)) + np.eye(5)
                    op = tri.T if ta == "T" else tri
                    rhs = bb if side == "L" else bb.T.copy()
                    x = O.dtrsm(side, uplo, ta, dg, 0.5, t, rhs)
                    np.testing.assert_allclose(op @ x if side == "L" else x @ op, 0.5 * rhs, atol=1e-12)


@pytest.mark.parametrize("P,Q", [(1, 1), (1, 2), (2, 2), (2, 3), (3, 2)])
def test_redistribution_expectation_against_the_reference_itself(O, P, Q):
    """oracle/_ref: the reference's own PDGEMR2D (REDIST/SRC/pdgemr.c + pgemraux.c + pdgemr2.c compiled in place, run with one thread
    per BLACS process over oracle/ref_redist.c's mini-BLACS) must produce, bit for bit, what the parity cases of tests/next_cases.py
    expect of the product: sub(B) = sub(A), everything else of B untouched, for every layout / grid / offset case."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import next_cases
    if O.ref_redist_lib() is None:
        pytest.skip("oracle/_ref not built (no reference sources on this machine)")
    ran = 0
    for cs in next_cases.F2_CASES:
        (Pa, Qa), (Pb, Qb) = cs.get("ga", (P, Q)), cs.get("gb", (P, Q))
        if Pa * Qa > P * Q or Pb * Qb > P * Q:
            continue
        m, n, ia, ja, ib, jb = cs["m"], cs["n"], cs.get("ia", 1), cs.get("ja", 1), cs.get("ib", 1), cs.get("jb", 1)
        (rsa, csa), (rsb, csb) = cs.get("src_a", (0, 0)), cs.get("src_b", (0, 0))
        rsa, csa, rsb, csb = rsa % Pa, csa % Qa, rsb % Pb, csb % Qb
        ag = np.asfortranarray((O.pzmatgen if cs.get("z") else O.pdmatgen)(*cs["shape_a"], 100))
        out = O.ref_pdgemr2d(ag, m, n, ia, ja, ib, jb, (Pa, Qa), cs["blk_a"], (rsa, csa), cs["shape_b"], (Pb, Qb), cs["blk_b"], (rsb, csb))
        want = np.full(cs["shape_b"], -9923.0 * (1 + 1j) if cs.get("z") else -9923.0, dtype=ag.dtype, order="F"); want[ib - 1:ib - 1 + m, jb - 1:jb - 1 + n] = ag[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n]
        for r in range(Pb * Qb):
            pr, pc = divmod(r, Qb)
            exp = O.scatter(want, cs["blk_b"][0], cs["blk_b"][1], Pb, Qb, pr, pc, rsrc=rsb, csrc=csb)
            ml, nl = out[r].shape
            assert np.array_equal(out[r], exp[:ml, :nl]), (cs, r)
        ran += 1
    assert ran >= 5
