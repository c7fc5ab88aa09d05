"""CPU tests of the oracle: golden vectors, the reference's own fixtures, independent LAPACK cross-check."""
import itertools
import os

import numpy as np
import pytest
import scipy.linalg as sla

from tests.helpers import load_example_6x6, lu_err

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(G, "golden.npz"))


def test_pdmatgen_golden(O, golden):
    assert np.array_equal(O.pdmatgen(8, 8, 100), golden["pdmatgen_8x8_seed100"])
    assert np.array_equal(O.pdmatgen(8, 1, 200), golden["pdmatgen_8x1_seed200"])
    assert abs(O.pdmatgen(4, 4, 100)[0, 0] - 0.22712259) < 1e-8          # SURVEY 8c first value
    assert np.array_equal(O.matgen64_tile(6, 42, 0, 6, 0, 5), golden["matgen64_6x5_seed42"])


@pytest.mark.parametrize("m,n,mb,P,Q", [(17, 13, 3, 2, 3), (10, 12, 4, 2, 2), (31, 31, 2, 3, 1), (13, 13, 5, 1, 4), (23, 50, 4, 2, 2)])
def test_pdmatgen_structural_equals_closed_form(O, m, n, mb, P, Q):
    """The jump-by-jump restatement of PDMATGEN (pdmatgen.f:448-505) == closed form, on every process."""
    ag = O.pdmatgen(m, n, 100)
    for r, c in itertools.product(range(P), range(Q)):
        al = O.pdmatgen_local(m, n, mb, mb, r, c, P, Q)
        sl = O.scatter(ag, mb, mb, P, Q, r, c)
        assert np.array_equal(al, sl[:al.shape[0], :al.shape[1]])


def test_matgen64_tiles_consistent(O):
    full = O.matgen64_tile(40, 7, 0, 40, 0, 30)
    assert np.array_equal(O.matgen64_tile(40, 7, 5, 11, 3, 9), full[5:16, 3:12])
    z = O.matgen64_tile(12, 7, 0, 12, 0, 4, complex_=True)
    assert z.dtype == np.complex128 and np.all(np.abs(z.real) <= 0.5) and np.all(np.abs(z.imag) <= 0.5)
    assert len(np.unique(full)) == full.size


def test_scatter_gather_roundtrip(O):
    a = O.pdmatgen(23, 31, 5)
    for (mb, nb, P, Q) in [(2, 3, 2, 3), (4, 4, 1, 4), (5, 2, 4, 1)]:
        back = np.zeros_like(a)
        for r, c in itertools.product(range(P), range(Q)):
            O.gather_into(back, O.scatter(a, mb, nb, P, Q, r, c), mb, nb, P, Q, r, c)
        assert np.array_equal(a, back)


def test_example_6x6_fixture(O, golden):
    """EXAMPLE/pdscaex.f: PDGESV on the 6x6 system, NB=2; accept resid < 10 (pdscaex.f:181-192)."""
    A, B = load_example_6x6(os.path.join(G, "DSCAEXMAT.dat"), os.path.join(G, "DSCAEXRHS.dat"))
    lu = A.copy(order="F")
    ipiv, info = O.getrf(lu, 2)
    x = B.copy(order="F")
    O.getrs(lu, ipiv, x)
    assert info == 0
    assert np.array_equal(ipiv, golden["ex6_ipiv"]) and np.array_equal(ipiv, np.arange(1, 7))
    assert np.allclose(x, golden["ex6_x"], rtol=1e-14)
    assert np.allclose(x.ravel(), [14.6461538, 14.6461538, 15.8769231, 15.5412587, 14.6461538, 15.8769231], rtol=1e-7)
    # the example's own acceptance test: ||Ax-b|| / (||x|| ||A|| eps N) < 10
    assert O.sresid(A, x, B) < 10.0


LU_DAT = dict(MN=[(4, 4), (10, 12), (17, 13), (13, 13)], NB=[2, 3, 4], NRHS=[1, 3, 9])


@pytest.mark.parametrize("mn", LU_DAT["MN"])
@pytest.mark.parametrize("nb", LU_DAT["NB"])
def test_lu_dat_cases(O, mn, nb):
    """TESTING/traditional/LU.dat grid (threshold 1.0, LU.dat:17): factor residual always for M != N,
    solve residuals for the square cases, all NRHS of the file."""
    m, n = mn
    a0 = O.pdmatgen(m, n, 100)
    lu = a0.copy(order="F")
    ipiv, info = O.getrf(lu, nb)
    assert info == 0
    fres = O.fresid(lu, ipiv, a0)
    assert fres < 1.0 and fres - fres == 0.0
    if m == n:
        for nrhs in LU_DAT["NRHS"]:
            b0 = O.pdmatgen(n, nrhs, 200)
            x = b0.copy(order="F")
            O.getrs(lu, ipiv, x)
            assert O.sresid(a0, x, b0) < 1.0
        # LAPACK gives the same pivots on tie-free input
        _, piv = sla.lu_factor(a0)
        assert np.array_equal(ipiv - 1, piv)


@pytest.mark.parametrize("n,nb", [(64, 8), (200, 64), (500, 32)])
def test_oracle_vs_lapack(O, golden, n, nb):
    a0 = O.pdmatgen(n, n, 100)
    lu = a0.copy(order="F")
    ipiv, info = O.getrf(lu, nb)
    lus, piv = sla.lu_factor(a0)
    assert info == 0 and np.array_equal(ipiv - 1, piv)
    assert lu_err(lu, lus, a0) < 1.0
    if n == 64:
        assert np.array_equal(ipiv, golden["n64_ipiv"]) and np.allclose(lu, golden["n64_lu"], rtol=0, atol=1e-13)
    # transposed solve restated too (pdgetrs.f:273-283)
    b0 = O.pdmatgen(n, 2, 200)
    xt = b0.copy(order="F")
    O.getrs(lu, ipiv, xt, "T")
    assert np.abs(a0.T @ xt - b0).max() < 1e-9


def test_zero_pivot_info(O):
    """INFO = first zero pivot column, factorisation continues (pdgetf2.f:214-227, pdgetrf.f:263-264)."""
    a = O.pdmatgen(12, 12, 100)
    a[:, 5] = 0.0
    lu = a.copy(order="F")
    ipiv, info = O.getrf(lu, 4)
    assert info == 6 and ipiv[5] == 6


def test_complex_oracle(O):
    n = 60
    a0 = O.pzmatgen(n, n, 100)
    lu = a0.copy(order="F")
    ipiv, info = O.getrf(lu, 8)
    assert info == 0 and O.fresid(lu, ipiv, a0) < 1.0
    b0 = O.pzmatgen(n, 3, 200)
    x = b0.copy(order="F")
    O.getrs(lu, ipiv, x)
    assert O.sresid(a0, x, b0) < 1.0
    # pivot metric is |Re|+|Im| (pzamax_.c:494-497), not the modulus
    col = a0[:, 0]
    assert ipiv[0] - 1 == int(np.argmax(np.abs(col.real) + np.abs(col.imag)))


def test_index_tools(O):
    for (n, nb, P) in [(17, 3, 2), (64, 8, 4), (5, 7, 3), (100, 1, 7)]:
        assert sum(O.numroc(n, nb, p, 0, P) for p in range(P)) == n
        for src in range(P):
            for ig in range(1, n + 1):
                p = O.indxg2p(ig, nb, 0, src, P)
                il = O.indxg2l(ig, nb, 0, 0, P)
                assert O.indxl2g(il, nb, p, src, P) == ig
                assert 1 <= il <= O.numroc(n, nb, p, src, P)
