"""The oracle's restatement of the LU algorithm against the reference's OWN Fortran control flow, executed.

tests/golden/lu_reference.npz holds what SRC/pdgetrf.f + pdgetf2.f + pdlaswp.f + pdgetrs.f produce on a 1 x 1 grid when their source
text is run by the mini interpreter of tests/fortran77_mini.py (numpy standing in for the PBLAS leaves, tests/fortran_lu_runner.py;
generator: tests/golden/make_lu_golden.py).  The oracle (oracle/oracle.c) must agree: INFO and IPIV exactly -- including the peeled
first block of sub-matrix operands (pdgetrf.f:219-250), partial last blocks, M != N, NB > N and exactly-zero pivot columns -- and the
factors / solutions to rounding.  Where the reference tree is present the same is done live on fresh matrices."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS = 2.0 ** -53


def _check(O, hdr, lu_ref, ipiv_ref, xs):
    m, n, nb, mg, ng, ia, ja, zero_col, info_ref = [int(v) for v in hdr]
    a0 = O.pdmatgen(mg, ng, 100).copy(order="F")
    if zero_col >= 0:
        a0[:, zero_col] = 0.0
    sub = np.asfortranarray(a0[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n])
    lu = sub.copy(order="F")
    ipiv, info = O.getrf(lu, nb)
    assert info == info_ref
    mn = min(m, n)
    # the reference's IPIV entries are row indices of A, stored at the rows' positions (1 x 1 grid: position = global row)
    assert np.array_equal(ipiv[:mn] + (ia - 1), ipiv_ref[ia - 1:ia - 1 + mn])
    anorm = max(np.abs(sub).sum(axis=1).max(), 1e-300)
    assert np.abs(lu - lu_ref[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n]).max() / (anorm * max(m, n) * EPS) < 1.0
    outside = np.ones((mg, ng), bool); outside[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n] = False
    assert np.array_equal(lu_ref[outside], a0[outside])          # the reference touched nothing outside sub(A)
    for trans, xref in xs.items():
        x = O.pdmatgen(n, 3, 200).copy(order="F")
        O.getrs(lu, ipiv, x, trans)
        assert np.abs(x - xref).max() <= 1e-9 * np.abs(xref).max()


def test_oracle_against_the_executed_reference_fortran(O):
    g = np.load(os.path.join(ROOT, "tests", "golden", "lu_reference.npz"))
    ncases = sum(1 for k in g.files if k.startswith("case"))
    assert ncases >= 20
    for i in range(ncases):
        xs = {t: g[f"x{t}{i}"] for t in "NT" if f"x{t}{i}" in g.files}
        _check(O, g[f"case{i}"], g[f"lu{i}"], g[f"ipiv{i}"], xs)


def test_reference_fortran_lu_executed_live(O):
    if not os.path.exists("/root/reference/SRC/pdgetrf.f"):
        pytest.skip("no reference tree here")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fortran_lu_runner as R
    it = R.make()
    rng = np.random.default_rng(3)
    for _ in range(6):
        m, n, nb = int(rng.integers(1, 40)), int(rng.integers(1, 40)), int(rng.integers(1, 9))
        a0 = np.asfortranarray(rng.uniform(-1, 1, (m, n)))
        a = a0.copy(order="F")
        ipiv_ref, info_ref = R.pdgetrf(it, a, nb)
        lu = a0.copy(order="F"); ipiv, info = O.getrf(lu, nb)
        mn = min(m, n)
        assert info == info_ref and np.array_equal(ipiv[:mn], ipiv_ref[:mn])
        assert np.abs(lu - a).max() <= 1e-12 * max(1.0, np.abs(a).max())
    assert it.log == []


def test_cholesky_oracle_against_the_executed_reference_fortran(O):
    """oracle/oracle_next.c's PDPOTRF / PDPOTRS against tests/golden/chol_reference.npz = SRC/pdpotrf.f + pdpotf2.f + pdpotrs.f executed
    (tests/fortran_chol_runner.py): INFO exactly (including matrices that are not positive definite), the factored triangle and the
    solutions to rounding, the other triangle untouched."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_chol_golden as G
    g = np.load(os.path.join(ROOT, "tests", "golden", "chol_reference.npz"))
    ncases = sum(1 for k in g.files if k.startswith("case"))
    assert ncases >= 30
    for i in range(ncases):
        n, nb, u, notpd, info_ref = [int(v) for v in g[f"case{i}"]]
        uplo = chr(u)
        a0 = G.matrix(n, None if notpd < 0 else notpd)
        a = a0.copy(order="F")
        assert O.dpotrf(uplo, a, nb) == info_ref
        if info_ref != 0:
            continue
        f = g[f"f{i}"]
        tri = np.tril if uplo == "L" else np.triu
        assert np.abs(tri(a) - tri(f)).max() <= 1e-13 * np.abs(f).max()
        other = np.triu_indices(n, 1) if uplo == "L" else np.tril_indices(n, -1)
        assert np.array_equal(f[other], a0[other]) and np.array_equal(a[other], a0[other])
        x = O.pdmatgen(n, 3, 200).copy(order="F")
        O.dpotrs(uplo, a, x)
        assert np.abs(x - g[f"x{i}"]).max() <= 1e-12 * np.abs(g[f"x{i}"]).max()


def test_matrix_generator_against_the_executed_reference_fortran(O):
    """Every parity test's input comes from the oracle's closed-form (jump-ahead) PDMATGEN / PZMATGEN.  tests/golden/matgen_reference.npz
    holds what TESTING/traditional/LIN/pdmatgen.f + pzmatgen.f + pmatgeninc.f produce when their source text is executed once per process
    of P x Q grids (tests/fortran_matgen_runner.py): the oracle must reproduce every local piece and the assembled global matrix BIT FOR
    BIT, for any block size, grid shape, source process and seed -- which is also the reference's own claim that the matrix does not
    depend on the distribution."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "matgen_reference.npz"))
    ncases = sum(1 for k in g.files if k.startswith("case"))
    assert ncases >= 16
    for i in range(ncases):
        m, n, mb, nb, p, q, seed, ir, ic, z = [int(v) for v in g[f"case{i}"]]
        if z:
            assert np.array_equal(O.pzmatgen(m, n, seed), g[f"g{i}"])
            continue
        assert np.array_equal(O.pdmatgen(m, n, seed), g[f"g{i}"])
        for pr in range(p):
            for pc in range(q):
                assert np.array_equal(O.pdmatgen_local(m, n, mb, nb, pr, pc, p, q, seed, ir, ic), g[f"l{i}_{pr}_{pc}"])


def test_reference_matrix_generator_executed_live(O):
    if not os.path.exists("/root/reference/TESTING/traditional/LIN/pdmatgen.f"):
        pytest.skip("no reference tree here")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fortran_matgen_runner as R
    it = R.make()
    rng = np.random.default_rng(11)
    for _ in range(3):
        m, n, mb, nb = (int(rng.integers(1, 14)) for _ in range(4))
        p, q = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        seed = int(rng.integers(0, 2 ** 31))
        ir, ic = int(rng.integers(0, p)), int(rng.integers(0, q))
        assert np.array_equal(R.global_(it, m, n, mb, nb, p, q, seed, ir, ic), O.pdmatgen(m, n, seed))
        pr, pc = int(rng.integers(0, p)), int(rng.integers(0, q))
        assert np.array_equal(R.local(it, m, n, mb, nb, pr, pc, p, q, seed, ir, ic), O.pdmatgen_local(m, n, mb, nb, pr, pc, p, q, seed, ir, ic))
    assert np.array_equal(R.global_(it, 7, 6, 3, 2, 2, 2, 100, complex_=True), O.pzmatgen(7, 6, 100))
    assert it.log == []
