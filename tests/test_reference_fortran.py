"""The oracle's restatement of the LU algorithm against the reference's OWN Fortran control flow, executed.

tests/golden/lu_reference.npz holds what SRC/pdgetrf.f + pdgetf2.f + pdlaswp.f + pdgetrs.f produce on a 1 x 1 grid when their source
text is run by the mini interpreter of tests/fortran77_mini.py (numpy standing in for the PBLAS leaves, tests/fortran_lu_runner.py;
generator: tests/golden/make_lu_golden.py).  The oracle (oracle/oracle.c) must agree: INFO and IPIV exactly -- including the peeled
first block of sub-matrix operands (pdgetrf.f:219-250), partial last blocks, M != N, NB > N and exactly-zero pivot columns -- and the
factors / solutions to rounding.  Where the reference tree is present the same is done live on fresh matrices."""
import os
import re
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


def test_condition_estimate_against_the_executed_reference_fortran(O):
    """oracle/oracle_next.c's PDGECON (default mode) against tests/golden/refine_reference.npz = SRC/pdgecon.f + pdlacon.f + pdlatrs.f
    executed (tests/fortran_refine_runner.py).  The executed source is what showed that the reference's PDLACON resets EST on every call
    (pdlacon.f:188-189) and therefore returns the alternating-sign value, not the maximum the iteration found: LAPACK's DGECON -- the
    oracle's previous pin -- returns a different (smaller) RCOND on the same factors."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_refine_golden as G
    from scipy.linalg import lapack
    g = np.load(os.path.join(ROOT, "tests", "golden", "refine_reference.npz"))
    ncases = sum(1 for k in g.files if k.startswith("case"))
    assert ncases >= 13
    O.lacon_keep_est(False)
    differs = 0
    for i in range(ncases):
        n, nb, scale, off = [int(v) for v in g[f"case{i}"]]
        a = G.matrix(dict(n=n, scale=scale))
        lu = a.copy(order="F"); ipiv, info = O.getrf(lu, nb)
        for k, norm in enumerate("1I"):
            anorm = np.abs(a).sum(axis=0).max() if norm == "1" else np.abs(a).sum(axis=1).max()
            rc = O.dgecon(norm, lu, anorm)
            assert rc == pytest.approx(float(g[f"rcond{i}"][k]), rel=1e-9), (n, nb, norm)
            rc_lapack = lapack.dgecon(lu, anorm, norm=norm)[0]
            assert rc_lapack <= rc * (1 + 1e-9)
            differs += rc_lapack < 0.99 * rc
    assert differs >= ncases                                     # the two estimators do not agree: the pin matters
    # INFO codes / quick returns of the executed source (RCOND untouched = -7 when an argument is illegal)
    for nm, n, anorm, lwork, info, rcond in g["quick"]:
        if info == 0:
            lu = O.pdmatgen(8, 8, 100)[:int(n), :int(n)].copy(order="F")
            assert O.dgecon(chr(int(nm)), lu, float(anorm)) == rcond


def test_reference_condition_estimate_executed_live(O):
    if not os.path.exists("/root/reference/SRC/pdlacon.f"):
        pytest.skip("no reference tree here")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fortran_refine_runner as R
    it = R.make()
    rng = np.random.default_rng(5)
    O.lacon_keep_est(False)
    for _ in range(4):
        n, nb = int(rng.integers(2, 30)), int(rng.integers(1, 9))
        a = np.asfortranarray(rng.uniform(-1, 1, (n, n)))
        lu = a.copy(order="F"); ipiv, info = O.getrf(lu, nb)
        for norm in "1I":
            anorm = O.dlange(norm, a)
            rc, info = R.pdgecon(it, norm, lu, anorm, nb)
            assert info == 0 and O.dgecon(norm, lu, anorm) == pytest.approx(rc, rel=1e-9)
    assert it.log == []


def test_refinement_against_the_executed_reference_fortran(O):
    """oracle/oracle_next.c's PDGERFS against SRC/pdgerfs.f + pdlacon.f executed (tests/golden/refine_reference.npz): the refined solution
    within its own error bound, BERR at rounding level on both sides, FERR -- built on the reference's PDLACON, i.e. on its
    alternating-sign value -- to a few per cent (it is an estimate of a quantity that moves with the rounding of the residual)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_refine_golden as G
    g = np.load(os.path.join(ROOT, "tests", "golden", "refine_reference.npz"))
    ncases = sum(1 for k in g.files if re.fullmatch(r"rfs\d+", k))
    assert ncases >= 8
    O.lacon_keep_est(False)
    for i in range(ncases):
        n, nb, nrhs, tr, scale = [int(v) for v in g[f"rfs{i}"]]
        cs = dict(n=n, nb=nb, nrhs=nrhs, trans=chr(tr), scale=scale, perturb=float(g[f"rfs_pert{i}"][0]))
        a, lu, ipiv, b, x = G.rfs_inputs(cs)
        ferr, berr = O.dgerfs(cs["trans"], a, lu, ipiv, b, x)
        xr, fr, br = g[f"rfs_x{i}"], g[f"rfs_ferr{i}"], g[f"rfs_berr{i}"]
        # LAPACK's estimator gives a different (larger) FERR on the same data: the pin distinguishes the two.  It is also the reliable
        # bound: with the alternating-sign value alone FERR can fall an order of magnitude short of the true error (n = 17, TRANS = T
        # below: two refinements that both reach BERR ~ eps differ by 12 x the reference's FERR), so X is compared within LAPACK's.
        O.lacon_keep_est(True)
        x2 = G.rfs_inputs(cs)[4]
        f2, _ = O.dgerfs(cs["trans"], a, lu, ipiv, b, x2)
        O.lacon_keep_est(False)
        assert np.all(f2 >= ferr * (1 - 1e-9))
        for k in range(nrhs):
            assert np.abs(x[:, k] - xr[:, k]).max() <= 2.0 * f2[k] * np.abs(xr[:, k]).max(), (cs, k)
            assert berr[k] <= 4 * EPS * (n + 1) and br[k] <= 4 * EPS * (n + 1)
            assert ferr[k] == pytest.approx(fr[k], rel=0.1), (cs, k)


def test_expert_driver_against_the_executed_reference_fortran(O):
    """oracle/oracle_next.c's PDGESVX against SRC/pdgesvx.f executed WITH its callees pdgeequ.f, pdlaqge.f, pdlange.f, pdgecon.f,
    pdlacon.f, pdgerfs.f (PDGETRF / PDGETRS = the oracle's, themselves pinned by the executed pdgetrf.f / pdgetrs.f).  Discrete outputs
    exactly: INFO (0, N + 1 for a matrix singular to working precision, k for a zero pivot), EQUED (N / R / C / B), IPIV; R, C and the
    equilibrated A and B exactly (maxima and reciprocals); RCOND to 1e-9; FERR to 10 %; X within LAPACK's bound; FACT = 'F' on the
    returned factors reproduces the solution."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_refine_golden as G
    g = np.load(os.path.join(ROOT, "tests", "golden", "refine_reference.npz"))
    assert len(G.SVX_CASES) >= 13
    seen = set()
    for i, cs in enumerate(G.SVX_CASES):
        n, nrhs, nb = cs["n"], cs["nrhs"], cs["nb"]
        info_ref, eq_ref = int(g[f"svx{i}"][0]), chr(int(g[f"svx{i}"][1]))
        seen.add((min(info_ref, 1) if info_ref <= n else 2, eq_ref))

        def run(fact, equed, a, af, ip, r, c, b, keep=False):
            O.lacon_keep_est(keep)
            x = np.zeros((n, nrhs), order="F")
            out = O.dgesvx(fact, cs["trans"], a, af, ip, equed, r, c, b, x, nb=nb)
            O.lacon_keep_est(False)
            return out, x
        a, b = G.svx_inputs(cs)
        af, ip, r, c = np.zeros((n, n), order="F"), np.zeros(n, np.int32), np.zeros(n), np.zeros(n)
        (eq, rcond, ferr, berr, info), x = run(cs["fact"], "N", a, af, ip, r, c, b)
        assert (info, eq) == (info_ref, eq_ref), cs
        assert rcond == pytest.approx(float(g[f"svx_rcond{i}"][0]), rel=1e-9, abs=1e-300), cs
        assert np.array_equal(a, g[f"svx_a{i}"]) and np.array_equal(b, g[f"svx_b{i}"]), cs          # equilibrated in place (or untouched)
        if eq_ref in "RB":
            assert np.array_equal(r, g[f"svx_r{i}"])
        if eq_ref in "CB":
            assert np.array_equal(c, g[f"svx_c{i}"])
        if info_ref == 0 or info_ref == n + 1:
            assert np.array_equal(ip, g[f"svx_ipiv{i}"]), cs
        if info_ref != 0:
            continue
        a2, b2 = G.svx_inputs(cs)
        (_, _, fbound, _, _), _ = run(cs["fact"], "N", a2, np.zeros((n, n), order="F"), np.zeros(n, np.int32), np.zeros(n), np.zeros(n), b2, keep=True)
        xr, fr = g[f"svx_x{i}"], g[f"svx_ferr{i}"]
        for k in range(nrhs):
            assert np.abs(x[:, k] - xr[:, k]).max() <= 2.0 * max(fbound[k], 1e-15) * np.abs(xr[:, k]).max(), (cs, k)
            assert ferr[k] == pytest.approx(fr[k], rel=0.1, abs=1e-300), (cs, k)
            assert berr[k] <= 4 * EPS * (n + 1) and g[f"svx_berr{i}"][k] <= 4 * EPS * (n + 1)
        # FACT = 'F' with the factors, scalings and EQUED just returned
        a3, b3 = G.svx_inputs(cs)
        if eq in "RB":
            a3 = np.asfortranarray(r[:, None] * a3)
        if eq in "CB":
            a3 = np.asfortranarray(a3 * c[None, :])
        (eqf, rcf, _, _, inff), xf = run("F", eq, a3, af.copy(order="F"), ip.copy(), r.copy(), c.copy(), b3)
        assert (inff, eqf) == (int(g[f"svxF{i}"][0]), chr(int(g[f"svxF{i}"][1])))
        assert rcf == pytest.approx(float(g[f"svxF_rcond{i}"][0]), rel=1e-9)
        stale_work = eq in "CB" and cs["trans"] == "N"
        for k in range(nrhs):
            tol = 2.0 * max(fbound[k], 1e-15) * np.abs(xr[:, k]).max()
            assert np.abs(xf[:, k] - x[:, k]).max() <= tol                     # the oracle's FACT = 'F' reproduces its own solution
            if not stale_work:
                assert np.abs(xf[:, k] - g[f"svxF_x{i}"][:, k]).max() <= tol
            else:
                # A defect of the reference, shown by the execution and NOT reproduced: with FACT = 'F', column scaling and TRANS = 'N'
                # ICOLEQU stays 0 (it is only set under FACT = 'E', pdgesvx.f:665-680), so C is never copied into WORK and X is
                # multiplied by whatever PDGERFS left there (pdgesvx.f:796-822): the executed result is not the solution.
                assert np.abs(g[f"svxF_x{i}"][:, k] - xr[:, k]).max() > 0.5 * np.abs(xr[:, k]).max()
    assert {(0, "N"), (0, "B"), (0, "R"), (0, "C"), (2, "N"), (1, "N")} <= seen


def test_inverse_against_the_executed_reference_fortran(O):
    """oracle/oracle_next.c's PDGETRI against SRC/pdgetri.f + pdtrtri.f + pdtrti2.f executed (tests/golden/refine_reference.npz): INFO
    exactly (the first exactly-zero U(i,i); nothing is overwritten then), the inverse to rounding, everything outside sub(A) untouched,
    and the workspace sizes the executed source reports against the formulas of pdgetri.f:202-229."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_refine_golden as G
    g = np.load(os.path.join(ROOT, "tests", "golden", "refine_reference.npz"))
    assert len(G.TRI_CASES) >= 12
    for i, cs in enumerate(G.TRI_CASES):
        n, nb, off = cs["n"], cs["nb"], cs.get("off", 0)
        big, ipiv, lu, a = G.tri_inputs(cs)
        info_ref, lw, liw = [int(v) for v in g[f"tri{i}"]]
        ref = g[f"tri_inv{i}"]
        inv = lu.copy(order="F")
        info = O.dgetri(inv, (ipiv - off).astype(np.int32), nb)
        assert info == info_ref, cs
        assert (lw, liw) == (n * nb, (n + off) + nb)                     # LWMIN = LOCr(N + IROFF) NB, LIWMIN = LOCc(N_A) + NB on a square grid
        outside = np.ones(big.shape, bool); outside[off:, off:] = False
        assert np.array_equal(ref[outside], big[outside])
        if info_ref > 0:
            assert info_ref == cs["zero"] + 1 and np.array_equal(ref, big) and np.array_equal(inv, lu)
            continue
        scale = np.abs(ref[off:, off:]).max()
        assert np.abs(inv - ref[off:, off:]).max() <= 1e-12 * scale * max(1.0, np.linalg.cond(a) * 1e-3), cs
        assert np.abs(ref[off:, off:] @ a - np.eye(n)).max() <= 1e-10 * max(1.0, np.linalg.cond(a) * 1e-3)


def test_reference_row_interchanges_of_the_solve_executed_live(O):
    """SRC/pdlapiv.f + pdlapv2.f (the PDLAPIV of PDGETRS, forward for TRANS = N and backward for T) executed as well: the solutions are
    bit-identical to the ones the golden vectors were made with (where PDLAPIV was a numpy stand-in written from its Purpose block)."""
    if not os.path.exists("/root/reference/SRC/pdlapv2.f"):
        pytest.skip("no reference tree here")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fortran_lu_runner as R
    it0, it1 = R.make(), R.make(real_lapiv=True)
    for n, nb, nrhs in [(6, 2, 2), (17, 4, 3), (13, 5, 1), (9, 16, 2)]:
        lu = O.pdmatgen(n, n, 100).copy(order="F")
        ipiv, info = R.pdgetrf(it0, lu, nb)
        for trans in "NT":
            b0 = O.pdmatgen(n, nrhs, 200).copy(order="F"); b1 = b0.copy(order="F")
            assert R.pdgetrs(it0, trans, lu, ipiv, b0, nb) == 0 and R.pdgetrs(it1, trans, lu, ipiv, b1, nb) == 0
            assert np.array_equal(b0, b1)
            x = O.pdmatgen(n, nrhs, 200).copy(order="F")
            O.getrs(lu, ipiv[:n], x, trans)
            assert np.abs(x - b1).max() <= 1e-9 * np.abs(b1).max()
    assert it1.log == []


def test_complex_lu_source_is_the_real_one_with_type_names_swapped():
    """The complex path (PZGETRF / PZGETF2 / PZLASWP / PZGETRS) needs no separate execution: its source IS the real routines'
    source with the type names swapped (COMPLEX*16, PZ*, PZGERU for PDGER, complex constants), statement for statement.  Checked on the
    reference tree where it is present; with it, the executed-source pins of the real LU path carry over to the oracle's complex one."""
    if not os.path.exists("/root/reference/SRC/pzgetrf.f"):
        pytest.skip("no reference tree here")

    def statements(path, swap):
        out = []
        for raw in open(path).read().splitlines():
            if not raw.strip() or raw[0] in "*Cc!":
                continue
            body = raw[6:72]
            if len(raw) > 5 and raw[5] not in " 0" and out:
                out[-1] += " " + body.strip()
            else:
                out.append(body.strip())
        norm = []
        for st in out:
            st = st.upper()
            if swap:
                st = st.replace("COMPLEX*16", "DOUBLE PRECISION").replace("PZGERU", "PDGER").replace("PZ", "PD")
                st = re.sub(r"\(\s*([0-9.D+-]+)\s*,\s*0\.0D\+0\s*\)", r"\1", st)        # ( 1.0D+0, 0.0D+0 ) -> 1.0D+0
            if st.startswith("EXTERNAL"):
                st = "EXTERNAL " + ",".join(sorted(x.strip() for x in st[8:].split(",")))   # the lists are ordered alphabetically
            norm.append(re.sub(r"\s+", "", st))
        return norm
    for name in ("getrf", "getf2", "laswp", "getrs"):
        z = statements(f"/root/reference/SRC/pz{name}.f", True)
        d = statements(f"/root/reference/SRC/pd{name}.f", False)
        diff = [(a, b) for a, b in zip(z, d) if a != b]
        if name == "getrs":
            # the one real difference: the transposed solves pass TRANS ('T' or 'C') on where the real routine writes 'Transpose'
            assert len(diff) == 2 and all(a.replace(",TRANS,", ",'TRANSPOSE',") == b and a.startswith("CALLPDTRSM(") for a, b in diff), diff
            diff = []
        assert len(z) == len(d) and not diff, (name, diff[:3])


def test_solve_residual_against_the_executed_reference_fortran(O):
    """SRESID -- the number every bench line and parity test reports for a solve -- is TESTING/traditional/LIN/pdlaschk.f's:
    max_j ||b_j - A x_j||_inf / (||x_j||_inf ANORM eps N) with A and b REGENERATED block by block by PDMATGEN.  The oracle's orc_sresid
    against that file executed (with the executed generator underneath, tests/fortran_matgen_runner.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_matgen_golden as G
    g = np.load(os.path.join(ROOT, "tests", "golden", "matgen_reference.npz"))
    for i, (n, nrhs, nb, nbr, pert) in enumerate(G.CHK_CASES):
        a, b, x = G.chk_solution(O, n, nrhs, pert)
        ref = float(g[f"chk{i}"][0])
        mine = O.sresid(a, x, b)
        if pert == 0.0:
            assert ref < 10.0 and mine < 10.0                   # rounding level on both sides (the reference's threshold is 1 ... 3)
        else:
            assert mine == pytest.approx(ref, rel=1e-3), (n, nrhs, nb, nbr)


def test_factorisation_residual_against_the_executed_reference_fortran(O):
    """FRESID -- the parity tests' and the bench pre-flight's measure of a factorisation -- is the LU driver's: PDGETRRV rebuilds P L U
    from the factors, PDLAFCHK subtracts the regenerated A and divides ||.||_inf by max(M, N) eps ||A||_inf (pdludriver.f:540-556).  The
    oracle's orc_fresid against TESTING/traditional/LIN/pdgetrrv.f + pdlafchk.f + SRC/pdlange.f + the generator, executed."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_refine_golden as G
    g = np.load(os.path.join(ROOT, "tests", "golden", "refine_reference.npz"))
    for i, (m, n, nb, pert) in enumerate(G.FCHK_CASES):
        a, lu, ipiv, _ = G.fchk_inputs(m, n, nb, pert)
        ref, anorm = [float(v) for v in g[f"fchk{i}"]]
        assert anorm == np.abs(a).sum(axis=1).max()
        mine = O.fresid(lu, ipiv, a)
        if pert == 0.0:
            assert ref < 1.0 and mine < 1.0                     # rounding only: the order of the products differs, the level does not
        else:
            assert mine == pytest.approx(ref, rel=1e-3), (m, n, nb)
