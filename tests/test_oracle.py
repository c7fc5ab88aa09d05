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


def _blocked_right_looking(a, nb):
    """Plain blocked right-looking LU with partial pivoting (the order PDGETRF works in), numpy only."""
    a = a.copy(order="F"); m, n = a.shape; mn = min(m, n)
    piv = np.zeros(mn, np.int64)
    for j0 in range(0, mn, nb):
        jb = min(nb, mn - j0)
        for j in range(j0, j0 + jb):                                   # unblocked panel
            p = j + int(np.argmax(np.abs(a[j:, j])))
            piv[j] = p
            if p != j:
                a[[j, p], j0:j0 + jb] = a[[p, j], j0:j0 + jb]
            if a[j, j] != 0.0:
                a[j + 1:, j] *= 1.0 / a[j, j]
            a[j + 1:, j + 1:j0 + jb] -= np.outer(a[j + 1:, j], a[j, j + 1:j0 + jb])
        for j in range(j0, j0 + jb):                                   # interchanges left and right of the panel
            p = piv[j]
            if p != j:
                a[[j, p], :j0] = a[[p, j], :j0]
                a[[j, p], j0 + jb:] = a[[p, j], j0 + jb:]
        l11 = np.tril(a[j0:j0 + jb, j0:j0 + jb], -1) + np.eye(jb)
        a[j0:j0 + jb, j0 + jb:] = np.linalg.solve(l11, a[j0:j0 + jb, j0 + jb:])
        a[j0 + jb:, j0 + jb:] -= a[j0 + jb:, j0:j0 + jb] @ a[j0:j0 + jb, j0 + jb:]
    return a, piv


def _slab_left_looking(a, nb, ws):
    """The schedule planned for overlapping the host->device upload (open item 2): columns arrive in slabs of
    ws; a slab is first CAUGHT UP with every panel factored so far (interchanges, U12 solve, update -- in step order), then
    its own panels are factored right-looking inside the slab.  The interchanges of the already factored columns are
    DEFERRED: L stays in the row order of the step that produced it, and one permutation pass per block column at the
    end brings it to LAPACK's layout."""
    a = a.copy(order="F"); m, n = a.shape; mn = min(m, n)
    piv = np.zeros(mn, np.int64)
    steps = []                                                          # (j0, jb) of the panels factored so far

    def apply_step(j0, jb, c0, c1):                                     # step (j0, jb) applied to columns [c0, c1)
        for j in range(j0, j0 + jb):
            p = piv[j]
            if p != j:
                a[[j, p], c0:c1] = a[[p, j], c0:c1]
        l11 = np.tril(a[j0:j0 + jb, j0:j0 + jb], -1) + np.eye(jb)
        a[j0:j0 + jb, c0:c1] = np.linalg.solve(l11, a[j0:j0 + jb, c0:c1])
        a[j0 + jb:, c0:c1] -= a[j0 + jb:, j0:j0 + jb] @ a[j0:j0 + jb, c0:c1]

    for s0 in range(0, n, ws):
        s1 = min(n, s0 + ws)
        for (j0, jb) in steps:                                          # catch-up, in step order
            apply_step(j0, jb, s0, s1)
        for j0 in range(s0, min(s1, mn), nb):                           # the slab's own panels
            jb = min(nb, mn - j0)
            for j in range(j0, j0 + jb):
                p = j + int(np.argmax(np.abs(a[j:, j])))
                piv[j] = p
                if p != j:
                    a[[j, p], j0:j0 + jb] = a[[p, j], j0:j0 + jb]
                if a[j, j] != 0.0:
                    a[j + 1:, j] *= 1.0 / a[j, j]
                a[j + 1:, j + 1:j0 + jb] -= np.outer(a[j + 1:, j], a[j, j + 1:j0 + jb])
            steps.append((j0, jb))
            apply_step(j0, jb, j0 + jb, s1)                             # trailing columns of THIS slab only
    for (j0, jb) in steps:                                              # deferred left interchanges, one pass per block column
        for (k0, kb) in steps:
            if k0 <= j0:
                continue
            for j in range(k0, k0 + kb):
                p = piv[j]
                if p != j:
                    a[[j, p], j0:j0 + jb] = a[[p, j], j0:j0 + jb]
    return a, piv


@pytest.mark.parametrize("m,n,nb,ws", [(96, 96, 8, 32), (120, 90, 16, 48), (70, 110, 8, 24), (64, 64, 16, 16)])
def test_slab_left_looking_schedule_is_the_same_factorisation(O, m, n, nb, ws):
    """Schedule invariance behind the planned upload overlap: slab-wise left-looking with deferred left interchanges gives
    the pivots of the right-looking order (and of the oracle) and the same factors up to round-off of the GEMM blocking."""
    a0 = O.matgen64_tile(max(m, n), 77, 0, m, 0, n)
    lu_r, piv_r = _blocked_right_looking(a0, nb)
    lu_s, piv_s = _slab_left_looking(a0, nb, ws)
    assert np.array_equal(piv_r, piv_s)
    assert np.allclose(lu_r, lu_s, rtol=0, atol=1e-12)
    ref = a0.copy(order="F")
    ipr, info = O.getrf(ref, nb)
    assert info == 0 and np.array_equal(ipr - 1, piv_r)
    assert np.allclose(ref, lu_r, rtol=0, atol=1e-12)


def _rl_step_exact(a, piv, j0, jb, c0, c1, panel=None):
    """Block step (j0, jb) applied to columns [c0, c1) with ELEMENTWISE-deterministic arithmetic (rank-1 updates in k order, no
    BLAS): interchanges, U12 = L11^-1 A12, A22 -= L21 U12.  panel: the step's panel in its step-time row order (rows j0..)."""
    pan = a[j0:, j0:j0 + jb] if panel is None else panel
    for j in range(j0, j0 + jb):
        p = piv[j]
        if p != j:
            a[[j, p], c0:c1] = a[[p, j], c0:c1]
    for kk in range(jb):                                                # forward substitution with the unit lower L11, row by row
        a[j0 + kk + 1:j0 + jb, c0:c1] -= np.outer(pan[kk + 1:jb, kk], a[j0 + kk, c0:c1])
    for kk in range(jb):                                                # trailing update, one rank-1 term at a time
        a[j0 + jb:, c0:c1] -= np.outer(pan[jb:, kk], a[j0 + kk, c0:c1])


def _panel_exact(a, piv, j0, jb):
    for j in range(j0, j0 + jb):
        p = j + int(np.argmax(np.abs(a[j:, j])))
        piv[j] = p
        if p != j:
            a[[j, p], j0:j0 + jb] = a[[p, j], j0:j0 + jb]
        if a[j, j] != 0.0:
            a[j + 1:, j] *= 1.0 / a[j, j]
        a[j + 1:, j + 1:j0 + jb] -= np.outer(a[j + 1:, j], a[j, j + 1:j0 + jb])


def _streamed_right_looking(a0, nb, present):
    """The schedule lu.cu runs for a HOST-RESIDENT caller: the right-looking sweep works on the columns [0, Np) that
    have arrived; columns that arrive later JOIN at the top of a step k and are first taken through steps 0 .. k-1 with the KEPT
    copy of each panel (step-time row order: the in-place panel has since been permuted by later left interchanges).
    present(k) = number of columns that have arrived when step k starts."""
    m, n = a0.shape; mn = min(m, n)
    a = np.full_like(a0, np.nan, order="F")
    piv = np.zeros(mn, np.int64); kept = {}
    nsteps = (mn + nb - 1) // nb
    Np = 0
    for k in range(nsteps):
        j0 = k * nb; jb = min(nb, mn - j0)
        want = max(present(k), min(n, j0 + jb + nb))                     # the next panel's columns are waited for
        if k == nsteps - 1:
            want = n
        if want > Np:
            a[:, Np:want] = a0[:, Np:want]
            for kk in range(k):                                          # replay, in step order, with the kept panels
                _rl_step_exact(a, piv, kk * nb, nb, Np, want, panel=kept[kk])
            Np = want
        _panel_exact(a, piv, j0, jb)
        kept[k] = a[j0:, j0:j0 + jb].copy()
        for j in range(j0, j0 + jb):                                     # left interchanges
            p = piv[j]
            if p != j:
                a[[j, p], :j0] = a[[p, j], :j0]
        _rl_step_exact(a, piv, j0, jb, j0 + jb, Np)
    return a, piv


@pytest.mark.parametrize("m,n,nb", [(96, 96, 8), (120, 90, 16), (70, 110, 8)])
@pytest.mark.parametrize("rate", [0, 5, 23, 10 ** 6])
def test_streamed_schedule_is_bit_identical_to_the_resident_sweep(O, m, n, nb, rate):
    """Host-resident callers: whatever the arrival order of the column slabs (rate = columns arriving per step; 0 = only the
    columns that are waited for, 10^6 = everything at once), the streamed schedule with replay performs, for every element, the
    same operations in the same order as the device-resident right-looking sweep -- so with elementwise-deterministic arithmetic
    the factors are BIT-identical (the GPU kernels have that property: one accumulation chain per element in k order)."""
    a0 = O.matgen64_tile(max(m, n), 78, 0, m, 0, n)
    ref, pr = _streamed_right_looking(a0, nb, lambda k: n)             # everything present from the start = the resident sweep
    lu, piv = _streamed_right_looking(a0, nb, lambda k: min(n, nb + rate * (k + 1)))
    assert np.array_equal(piv, pr)
    assert np.array_equal(lu, ref)
    orc = a0.copy(order="F"); ipr, info = O.getrf(orc, nb)
    assert info == 0 and np.array_equal(ipr - 1, pr) and np.allclose(orc, ref, rtol=0, atol=1e-12)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("trans", ["N", "T", "C"])
def test_oracle_getrs_trans_against_numpy(O, cplx, trans):
    """The restated PDGETRS / PZGETRS (SRC/pdgetrs.f:255-284) for every TRANS against an independent dense solve."""
    n, nb = 70, 8
    a0 = (O.pzmatgen if cplx else O.pdmatgen)(n, n, 100); b0 = (O.pzmatgen if cplx else O.pdmatgen)(n, 4, 200)
    ref = a0.copy(order="F"); ipiv, info = O.getrf(ref, nb)
    assert info == 0
    x = b0.copy(order="F"); O.getrs(ref, ipiv, x, trans)
    aop = {"N": a0, "T": a0.T, "C": a0.conj().T}[trans]
    assert np.abs(x - np.linalg.solve(aop, b0)).max() < 1e-11


def test_exact_ties_are_broken_like_pdamax_on_a_grid(O):
    """PBLAS/SRC/pdamax_.c:436-458: the local candidates go up a binary tree towards process row 0 and the receiver keeps its own on a tie,
    so among equal maxima the one on the lowest absolute process row wins (first local index inside a row).  NB = 1, three process rows:
    rows 0, 3, 6 live on process row 0, rows 1, 4, 7 on row 1, rows 2, 5, 8 on row 2."""
    a = np.asfortranarray(np.eye(9) * 0.5 + 0.01 * np.arange(81).reshape(9, 9) / 81.0)
    a[:, 0] = 0.25
    a[[4, 6, 8], 0] = 2.0                                       # equal maxima at global rows 4 (process row 1), 6 (row 0), 8 (row 2)
    try:
        lu = a.copy(order="F"); ip, _ = O.getrf(lu, 1)
        assert ip[0] == 5                                      # one process row: the first global index
        O.tie_grid(3, 0)
        lu = a.copy(order="F"); ip, _ = O.getrf(lu, 1)
        assert ip[0] == 7                                      # row 6 sits on process row 0
        O.tie_grid(3, 1)                                       # the first block row on process row 1: rows 0, 3, 6 -> process row 1; 2, 5, 8 -> row 0
        lu = a.copy(order="F"); ip, _ = O.getrf(lu, 1)
        assert ip[0] == 9
        O.tie_grid(2, 0)                                       # two process rows: 4, 6, 8 all on process row 0 -> the first of them
        lu = a.copy(order="F"); ip, _ = O.getrf(lu, 1)
        assert ip[0] == 5
    finally:
        O.tie_grid(1, 0)
