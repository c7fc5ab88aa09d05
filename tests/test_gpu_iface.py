"""GPU parity tests of the rest of the drop-in interface on a 1x1 grid: PDGETRS with TRANS = 'T' / 'C'
(SRC/pdgetrs.f:268-284), block-aligned sub-matrix operands IA, JA > 1 and M, N smaller than the descriptor's matrix
(SRC/pdgetrf.f:178-186,219-250), right-hand sides anywhere in B (JB > 1).  The oracle is the checker."""
import numpy as np
import pytest

from tests.helpers import lu_err, PADVAL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,nb,nrhs", [(13, 4, 3), (200, 64, 2), (1000, 128, 1), (1536, 512, 5)])
@pytest.mark.parametrize("trans", ["N", "T", "C"])
def test_pdgetrs_trans(S, O, ctx11, n, nb, nrhs, trans):
    a0 = O.pdmatgen(n, n, 100); b0 = O.pdmatgen(n, nrhs, 200)
    lu = a0.copy(order="F")
    da, _ = S.descinit(n, n, nb, nb, 0, 0, ctx11, n); db, _ = S.descinit(n, nrhs, nb, 2, 0, 0, ctx11, n)
    ipiv = np.zeros(n + nb, np.int32)
    assert S.pdgetrf(n, n, lu, 1, 1, da, ipiv) == 0
    x = b0.copy(order="F")
    assert S.pdgetrs(trans, n, nrhs, lu, 1, 1, da, ipiv, x, 1, 1, db) == 0
    ref = a0.copy(order="F"); ipr, _ = O.getrf(ref, nb)
    xr = b0.copy(order="F"); O.getrs(ref, ipr, xr, trans)
    assert np.abs(x - xr).max() / np.abs(xr).max() < 1e-9
    aop = a0 if trans == "N" else a0.T
    assert O.sresid(np.asfortranarray(aop), x, b0) < 1.0


@pytest.mark.parametrize("n,nb,nrhs", [(60, 8, 3), (500, 64, 2)])
@pytest.mark.parametrize("trans", ["T", "C"])
def test_pzgetrs_trans(S, O, ctx11, n, nb, nrhs, trans):
    a0 = O.pzmatgen(n, n, 100); b0 = O.pzmatgen(n, nrhs, 200)
    lu = a0.copy(order="F")
    da, _ = S.descinit(n, n, nb, nb, 0, 0, ctx11, n); db, _ = S.descinit(n, nrhs, nb, 1, 0, 0, ctx11, n)
    ipiv = np.zeros(n + nb, np.int32)
    assert S.pzgetrf(n, n, lu, 1, 1, da, ipiv) == 0
    x = b0.copy(order="F")
    assert S.pzgetrs(trans, n, nrhs, lu, 1, 1, da, ipiv, x, 1, 1, db) == 0
    aop = a0.T if trans == "T" else a0.conj().T
    xr = np.linalg.solve(aop, b0)
    assert np.abs(x - xr).max() / np.abs(xr).max() < 1e-8
    assert O.sresid(np.asfortranarray(aop), x, b0) < 1.0
    ref = a0.copy(order="F"); ipr, _ = O.getrf(ref, nb)
    xo = b0.copy(order="F"); O.getrs(ref, ipr, xo, trans)
    assert np.abs(x - xo).max() / np.abs(xo).max() < 1e-9


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("mg,ng,nb,ia,ja,m,n", [(40, 40, 4, 9, 5, 20, 24), (300, 260, 32, 65, 33, 200, 200), (1024, 1024, 128, 257, 129, 600, 700),
                                                (500, 500, 64, 1, 1, 300, 200), (500, 500, 64, 129, 1, 371, 300)])
def test_pdgetrf_submatrix(S, O, ctx11, mg, ng, nb, ia, ja, m, n, device):
    """sub(A) = A(IA:IA+M-1, JA:JA+N-1), block-aligned offsets, M and N smaller than the descriptor's matrix: the factors of
    sub(A) equal the oracle's factorisation of that block, IPIV(IA-1+i) holds row indices OF A, nothing outside sub(A) and
    its IPIV entries is written."""
    ag = O.pdmatgen(mg, ng, 100)
    sub0 = np.asfortranarray(ag[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n])
    ref = sub0.copy(order="F"); ipr, infr = O.getrf(ref, nb)
    lld = mg + 2
    al = np.full((lld, ng), PADVAL, order="F"); al[:mg, :] = ag
    desc, info = S.descinit(mg, ng, nb, nb, 0, 0, ctx11, lld)
    assert info == 0
    ipiv = np.full(mg + nb, -77, np.int32)
    if device:
        import torch
        t = torch.from_numpy(np.ascontiguousarray(al.T)).cuda()
        info = S.pdgetrf(m, n, t, ia, ja, desc, ipiv)
        al = np.asfortranarray(t.cpu().numpy().T)
    else:
        info = S.pdgetrf(m, n, al, ia, ja, desc, ipiv)
    assert info == infr
    mn = min(m, n)
    assert np.array_equal(ipiv[ia - 1:ia - 1 + mn], ipr + (ia - 1))
    assert np.all(ipiv[:ia - 1] == -77) and np.all(ipiv[ia - 1 + mn:] == -77)
    got = al[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n]
    assert lu_err(got, ref, sub0) < 1.0
    expect = np.full((lld, ng), PADVAL, order="F"); expect[:mg, :] = ag
    expect[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n] = got
    assert np.array_equal(al, expect), "something outside sub(A) was written"


@pytest.mark.parametrize("trans", ["N", "T"])
def test_pdgetrs_submatrix_and_rhs_window(S, O, ctx11, trans):
    """PDGETRS on sub(A) with the right-hand sides in the middle of a wider B (IB = IA aligned, JB arbitrary)."""
    mg, nb, ia, ja, n = 700, 64, 129, 65, 400
    ag = O.pdmatgen(mg, mg, 100)
    desc, _ = S.descinit(mg, mg, nb, nb, 0, 0, ctx11, mg)
    al = ag.copy(order="F")
    ipiv = np.zeros(mg + nb, np.int32)
    assert S.pdgetrf(n, n, al, ia, ja, desc, ipiv) == 0
    nbg, jb, nrhs, nbb = 11, 4, 5, 3
    bg = O.pdmatgen(mg, nbg, 200)
    bl = bg.copy(order="F")
    descb, _ = S.descinit(mg, nbg, nb, nbb, 0, 0, ctx11, mg)
    assert S.pdgetrs(trans, n, nrhs, al, ia, ja, desc, ipiv, bl, ia, jb, descb) == 0
    sub0 = np.asfortranarray(ag[ia - 1:ia - 1 + n, ja - 1:ja - 1 + n]); b0 = np.asfortranarray(bg[ia - 1:ia - 1 + n, jb - 1:jb - 1 + nrhs])
    x = np.asfortranarray(bl[ia - 1:ia - 1 + n, jb - 1:jb - 1 + nrhs])
    aop = sub0 if trans == "N" else np.asfortranarray(sub0.T)
    assert O.sresid(aop, x, b0) < 1.0
    expect = bg.copy(order="F"); expect[ia - 1:ia - 1 + n, jb - 1:jb - 1 + nrhs] = x
    assert np.array_equal(bl, expect), "PDGETRS wrote outside sub(B)"
    # PDGESV on the same windows gives the same solution
    al2 = ag.copy(order="F"); bl2 = bg.copy(order="F"); ip2 = np.zeros(mg + nb, np.int32)
    if trans == "N":
        assert S.pdgesv(n, nrhs, al2, ia, ja, desc, ip2, bl2, ia, jb, descb) == 0
        assert np.array_equal(bl2, bl) and np.array_equal(ip2, ipiv)
