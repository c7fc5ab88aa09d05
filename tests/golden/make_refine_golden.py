"""Golden vectors of the condition estimate and of the iterative refinement produced by EXECUTING the reference's own Fortran
(SRC/pdgecon.f, pdlacon.f, pdlatrs.f, pdgerfs.f, pdgesvx.f, pdgeequ.f, pdlaqge.f, pdlange.f, pdgetri.f, pdtrtri.f, pdtrti2.f under
/root/reference) on a 1 x 1 grid with tests/fortran_refine_runner.py.  Inputs: the oracle's LU factors (block size NB) of
A = PDMATGEN(seed 100), optionally badly scaled; sub-matrix cases factor A(IA:, JA:) in place inside a larger matrix.
Writes tests/golden/refine_reference.npz.  python tests/golden/make_refine_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(HERE))
import fortran_refine_runner as R  # noqa: E402
import oracle as O  # noqa: E402

CASES = [dict(n=n, nb=nb) for n, nb in ((2, 2), (3, 2), (6, 2), (10, 3), (17, 4), (30, 8), (40, 64), (64, 8), (90, 16))] + \
        [dict(n=24, nb=4, scale=3), dict(n=45, nb=4, scale=6), dict(n=33, nb=8, off=8), dict(n=20, nb=4, off=12)]


def matrix(cs):
    n = cs["n"]
    a = O.pdmatgen(n, n, 100).copy(order="F")
    if cs.get("scale"):
        k = np.arange(n)
        a = np.asfortranarray((10.0 ** (cs["scale"] * np.sin(k))[:, None]) * a * (10.0 ** (cs["scale"] * np.cos(2 * k))[None, :]))
    return a


RFS_CASES = [dict(n=6, nb=2, nrhs=1, trans="N"), dict(n=17, nb=4, nrhs=3, trans="N"), dict(n=17, nb=4, nrhs=2, trans="T"), dict(n=30, nb=8, nrhs=9, trans="N"),
             dict(n=24, nb=4, nrhs=2, trans="N", scale=2), dict(n=24, nb=4, nrhs=2, trans="T", scale=2), dict(n=2, nb=2, nrhs=1, trans="N"),
             dict(n=40, nb=64, nrhs=2, trans="N", perturb=1e-3)]


def rfs_inputs(cs):
    """(a, lu, ipiv, b, x0): x0 = the PDGETRS solution, perturbed so that the refinement has something to do"""
    a = matrix(cs)
    n, nb, nrhs = cs["n"], cs["nb"], cs["nrhs"]
    b = O.pdmatgen(n, nrhs, 200).copy(order="F")
    lu = a.copy(order="F"); ipiv, info = O.getrf(lu, nb)
    assert info == 0
    x0 = b.copy(order="F"); O.getrs(lu, ipiv, x0, cs["trans"])
    x0 *= 1.0 + cs.get("perturb", 1e-7)
    return a, lu, ipiv, b, x0


SVX_CASES = [dict(n=6, nb=2, nrhs=1, fact="N", trans="N"), dict(n=17, nb=4, nrhs=2, fact="E", trans="N", scale=5), dict(n=17, nb=4, nrhs=2, fact="E", trans="T", scale=5),
             dict(n=12, nb=4, nrhs=2, fact="E", trans="N"), dict(n=20, nb=8, nrhs=3, fact="N", trans="T"), dict(n=24, nb=4, nrhs=2, fact="E", trans="N", scale=2),
             dict(n=24, nb=4, nrhs=2, fact="E", trans="N", rowscale=6), dict(n=24, nb=4, nrhs=2, fact="E", trans="T", colscale=6),
             dict(n=16, nb=4, nrhs=1, fact="N", trans="N", twin=(2, 3)), dict(n=16, nb=4, nrhs=1, fact="E", trans="N", zero_row=5),
             dict(n=16, nb=4, nrhs=1, fact="N", trans="N", zero_col=7), dict(n=2, nb=2, nrhs=1, fact="E", trans="N"), dict(n=1, nb=2, nrhs=2, fact="N", trans="N")]


def svx_inputs(cs):
    n, nrhs = cs["n"], cs["nrhs"]
    a = matrix(cs)
    k = np.arange(n)
    if cs.get("rowscale"):
        a = np.asfortranarray((10.0 ** (cs["rowscale"] * np.sin(k))[:, None]) * a)
    if cs.get("colscale"):
        a = np.asfortranarray(a * (10.0 ** (cs["colscale"] * np.cos(k))[None, :]))
    if cs.get("twin"):
        a[:, cs["twin"][1]] = a[:, cs["twin"][0]]                 # singular to working precision
    if cs.get("zero_row") is not None:
        a[cs["zero_row"], :] = 0.0
    if cs.get("zero_col") is not None:
        a[:, cs["zero_col"]] = 0.0                                # PDGETRF reports INFO > 0
    return np.asfortranarray(a), O.pdmatgen(n, nrhs, 200).copy(order="F")


TRI_CASES = [dict(n=1, nb=2), dict(n=2, nb=1), dict(n=6, nb=2), dict(n=10, nb=3), dict(n=17, nb=4), dict(n=30, nb=8), dict(n=20, nb=64), dict(n=24, nb=4, scale=2),
             dict(n=16, nb=4, zero=9), dict(n=16, nb=4, zero=0), dict(n=12, nb=4, off=8), dict(n=9, nb=3, off=3, zero=4)]


def tri_inputs(cs):
    """(big, ipiv, lu): the oracle's factors of A placed at (off, off) of a larger matrix; zero: U(zero, zero) set to 0 (singular)"""
    n, nb, off = cs["n"], cs["nb"], cs.get("off", 0)
    a = matrix(cs)
    lu = a.copy(order="F"); ipiv, info = O.getrf(lu, nb)
    if cs.get("zero") is not None:
        lu[cs["zero"], cs["zero"]] = 0.0
    big = O.pdmatgen(n + off, n + off, 55).copy(order="F"); big[off:, off:] = lu
    return big, np.asarray(ipiv, np.int64) + off, lu, a


#              m   n  nb  relative perturbation of the factors (0: FRESID is rounding only)
FCHK_CASES = [(6, 6, 2, 1e-6), (13, 13, 4, 1e-6), (10, 14, 3, 1e-5), (15, 9, 4, 1e-6), (20, 20, 32, 1e-3), (13, 13, 4, 0.0), (10, 14, 3, 0.0)]


def fchk_inputs(m, n, nb, pert):
    a = O.pdmatgen(m, n, 100).copy(order="F")
    lu = a.copy(order="F"); ipiv, info = O.getrf(lu, nb)
    lu *= 1.0 + pert * np.cos(np.arange(m))[:, None] * np.sin(1.0 + np.arange(n))[None, :]
    full_ip = np.arange(1, m + 1); full_ip[:min(m, n)] = ipiv[:min(m, n)]
    return a, np.asfortranarray(lu), ipiv, full_ip


if __name__ == "__main__":
    it = R.make(extra=R.SVX_UNITS + R.TRI_UNITS)
    store = {}
    for i, cs in enumerate(CASES):
        a = matrix(cs)
        n, nb, off = cs["n"], cs["nb"], cs.get("off", 0)
        lu = a.copy(order="F"); ipiv, info = O.getrf(lu, nb)
        assert info == 0
        big = O.pdmatgen(n + off, n + off, 55).copy(order="F"); big[off:, off:] = lu
        vals = []
        for norm in "1I":
            anorm = np.abs(a).sum(axis=0).max() if norm == "1" else np.abs(a).sum(axis=1).max()
            rc, info = R.pdgecon(it, norm, big, anorm, nb, ia=off + 1, ja=off + 1, n=n)
            assert info == 0
            vals.append(rc)
        store[f"case{i}"] = np.array([n, nb, cs.get("scale", 0), off], np.int64)
        store[f"rcond{i}"] = np.array(vals)
    for i, cs in enumerate(RFS_CASES):
        a, lu, ipiv, b, x = rfs_inputs(cs)
        ferr, berr, info = R.pdgerfs(it, cs["trans"], a, lu, ipiv, b, x, cs["nb"])
        assert info == 0
        store[f"rfs{i}"] = np.array([cs["n"], cs["nb"], cs["nrhs"], ord(cs["trans"]), cs.get("scale", 0)], np.int64)
        store[f"rfs_pert{i}"] = np.array([cs.get("perturb", 1e-7)])
        store[f"rfs_x{i}"], store[f"rfs_ferr{i}"], store[f"rfs_berr{i}"] = x, ferr, berr
    for i, cs in enumerate(SVX_CASES):
        a, b = svx_inputs(cs)
        n = cs["n"]
        af, ipiv, r, c = np.zeros((n, n), order="F"), np.zeros(n, np.int64), np.zeros(n), np.zeros(n)
        res = R.pdgesvx(it, cs["fact"], cs["trans"], a, af, ipiv, "N", r, c, b, cs["nb"])
        store[f"svx{i}"] = np.array([res["info"], ord(res["equed"][0])], np.int64)
        store[f"svx_rcond{i}"] = np.array([res["rcond"]])
        store[f"svx_work{i}"] = np.array([res.get("lwork", 0), res.get("liwork", 0)], np.int64)
        for key, val in (("a", a), ("af", af), ("ipiv", ipiv), ("r", r), ("c", c), ("b", b)):
            store[f"svx_{key}{i}"] = val
        if "x" in res:
            store[f"svx_x{i}"], store[f"svx_ferr{i}"], store[f"svx_berr{i}"] = res["x"], res["ferr"], res["berr"]
        # FACT = 'F' on what the first call returned (only where it completed)
        if res["info"] == 0:
            a2, b2 = svx_inputs(cs)
            eq = res["equed"][0]
            if eq in "RB":
                a2 = np.asfortranarray(r[:, None] * a2)
            if eq in "CB":
                a2 = np.asfortranarray(a2 * c[None, :])
            resf = R.pdgesvx(it, "F", cs["trans"], a2, af.copy(order="F"), ipiv.copy(), eq, r.copy(), c.copy(), b2, cs["nb"])
            store[f"svxF{i}"] = np.array([resf["info"], ord(resf["equed"][0])], np.int64)
            store[f"svxF_rcond{i}"] = np.array([resf["rcond"]])
            store[f"svxF_x{i}"] = resf["x"]
    for i, cs in enumerate(TRI_CASES):
        big, ipiv, lu, a = tri_inputs(cs)
        off = cs.get("off", 0)
        info, lw, liw = R.pdgetri(it, big, ipiv, cs["nb"], ia=off + 1, ja=off + 1, n=cs["n"])
        store[f"tri{i}"] = np.array([info, lw, liw], np.int64)
        store[f"tri_inv{i}"] = big
    itf = R.make(extra=R.CHK_UNITS, matgen=True)
    for i, (m, n, nb, pert) in enumerate(FCHK_CASES):
        a, lu, ipiv, full_ip = fchk_inputs(m, n, nb, pert)
        store[f"fchk{i}"] = np.array(R.fresid(itf, lu, full_ip, nb, 100))          # (FRESID, ANORM)
    # argument errors and quick returns as the executed source reports them: (NORM, N, ANORM, LWORK) -> (INFO, RCOND)
    lu = O.pdmatgen(8, 8, 100).copy(order="F")
    a = lu.reshape(-1, order="F").copy()
    desc = [1, 0, 8, 8, 4, 4, 0, 0, 8]
    quick = []
    for norm, n, anorm, lwork in (("X", 8, 1.0, 100), ("1", 8, -1.0, 100), ("1", 8, 1.0, 1), ("1", 0, 1.0, 100), ("1", 8, 0.0, 100), ("I", 1, 2.0, 100)):
        out = it.call("PDGECON", norm, n, a, 1, 1, desc, anorm, -7.0, np.zeros(128), lwork, np.zeros(128, np.int64), 100, 0)
        quick.append([ord(norm), n, anorm, lwork, out["INFO"], out["RCOND"]])
    store["quick"] = np.array(quick)
    np.savez_compressed(os.path.join(HERE, "refine_reference.npz"), **store)
    print("wrote", len(CASES), "+", len(RFS_CASES), "+", len(SVX_CASES), "+", len(TRI_CASES), "cases; PXERBLA log:", it.log)
    print("PDGETRI (info, lwmin, liwmin):", [tuple(int(v) for v in store[f"tri{i}"]) for i in range(len(TRI_CASES))])
    print("PDGESVX (info, equed):", [(int(store[f"svx{i}"][0]), chr(int(store[f"svx{i}"][1]))) for i in range(len(SVX_CASES))])
    print(store["quick"])
