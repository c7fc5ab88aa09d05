"""Golden vectors of the LU path produced by EXECUTING the reference's own Fortran (SRC/pdgetrf.f, pdgetf2.f, pdlaswp.f, pdgetrs.f under
/root/reference) on a 1 x 1 grid with tests/fortran_lu_runner.py (mini interpreter; numpy stands in for the PBLAS leaves).  Inputs are
PDMATGEN matrices (seed 100 / 200), so only shapes, IPIV, INFO, factors and solutions are stored.  Writes tests/golden/lu_reference.npz.
Run here (the reference tree is not on the GPU boxes):  python tests/golden/make_lu_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(HERE))
import fortran_lu_runner as R  # noqa: E402
import oracle as O  # noqa: E402

# the LU.dat problem sizes and block sizes, a few larger ones, a sub-matrix (the peeled first block of pdgetrf.f:219-250), zero pivots
CASES = [dict(m=m, n=n, nb=nb) for (m, n) in [(4, 4), (10, 12), (17, 13), (13, 13)] for nb in (2, 3, 4)] + \
        [dict(m=64, n=64, nb=8), dict(m=100, n=100, nb=32), dict(m=50, n=50, nb=50), dict(m=30, n=30, nb=64), dict(m=70, n=40, nb=16), dict(m=40, n=70, nb=16),
         dict(m=24, n=24, nb=4, mg=40, ng=40, ia=9, ja=5), dict(m=20, n=28, nb=4, mg=40, ng=40, ia=5, ja=9),
         dict(m=20, n=20, nb=4, zero_col=7), dict(m=16, n=16, nb=8, zero_col=0)]


def run(it, cs):
    m, n, nb = cs["m"], cs["n"], cs["nb"]
    mg, ng, ia, ja = cs.get("mg", m), cs.get("ng", n), cs.get("ia", 1), cs.get("ja", 1)
    a = O.pdmatgen(mg, ng, 100).copy(order="F")
    if cs.get("zero_col") is not None:
        a[:, cs["zero_col"]] = 0.0
    ipiv, info = R.pdgetrf(it, a, nb, ia, ja, m, n)
    out = dict(lu=a, ipiv=ipiv, info=info)
    if m == n and info == 0 and ia == 1 and ja == 1:
        for trans in "NT":
            b = O.pdmatgen(n, 3, 200).copy(order="F")
            R.pdgetrs(it, trans, a, ipiv, b, nb)
            out["x" + trans] = b
    return out


if __name__ == "__main__":
    it = R.make()
    store = {}
    for i, cs in enumerate(CASES):
        o = run(it, cs)
        store[f"case{i}"] = np.array([cs["m"], cs["n"], cs["nb"], cs.get("mg", cs["m"]), cs.get("ng", cs["n"]), cs.get("ia", 1), cs.get("ja", 1),
                                      -1 if cs.get("zero_col") is None else cs["zero_col"], o["info"]], np.int64)
        store[f"lu{i}"] = o["lu"]; store[f"ipiv{i}"] = o["ipiv"]
        for t in "NT":
            if "x" + t in o:
                store[f"x{t}{i}"] = o["x" + t]
    np.savez_compressed(os.path.join(HERE, "lu_reference.npz"), **store)
    print("wrote", len(CASES), "cases; PXERBLA log:", it.log)
