"""Golden vectors of the Cholesky path produced by EXECUTING the reference's own Fortran (SRC/pdpotrf.f, pdpotf2.f, pdpotrs.f under
/root/reference) on a 1 x 1 grid with tests/fortran_chol_runner.py.  Inputs: A = G + G' + 2 n I with G = PDMATGEN(seed 100) (optionally
one diagonal entry set to -1), B = PDMATGEN(seed 200).  Writes tests/golden/chol_reference.npz.  python tests/golden/make_chol_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(HERE))
import fortran_chol_runner as R  # noqa: E402
import oracle as O  # noqa: E402

CASES = [dict(n=n, nb=nb, uplo=u) for n in (4, 10, 17, 13) for nb in (2, 3, 4) for u in "LU"] + \
        [dict(n=64, nb=8, uplo=u) for u in "LU"] + [dict(n=50, nb=64, uplo=u) for u in "LU"] + [dict(n=90, nb=40, uplo=u) for u in "LU"] + \
        [dict(n=30, nb=8, uplo=u, notpd=k) for u in "LU" for k in (0, 13, 29)]


def matrix(n, notpd=None):
    g = O.pdmatgen(n, n, 100)
    a = np.asfortranarray(g + g.T + 2.0 * n * np.eye(n))
    if notpd is not None:
        a[notpd, notpd] = -1.0
    return a


if __name__ == "__main__":
    it = R.make()
    store = {}
    for i, cs in enumerate(CASES):
        a = matrix(cs["n"], cs.get("notpd"))
        info = R.pdpotrf(it, cs["uplo"], a, cs["nb"])
        store[f"case{i}"] = np.array([cs["n"], cs["nb"], ord(cs["uplo"]), -1 if cs.get("notpd") is None else cs["notpd"], info], np.int64)
        store[f"f{i}"] = a
        if info == 0:
            x = O.pdmatgen(cs["n"], 3, 200).copy(order="F")
            R.pdpotrs(it, cs["uplo"], a, x, cs["nb"])
            store[f"x{i}"] = x
    np.savez_compressed(os.path.join(HERE, "chol_reference.npz"), **store)
    print("wrote", len(CASES), "cases; PXERBLA log:", it.log)
