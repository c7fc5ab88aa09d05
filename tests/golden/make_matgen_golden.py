"""Golden vectors of the reference's test-matrix generator produced by EXECUTING its own Fortran (TESTING/traditional/LIN/pdmatgen.f,
pzmatgen.f, pmatgeninc.f under /root/reference) once per process of emulated NPROW x NPCOL grids with tests/fortran_matgen_runner.py.
Every parity test's input matrix comes from the oracle's closed-form generator; this file is what pins that generator to the
reference.  Writes tests/golden/matgen_reference.npz.  python tests/golden/make_matgen_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import fortran_matgen_runner as R  # noqa: E402

#        m   n  mb  nb  P  Q  seed  iarow iacol  complex
CASES = [(6, 6, 2, 2, 1, 1, 100, 0, 0, 0), (6, 6, 2, 2, 2, 3, 100, 0, 0, 0), (7, 5, 2, 3, 2, 3, 100, 0, 0, 0), (13, 13, 4, 4, 2, 2, 200, 1, 1, 0),
         (9, 11, 3, 2, 3, 2, 12345, 2, 0, 0), (16, 16, 8, 8, 1, 1, 100, 0, 0, 0), (10, 10, 64, 64, 2, 2, 7, 0, 1, 0), (1, 1, 1, 1, 2, 2, 100, 0, 0, 0),
         (24, 3, 5, 2, 4, 2, 200, 3, 0, 0), (3, 24, 2, 5, 2, 4, 300, 0, 3, 0), (32, 32, 4, 4, 2, 4, 2147483647, 0, 0, 0), (20, 20, 3, 3, 4, 2, 1, 0, 0, 0),
         (6, 6, 2, 2, 2, 3, 100, 0, 0, 1), (9, 7, 3, 2, 3, 2, 200, 1, 0, 1), (12, 12, 4, 4, 1, 1, 100, 0, 0, 1), (17, 5, 4, 3, 2, 2, 31, 0, 1, 1)]


#             n nrhs nb nbrhs perturbation of the solution
CHK_CASES = [(6, 1, 2, 1, 1e-6), (13, 3, 4, 2, 1e-6), (20, 5, 8, 3, 1e-3), (9, 2, 16, 2, 1e-6), (16, 4, 4, 4, 0.0), (1, 1, 1, 1, 1e-6)]


def chk_solution(O, n, nrhs, pert):
    """the solution PDLASCHK is given: the exact one, perturbed (pert = 0: whatever rounding leaves)"""
    a, b = O.pdmatgen(n, n, 100).copy(order="F"), O.pdmatgen(n, nrhs, 200).copy(order="F")
    x = np.linalg.solve(a, b)
    return a, b, np.asfortranarray(x * (1.0 + pert * np.cos(np.arange(n))[:, None]))


if __name__ == "__main__":
    it = R.make()
    store = {}
    for i, (m, n, mb, nb, p, q, seed, ir, ic, z) in enumerate(CASES):
        store[f"case{i}"] = np.array([m, n, mb, nb, p, q, seed, ir, ic, z], np.int64)
        store[f"g{i}"] = R.global_(it, m, n, mb, nb, p, q, seed, ir, ic, complex_=bool(z))
        for pr in range(p):
            for pc in range(q):
                store[f"l{i}_{pr}_{pc}"] = R.local(it, m, n, mb, nb, pr, pc, p, q, seed, ir, ic, complex_=bool(z))
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import oracle as O
    itc = R.make_checks()
    for i, (n, nrhs, nb, nbr, pert) in enumerate(CHK_CASES):
        a, b, x = chk_solution(O, n, nrhs, pert)
        store[f"chk{i}"] = np.array([R.pdlaschk(itc, x, n, nrhs, nb, nbr, 100, 200, np.abs(a).sum(axis=1).max())])
    np.savez_compressed(os.path.join(HERE, "matgen_reference.npz"), **store)
    print("wrote", len(CASES), "cases; PXERBLA log:", it.log)
