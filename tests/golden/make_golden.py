"""Regenerates tests/golden/golden.npz from the CPU oracle, cross-checked against scipy's LAPACK.

Run in the build container:  python tests/golden/make_golden.py
The *.dat files next to this script are the reference's own data fixtures (EXAMPLE/DSCAEXMAT.dat,
DSCAEXRHS.dat, SCAEX.dat; TESTING/traditional/LU.dat), copied verbatim: they are inputs, not sources.
The reference ships no expected outputs for this path, so the vectors stored here are produced by the
oracle and pinned by an independent implementation (scipy.linalg.lu_factor / solve) at generation time.
"""
import os
import sys

import numpy as np
import scipy.linalg as sla

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402
from tests.helpers import load_example_6x6  # noqa: E402

out = {}
# 1. PDMATGEN closed form, seed 100 (A) and 200 (B): first 8x8 / 8x1 values
out["pdmatgen_8x8_seed100"] = O.pdmatgen(8, 8, 100)
out["pdmatgen_8x1_seed200"] = O.pdmatgen(8, 1, 200)
out["matgen64_6x5_seed42"] = O.matgen64_tile(6, 42, 0, 6, 0, 5)
# 2. the 6x6 tutorial system: pivots, factors and solution
A, B = load_example_6x6(os.path.join(HERE, "DSCAEXMAT.dat"), os.path.join(HERE, "DSCAEXRHS.dat"))
lu = A.copy(order="F")
ipiv, info = O.getrf(lu, 2)
x = B.copy(order="F")
O.getrs(lu, ipiv, x)
lu_s, piv_s = sla.lu_factor(A)
assert info == 0 and np.array_equal(ipiv - 1, piv_s), (ipiv, piv_s)
assert np.allclose(x, np.linalg.solve(A, B), rtol=1e-13)
out["ex6_ipiv"] = ipiv
out["ex6_lu"] = lu
out["ex6_x"] = x
# 3. N=64 NB=8 PDMATGEN case: pivots + factors
a0 = O.pdmatgen(64, 64, 100)
lu = a0.copy(order="F")
ipiv, info = O.getrf(lu, 8)
lu_s, piv_s = sla.lu_factor(a0)
assert info == 0 and np.array_equal(ipiv - 1, piv_s)
assert np.abs(lu - lu_s).max() < 1e-12
out["n64_ipiv"] = ipiv
out["n64_lu"] = lu
np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
print("wrote golden.npz:", {k: v.shape for k, v in out.items()})
