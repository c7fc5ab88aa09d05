"""Shared helpers of the parity tests (mirrors the flow of TESTING/traditional/LIN/pdludriver.f)."""
import numpy as np

EPS = 2.0 ** -53
PADVAL = -9923.0          # guard-zone value of the reference driver (pdludriver.f:79)


def lu_err(lu_test, lu_ref, a0):
    """LU-factor tolerance: max|dLU| / (||A||_inf N eps); must be < 1 with identical IPIV."""
    n = max(a0.shape)
    anorm = np.abs(a0).sum(axis=1).max()
    return float(np.abs(lu_test - lu_ref).max() / (anorm * n * EPS))


def first_mismatch(a, b):
    idx = np.nonzero(np.asarray(a) != np.asarray(b))[0]
    return None if idx.size == 0 else int(idx[0])


def load_example_6x6(path_mat, path_rhs):
    """EXAMPLE/DSCAEXMAT.dat / DSCAEXRHS.dat reader (column-major element list, TOOLS/pdlaread.f:96-107)."""
    def rd(p):
        toks = open(p).read().split()
        m, n = int(toks[0]), int(toks[1])
        v = np.array([float(t.replace("D", "e")) for t in toks[2:2 + m * n]])
        return v.reshape((m, n), order="F")
    return rd(path_mat), rd(path_rhs)
