"""Runs the reference's OWN Fortran of the Cholesky path -- SRC/pdpotrf.f, pdpotf2.f, pdpotrs.f read from /root/reference -- on a 1 x 1
grid with tests/fortran77_mini.py.  The blocked algorithm and the unblocked diagonal-block factorisation (with its DDOT / DGEMV / DSCAL
calls on the local array, addressed A( IOFFA ) like in the source) are the reference's statements, executed; the BLAS / PBLAS leaves
(DDOT, DGEMV, DSCAL, PDTRSM, PDSYRK) are numpy stand-ins written from their Purpose blocks.  TEST INFRASTRUCTURE."""
import os

import numpy as np

import fortran77_mini as F


def make(ref_root="/root/reference", extra=()):
    units = [F.parse(open(os.path.join(ref_root, d, f + ".f")).read())
             for d, f in (("TOOLS", "numroc"), ("TOOLS", "indxg2p"), ("TOOLS", "iceil"), ("TOOLS", "infog2l"), ("TOOLS", "chk1mat"),
                          ("SRC", "pdpotrf"), ("SRC", "pdpotf2"), ("SRC", "pdpotrs")) + tuple(extra)]
    log = []

    def ev(it, env, parts, k):
        return it.eval(parts[k], env)

    def gridinfo(it, env, parts):
        for name, v in zip(parts[1:], (1, 1, 0, 0)):
            env[name] = v

    def nop(it, env, parts):
        pass

    def topget(it, env, parts):
        it.assign(parts[3], env, " ")

    def pxerbla(it, env, parts):
        log.append(("PXERBLA", it.eval(parts[1], env), it.eval(parts[2], env)))

    def mat(env, name, desc_name):
        """the flat local array as its lld x n matrix (a view: writes go to the same memory)"""
        a = env[name]; lld = env[desc_name][8]
        return a.reshape((lld, a.size // lld), order="F")

    def strided(arr, off, n, inc):
        return arr[off:off + (n - 1) * inc + 1:inc] if n > 0 else arr[off:off]

    # ---- local BLAS on the flat array, arguments by address (BLAS level 1 / 2 semantics) ----
    def ddot(it, env, parts):
        n = it.eval(parts[0], env)
        x, ox = it.address(parts[1], env); incx = it.eval(parts[2], env)
        y, oy = it.address(parts[3], env); incy = it.eval(parts[4], env)
        return float(strided(x, ox, n, incx) @ strided(y, oy, n, incy)) if n > 0 else 0.0

    def dscal(it, env, parts):
        n, alpha = ev(it, env, parts, 0), ev(it, env, parts, 1)
        x, ox = it.address(parts[2], env); incx = ev(it, env, parts, 3)
        if n > 0:
            strided(x, ox, n, incx)[...] *= alpha

    def dgemv(it, env, parts):
        trans = ev(it, env, parts, 0)[0].upper()
        m, n, alpha = ev(it, env, parts, 1), ev(it, env, parts, 2), ev(it, env, parts, 3)
        a, oa = it.address(parts[4], env); lda = ev(it, env, parts, 5)
        x, ox = it.address(parts[6], env); incx = ev(it, env, parts, 7)
        beta = ev(it, env, parts, 8)
        y, oy = it.address(parts[9], env); incy = ev(it, env, parts, 10)
        ny, nx = (m, n) if trans == "N" else (n, m)
        if ny <= 0:
            return
        yv = strided(y, oy, ny, incy)
        if m > 0 and n > 0:
            A = np.array([[a[oa + i + j * lda] for j in range(n)] for i in range(m)])
            xv = strided(x, ox, nx, incx).copy()
            yv[...] = alpha * ((A @ xv) if trans == "N" else (A.T @ xv)) + beta * yv
        else:
            yv[...] = beta * yv

    # ---- PBLAS on the distributed (here: whole) matrix ----
    def pdtrsm(it, env, parts):
        side, uplo, trans, diag = (ev(it, env, parts, k)[0].upper() for k in range(4))
        m, n, alpha = ev(it, env, parts, 4), ev(it, env, parts, 5), ev(it, env, parts, 6)
        ia, ja, ib, jb = ev(it, env, parts, 8), ev(it, env, parts, 9), ev(it, env, parts, 12), ev(it, env, parts, 13)
        if m <= 0 or n <= 0:
            return
        from scipy.linalg import solve_triangular
        na = m if side == "L" else n
        t = mat(env, parts[7], parts[10])[ia - 1:ia - 1 + na, ja - 1:ja - 1 + na]
        b = mat(env, parts[11], parts[14])[ib - 1:ib - 1 + m, jb - 1:jb - 1 + n]
        tt = 0 if trans == "N" else 1
        if side == "L":
            b[...] = solve_triangular(t, alpha * b, lower=(uplo == "L"), trans=tt, unit_diagonal=(diag == "U"))
        else:                                     # X op(T) = alpha B  <=>  op(T)' X' = alpha B'
            b[...] = solve_triangular(t, alpha * b.T, lower=(uplo == "L"), trans=1 - tt, unit_diagonal=(diag == "U")).T

    def pdsyrk(it, env, parts):
        uplo, trans = ev(it, env, parts, 0)[0].upper(), ev(it, env, parts, 1)[0].upper()
        n, k, alpha = ev(it, env, parts, 2), ev(it, env, parts, 3), ev(it, env, parts, 4)
        ia, ja, beta, ic, jc = ev(it, env, parts, 6), ev(it, env, parts, 7), ev(it, env, parts, 9), ev(it, env, parts, 11), ev(it, env, parts, 12)
        if n <= 0:
            return
        A = mat(env, parts[5], parts[8])
        a = A[ia - 1:ia - 1 + n, ja - 1:ja - 1 + k] if trans == "N" else A[ia - 1:ia - 1 + k, ja - 1:ja - 1 + n]
        upd = alpha * (a @ a.T if trans == "N" else a.T @ a)
        c = mat(env, parts[10], parts[13])[ic - 1:ic - 1 + n, jc - 1:jc - 1 + n]
        tri = np.tril(np.ones((n, n), bool)) if uplo == "L" else np.triu(np.ones((n, n), bool))
        c[tri] = (upd + beta * c)[tri]             # only the UPLO triangle is referenced

    cbs = {"BLACS_GRIDINFO": gridinfo, "PXERBLA": pxerbla, "BLACS_ABORT": nop, "PB_TOPGET": topget, "PB_TOPSET": nop, "PCHK1MAT": nop, "PCHK2MAT": nop,
           "IGEBS2D": nop, "IGEBR2D": nop, "DSCAL": dscal, "DGEMV": dgemv, "PDTRSM": pdtrsm, "PDSYRK": pdsyrk}
    it = F.Interp(units, cbs)
    it.raw_functions = {"DDOT": ddot}
    it.log = log
    return it


def pdpotrf(it, uplo, a, nb):
    """a: global symmetric matrix (Fortran order), the UPLO triangle factored in place by the reference's PDPOTRF.  Returns INFO."""
    n = a.shape[0]
    flat = a.reshape(-1, order="F")                # a view of the same memory (a is Fortran-contiguous)
    assert np.shares_memory(flat, a)
    desc = [1, 0, n, n, nb, nb, 0, 0, n]
    return it.call("PDPOTRF", uplo, n, flat, 1, 1, desc, 0)["INFO"]


def pdpotrs(it, uplo, a, b, nb):
    n, nrhs = b.shape
    fa, fb = a.reshape(-1, order="F"), b.reshape(-1, order="F")
    assert np.shares_memory(fb, b)
    desca = [1, 0, n, n, nb, nb, 0, 0, n]; descb = [1, 0, n, nrhs, nb, nb, 0, 0, n]
    return it.call("PDPOTRS", uplo, n, nrhs, fa, 1, 1, desca, fb, 1, 1, descb, 0)["INFO"]
