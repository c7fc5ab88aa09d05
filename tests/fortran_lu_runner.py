"""Runs the reference's OWN Fortran of the LU path -- SRC/pdgetrf.f, pdgetf2.f, pdlaswp.f, pdgetrs.f read from /root/reference -- on a
1 x 1 process grid with the mini interpreter of tests/fortran77_mini.py.  The control flow (blocking, the peeled first block, partial
last blocks, zero-pivot handling, IPIV semantics, the order of the interchanges / solve / update, INFO) is the reference's source text,
executed; only the PBLAS / BLACS leaves it calls (PDAMAX, PDSWAP, PDSCAL, PDGER, PDTRSM, PDGEMM, PDLAPIV, broadcasts) are numpy
stand-ins written from their Purpose blocks (PBLAS/SRC/pdamax_.c etc.).  TEST INFRASTRUCTURE: it pins the oracle's restatement of the
algorithm (oracle/oracle.c) against the reference's own control flow."""
import os

import numpy as np

import fortran77_mini as F


def make(ref_root="/root/reference", real_lapiv=False, extra=()):
    """real_lapiv: SRC/pdlapiv.f + pdlapv2.f are executed too (their PDSWAP calls reach the numpy leaf) instead of the PDLAPIV stand-in"""
    units = [F.parse(open(os.path.join(ref_root, d, f + ".f")).read())
             for d, f in (("TOOLS", "numroc"), ("TOOLS", "indxg2p"), ("TOOLS", "indxg2l"), ("TOOLS", "indxl2g"), ("TOOLS", "iceil"), ("TOOLS", "infog2l"),
                          ("TOOLS", "chk1mat"), ("TOOLS", "descset"), ("SRC", "pdgetrf"), ("SRC", "pdgetf2"), ("SRC", "pdlaswp"), ("SRC", "pdgetrs"))
             + ((("SRC", "pdlapiv"), ("SRC", "pdlapv2")) if real_lapiv else ()) + tuple(extra)]
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

    def sub(env, name, i, j, m, n):
        a = env[name]
        return a[i - 1:i - 1 + m, j - 1:j - 1 + n]

    # PBLAS/SRC/pdamax_.c:404-487: first index of the largest |x|; all zero -> INDX = IX; AMAX = the signed element
    def pdamax(it, env, parts):
        n, ix, jx, incx = ev(it, env, parts, 0), ev(it, env, parts, 4), ev(it, env, parts, 5), ev(it, env, parts, 7)
        assert incx == 1
        x = env[parts[3]][ix - 1:ix - 1 + n, jx - 1]
        if n < 1:
            it.assign(parts[1], env, 0.0); it.assign(parts[2], env, 0); return
        k = int(np.argmax(np.abs(x)))
        it.assign(parts[1], env, float(x[k])); it.assign(parts[2], env, ix + k)

    # PBLAS/SRC/pdswap_.c: rows when INCX = M_ (pdgetf2.f:218, pdlaswp.f:167-182), columns when INCX = 1
    def pdswap(it, env, parts):
        n = ev(it, env, parts, 0)
        ix, jx, incx = ev(it, env, parts, 2), ev(it, env, parts, 3), ev(it, env, parts, 5)
        iy, jy, incy = ev(it, env, parts, 7), ev(it, env, parts, 8), ev(it, env, parts, 10)
        X, Y = env[parts[1]], env[parts[6]]
        if n <= 0:
            return
        if incx == 1 and incy == 1:
            t = X[ix - 1:ix - 1 + n, jx - 1].copy(); X[ix - 1:ix - 1 + n, jx - 1] = Y[iy - 1:iy - 1 + n, jy - 1]; Y[iy - 1:iy - 1 + n, jy - 1] = t
        else:
            t = X[ix - 1, jx - 1:jx - 1 + n].copy(); X[ix - 1, jx - 1:jx - 1 + n] = Y[iy - 1, jy - 1:jy - 1 + n]; Y[iy - 1, jy - 1:jy - 1 + n] = t

    def pdscal(it, env, parts):
        n, alpha, ix, jx, incx = ev(it, env, parts, 0), ev(it, env, parts, 1), ev(it, env, parts, 3), ev(it, env, parts, 4), ev(it, env, parts, 6)
        assert incx == 1
        if n > 0:
            env[parts[2]][ix - 1:ix - 1 + n, jx - 1] *= alpha

    def pdger(it, env, parts):
        m, n, alpha = ev(it, env, parts, 0), ev(it, env, parts, 1), ev(it, env, parts, 2)
        ix, jx, incx = ev(it, env, parts, 4), ev(it, env, parts, 5), ev(it, env, parts, 7)
        iy, jy = ev(it, env, parts, 9), ev(it, env, parts, 10)
        ia, ja = ev(it, env, parts, 14), ev(it, env, parts, 15)
        assert incx == 1
        if m > 0 and n > 0:
            x = env[parts[3]][ix - 1:ix - 1 + m, jx - 1].copy(); y = env[parts[8]][iy - 1, jy - 1:jy - 1 + n].copy()
            sub(env, parts[13], ia, ja, m, n)[...] += alpha * np.outer(x, y)

    def pdtrsm(it, env, parts):
        side, uplo, trans, diag = (ev(it, env, parts, k)[0].upper() for k in range(4))
        m, n, alpha = ev(it, env, parts, 4), ev(it, env, parts, 5), ev(it, env, parts, 6)
        ia, ja, ib, jb = ev(it, env, parts, 8), ev(it, env, parts, 9), ev(it, env, parts, 12), ev(it, env, parts, 13)
        assert side == "L"
        if m <= 0 or n <= 0:
            return
        from scipy.linalg import solve_triangular
        t = sub(env, parts[7], ia, ja, m, m)
        b = sub(env, parts[11], ib, jb, m, n)
        b[...] = solve_triangular(t, alpha * b, lower=(uplo == "L"), trans=(0 if trans == "N" else 1), unit_diagonal=(diag == "U"))

    def pdgemm(it, env, parts):
        ta, tb = ev(it, env, parts, 0)[0].upper(), ev(it, env, parts, 1)[0].upper()
        m, n, k, alpha = (ev(it, env, parts, q) for q in (2, 3, 4, 5))
        ia, ja, ib, jb, beta, ic, jc = (ev(it, env, parts, q) for q in (7, 8, 11, 12, 14, 16, 17))
        assert ta == "N" and tb == "N"
        if m > 0 and n > 0:
            c = sub(env, parts[15], ic, jc, m, n)
            c[...] = alpha * (sub(env, parts[6], ia, ja, m, k) @ sub(env, parts[10], ib, jb, k, n)) + beta * c

    # SRC/pdlapiv.f Purpose: DIREC = 'F': rows i = 1..N swapped with IPIV(i) in order; 'B': in reverse order (ROWCOL = 'R', PIVROC = 'C')
    def pdlapiv(it, env, parts):
        direc, rowcol = ev(it, env, parts, 0)[0].upper(), ev(it, env, parts, 1)[0].upper()
        m, n, ia, ja = (ev(it, env, parts, q) for q in (3, 4, 6, 7))
        ip = ev(it, env, parts, 10)
        assert rowcol == "R"
        A, piv = env[parts[5]], env[parts[9]]
        order = range(m) if direc == "F" else range(m - 1, -1, -1)
        for i in order:
            p = int(piv[ip - 1 + i]) - 1                      # IPIV holds global row indices of A (pdgetrf.f:118-121)
            r = ia - 1 + i
            if p != r:
                A[[r, p], ja - 1:ja - 1 + n] = A[[p, r], ja - 1:ja - 1 + n]

    cbs = {"BLACS_GRIDINFO": gridinfo, "PXERBLA": pxerbla, "BLACS_ABORT": nop, "PB_TOPGET": topget, "PB_TOPSET": nop, "PCHK1MAT": nop, "PCHK2MAT": nop,
           "IGEBS2D": nop, "IGEBR2D": nop, "IGAMN2D": nop, "PDAMAX": pdamax, "PDSWAP": pdswap, "PDSCAL": pdscal, "PDGER": pdger, "PDTRSM": pdtrsm,
           "PDGEMM": pdgemm, "PDLAPIV": pdlapiv}
    if real_lapiv:
        del cbs["PDLAPIV"]
    it = F.Interp(units, cbs)
    it.log = log
    return it


def pdgetrf(it, a, nb, ia=1, ja=1, m=None, n=None):
    """a: global matrix (float64, Fortran order), factored in place by the reference's PDGETRF on a 1 x 1 grid with NB = nb.
    Returns (ipiv as the reference leaves it: LOCr(M_A) + MB_A entries, 1-based, INFO)."""
    M, N = a.shape
    m = M if m is None else m
    n = N if n is None else n
    desc = [1, 0, M, N, nb, nb, 0, 0, max(1, M)]
    ipiv = [0] * (M + nb)
    out = it.call("PDGETRF", m, n, a, ia, ja, desc, ipiv, 0)
    return np.array(ipiv, np.int32), out["INFO"]


def pdgetrs(it, trans, a, ipiv, b, nb):
    n = a.shape[0]
    desca = [1, 0, n, n, nb, nb, 0, 0, max(1, n)]
    descb = [1, 0, n, b.shape[1], nb, nb, 0, 0, max(1, n)]
    out = it.call("PDGETRS", trans, n, b.shape[1], a, 1, 1, desca, list(ipiv), b, 1, 1, descb, 0)
    return out["INFO"]
