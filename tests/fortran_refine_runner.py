"""Runs the reference's OWN Fortran of the condition estimate -- SRC/pdgecon.f, pdlacon.f (the reverse-communication estimator with its
SAVEd state and computed GO TO), pdlatrs.f, pdrscl.f, read from /root/reference -- on a 1 x 1 process grid with the mini interpreter of
tests/fortran77_mini.py.  The iteration (start vector, sign vectors, the ITMAX = 5 loop and its stopping tests, the alternating-sign
safeguard, which of the two solves runs for KASE = 1 / 2 and ONENRM, RCOND = (1 / AINVNM) / ANORM) is the reference's source text,
executed; the PBLAS leaves it calls (PDTRSV, PDASUM, PDAMAX, PDELGET, DCOPY, PDSCAL) are numpy / scipy stand-ins written from their
Purpose blocks.  TEST INFRASTRUCTURE: pins oracle/oracle_next.c's restatement of PDGECON / PDLACON against the reference's own control flow
(it was pinned to LAPACK's DGECON before, a different implementation of the same estimator)."""
import os

import numpy as np

import fortran77_mini as F

SAFMIN = float(np.finfo(np.float64).tiny)
EPS = 2.0 ** -53


def _pdlamch(ictxt, cmach):
    c = str(cmach)[:1].upper()      # PDLAMCH = DLAMCH combined over the grid (TOOLS/pdlamch? -> LAPACK DLAMCH): 'S' safe minimum, 'E' eps, 'P' eps * base
    return {"S": SAFMIN, "E": EPS, "P": 2.0 * EPS, "B": 2.0, "O": float(np.finfo(np.float64).max), "U": SAFMIN}[c]


def make(ref_root="/root/reference", extra=(), matgen=False):
    """matgen: also the test-matrix generator (TESTING/traditional/LIN/pdmatgen.f + pmatgeninc.f, 32-bit INTEGER wrap-around)"""
    units = [F.parse(open(os.path.join(ref_root, d, f + ".f")).read())
             for d, f in (("TOOLS", "numroc"), ("TOOLS", "indxg2p"), ("TOOLS", "indxg2l"), ("TOOLS", "indxl2g"), ("TOOLS", "iceil"), ("TOOLS", "infog2l"),
                          ("TOOLS", "chk1mat"), ("TOOLS", "descset"), ("SRC", "pdgecon"), ("SRC", "pdlacon"), ("SRC", "pdlatrs"), ("SRC", "pdrscl")) + tuple(extra)]
    if matgen:
        for f in ("pmatgeninc.f", "pdmatgen.f"):
            units += F.parse_file(open(os.path.join(ref_root, "TESTING", "traditional", "LIN", f)).read())
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

    def dmatadd(it, env, parts):                                                        # TESTING/.../pdmatgen? dmatadd: C := alpha A + beta C (by address)
        m, n, alpha, beta = ev(it, env, parts, 0), ev(it, env, parts, 1), ev(it, env, parts, 2), ev(it, env, parts, 5)
        a, ao = it.address(parts[3], env); c, co = it.address(parts[6], env)
        lda, ldc = ev(it, env, parts, 4), ev(it, env, parts, 7)
        for j in range(n):
            c[co + j * ldc:co + j * ldc + m] = alpha * np.array(a[ao + j * lda:ao + j * lda + m]) + beta * np.array(c[co + j * ldc:co + j * ldc + m])

    def vec(it, env, parts, kx, n):
        """the N entries of the distributed vector X(IX:IX+N-1, JX) given as (X, IX, JX, DESCX) from argument kx on: a numpy view"""
        arr, off = it.address(parts[kx], env)
        ix, jx, desc = ev(it, env, parts, kx + 1), ev(it, env, parts, kx + 2), env[parts[kx + 3]]
        lld = desc[8]
        o = off + (ix - 1) + (jx - 1) * lld
        if len(parts) > kx + 4 and ev(it, env, parts, kx + 4) == desc[2]:             # INCX = M_X: a ROW vector X(IX, JX:JX+N-1)
            return arr[o:o + (n - 1) * lld + 1:lld] if n > 0 else arr[o:o], jx
        return arr[o:o + n], ix

    # PBLAS/SRC/pdtrsv_.c: sub(X) := inv(op(sub(A))) sub(X), sub(A) = A(IA:IA+N-1, JA:JA+N-1) triangular
    def pdtrsv(it, env, parts):
        from scipy.linalg import solve_triangular
        uplo, trans, diag = (ev(it, env, parts, k)[0].upper() for k in range(3))
        n = ev(it, env, parts, 3)
        a, aoff = it.address(parts[4], env)
        ia, ja, desca = ev(it, env, parts, 5), ev(it, env, parts, 6), env[parts[7]]
        lld = desca[8]
        x, _ = vec(it, env, parts, 8, n)
        if n <= 0:
            return
        ncols = ja - 1 + n
        full = np.asarray(a[aoff:aoff + lld * ncols]).reshape((ncols, lld)).T       # column-major local array
        t = full[ia - 1:ia - 1 + n, ja - 1:ja - 1 + n]
        x[:] = solve_triangular(t, np.array(x), lower=(uplo == "L"), trans=(0 if trans == "N" else 1), unit_diagonal=(diag == "U"))

    def pdasum(it, env, parts):                                                     # PBLAS/SRC/pdasum_.c: sum of |x_i|
        x, _ = vec(it, env, parts, 2, ev(it, env, parts, 0))
        it.assign(parts[1], env, float(np.abs(x).sum()))

    def pdamax(it, env, parts):                                                     # PBLAS/SRC/pdamax_.c: first largest |x_i|, GLOBAL index
        n = ev(it, env, parts, 0)
        x, ix = vec(it, env, parts, 3, n)
        k = int(np.argmax(np.abs(x))) if n > 0 else 0
        it.assign(parts[1], env, float(x[k]) if n > 0 else 0.0); it.assign(parts[2], env, ix + k if n > 0 else 0)

    def pdelget(it, env, parts):                                                    # TOOLS/pdelget.f: ALPHA = A(IA, JA)
        arr, off = it.address(parts[3], env)
        ia, ja, desc = ev(it, env, parts, 4), ev(it, env, parts, 5), env[parts[6]]
        it.assign(parts[2], env, float(arr[off + (ia - 1) + (ja - 1) * desc[8]]))

    def dcopy(it, env, parts):
        n = ev(it, env, parts, 0)
        x, xo = it.address(parts[1], env); y, yo = it.address(parts[3], env)
        assert ev(it, env, parts, 2) == 1 and ev(it, env, parts, 4) == 1
        if n > 0:
            y[yo:yo + n] = np.array(x[xo:xo + n])

    def pdscal(it, env, parts):
        n, alpha = ev(it, env, parts, 0), ev(it, env, parts, 1)
        x, _ = vec(it, env, parts, 2, n)
        x[:] = alpha * np.array(x)

    def window(it, env, parts, k, m, n):
        """the m x n sub-matrix (A, IA, JA, DESCA) given from argument k on, as a numpy view of the flat column-major local array"""
        a, aoff = it.address(parts[k], env)
        ia, ja, desc = ev(it, env, parts, k + 1), ev(it, env, parts, k + 2), env[parts[k + 3]]
        lld = desc[8]
        ncols = ja - 1 + n
        full = np.asarray(a[aoff:aoff + lld * ncols]).reshape((ncols, lld)).T
        return full[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n]

    # PBLAS/SRC/pdgemv_.c: sub(Y) := alpha op(sub(A)) sub(X) + beta sub(Y);  pdagemv_.c: |alpha| |op(sub(A))| |sub(X)| + |beta sub(Y)|
    def pdgemv(it, env, parts, absolute=False):
        trans = ev(it, env, parts, 0)[0].upper()
        m, n, alpha, beta = ev(it, env, parts, 1), ev(it, env, parts, 2), ev(it, env, parts, 3), ev(it, env, parts, 13)
        a = window(it, env, parts, 4, m, n)
        lx, ly = (n, m) if trans == "N" else (m, n)
        x, _ = vec(it, env, parts, 8, lx)
        y, _ = vec(it, env, parts, 14, ly)
        assert ev(it, env, parts, 12) == 1 and ev(it, env, parts, 18) == 1
        op = a if trans == "N" else a.T
        if absolute:
            y[:] = abs(alpha) * (np.abs(op) @ np.abs(np.array(x))) + np.abs(beta * np.array(y))
        else:
            y[:] = alpha * (op @ np.array(x)) + beta * np.array(y)

    def pdagemv(it, env, parts):
        pdgemv(it, env, parts, absolute=True)

    def pdcopy(it, env, parts):
        n = ev(it, env, parts, 0)
        x, _ = vec(it, env, parts, 1, n); y, _ = vec(it, env, parts, 6, n)
        y[:] = np.array(x)

    def pdaxpy(it, env, parts):
        n, alpha = ev(it, env, parts, 0), ev(it, env, parts, 1)
        x, _ = vec(it, env, parts, 2, n); y, _ = vec(it, env, parts, 7, n)
        y[:] = np.array(y) + alpha * np.array(x)

    # SRC/pdgetrs.f (itself executed by tests/fortran_lu_runner.py): one right-hand side, the factors and IPIV of PDGETRF
    def pdgetrs(it, env, parts):
        from scipy.linalg import solve_triangular
        trans, n, nrhs = ev(it, env, parts, 0)[0].upper(), ev(it, env, parts, 1), ev(it, env, parts, 2)
        lu = window(it, env, parts, 3, n, n)
        iaf = ev(it, env, parts, 4)
        piv = [int(env[parts[7]][iaf - 1 + i]) - iaf for i in range(n)]                 # 0-based, relative to sub(A)
        xs = window(it, env, parts, 8, n, nrhs)
        for k in range(nrhs):
            v = np.array(xs[:, k])
            if trans == "N":
                for i in range(n):
                    v[[i, piv[i]]] = v[[piv[i], i]]
                v = solve_triangular(lu, solve_triangular(lu, v, lower=True, unit_diagonal=True))
            else:
                v = solve_triangular(lu, solve_triangular(lu, v, trans=1), lower=True, unit_diagonal=True, trans=1)
                for i in range(n - 1, -1, -1):
                    v[[i, piv[i]]] = v[[piv[i], i]]
            xs[:, k] = v
        it.assign(parts[12], env, 0)

    # SRC/pdgetrf.f (itself executed by tests/fortran_lu_runner.py, which pins oracle.getrf): the oracle's factorisation of the window
    def pdgetrf(it, env, parts):
        import oracle as O
        m, n = ev(it, env, parts, 0), ev(it, env, parts, 1)
        a = window(it, env, parts, 2, m, n)
        ia, desc = ev(it, env, parts, 3), env[parts[5]]
        lu = np.asfortranarray(np.array(a))
        ipiv, info = O.getrf(lu, desc[5])
        a[:, :] = lu
        for i in range(min(m, n)):
            env[parts[6]][ia - 1 + i] = int(ipiv[i]) + ia - 1
        it.assign(parts[7], env, int(info))

    def pdlacpy(it, env, parts):                                                        # SRC/pdlacpy.f, UPLO = 'Full'
        assert ev(it, env, parts, 0)[0].upper() not in "UL"
        m, n = ev(it, env, parts, 1), ev(it, env, parts, 2)
        window(it, env, parts, 7, m, n)[:, :] = np.array(window(it, env, parts, 3, m, n))

    def dlassq(it, env, parts):                                                         # LAPACK DLASSQ: (scale, sumsq) updated with x
        n = ev(it, env, parts, 0)
        x, xo = it.address(parts[1], env)
        assert ev(it, env, parts, 2) == 1
        scale, sumsq = ev(it, env, parts, 3), ev(it, env, parts, 4)
        for v in np.abs(np.array(x[xo:xo + n])):
            if v != 0.0:
                if scale < v:
                    sumsq = 1.0 + sumsq * (scale / v) ** 2; scale = float(v)
                else:
                    sumsq += (float(v) / scale) ** 2
        it.assign(parts[3], env, float(scale)); it.assign(parts[4], env, float(sumsq))

    def dcombssq(it, env, parts):                                                       # LAPACK DCOMBSSQ: V1 := V1 (+) V2 on (scale, sumsq) pairs
        v1, o1 = it.address(parts[0], env); v2, o2 = it.address(parts[1], env)
        if v1[o1] >= v2[o2]:
            if v1[o1] != 0.0:
                v1[o1 + 1] = v1[o1 + 1] + (v2[o2] / v1[o1]) ** 2 * v2[o2 + 1]
        else:
            v1[o1 + 1] = v2[o2 + 1] + (v1[o1] / v2[o2]) ** 2 * v1[o1 + 1]
            v1[o1] = v2[o2]

    # ---- leaves of SRC/pdgetri.f / pdtrtri.f / pdtrti2.f ----
    def tri_of(t, uplo, diag):
        t = np.tril(np.array(t)) if uplo == "L" else np.triu(np.array(t))
        if diag == "U":
            np.fill_diagonal(t, 1.0)
        return t

    def pdtrmm(it, env, parts):                                                         # PBLAS/SRC/pdtrmm_.c, SIDE = 'L': B := alpha op(A) B
        side, uplo, trans, diag = (ev(it, env, parts, k)[0].upper() for k in range(4))
        m, n, alpha = ev(it, env, parts, 4), ev(it, env, parts, 5), ev(it, env, parts, 6)
        assert side == "L" and trans == "N"
        if m > 0 and n > 0:
            t = tri_of(window(it, env, parts, 7, m, m), uplo, diag)
            b = window(it, env, parts, 11, m, n)
            b[:, :] = alpha * (t @ np.array(b))

    def pdtrsm(it, env, parts):                                                         # PBLAS/SRC/pdtrsm_.c, SIDE = 'R': B := alpha B inv(A)
        from scipy.linalg import solve_triangular
        side, uplo, trans, diag = (ev(it, env, parts, k)[0].upper() for k in range(4))
        m, n, alpha = ev(it, env, parts, 4), ev(it, env, parts, 5), ev(it, env, parts, 6)
        assert trans == "N"
        if m > 0 and n > 0 and side == "R":
            t = window(it, env, parts, 7, n, n)
            b = window(it, env, parts, 11, m, n)
            # X A = alpha B  <=>  A' X' = alpha B'
            b[:, :] = solve_triangular(np.array(t), alpha * np.array(b).T, lower=(uplo == "L"), trans=1, unit_diagonal=(diag == "U")).T
        elif m > 0 and n > 0:
            t = window(it, env, parts, 7, m, m)
            b = window(it, env, parts, 11, m, n)
            b[:, :] = solve_triangular(np.array(t), alpha * np.array(b), lower=(uplo == "L"), unit_diagonal=(diag == "U"))

    # callees of SRC/pdgetrf.f when that file is executed here (flat local arrays): the panel factorisation and the interchanges
    def pdgetf2(it, env, parts):                                                        # SRC/pdgetf2.f (executed by tests/fortran_lu_runner.py)
        import oracle as O
        m, n = ev(it, env, parts, 0), ev(it, env, parts, 1)
        a = window(it, env, parts, 2, m, n)
        ia = ev(it, env, parts, 3)
        lu = np.asfortranarray(np.array(a))
        ipiv, info = O.getrf(lu, max(n, 1))                                              # one block = the unblocked algorithm
        a[:, :] = lu
        for i in range(min(m, n)):
            env[parts[6]][ia - 1 + i] = int(ipiv[i]) + ia - 1
        it.assign(parts[7], env, int(info))

    def pdlaswp(it, env, parts):                                                        # SRC/pdlaswp.f: rows K1..K2 of the N columns from JA, forward
        direc, rowcol = ev(it, env, parts, 0)[0].upper(), ev(it, env, parts, 1)[0].upper()
        n, ja, k1, k2 = ev(it, env, parts, 2), ev(it, env, parts, 5), ev(it, env, parts, 7), ev(it, env, parts, 8)
        assert direc == "F" and rowcol == "R"
        if n <= 0:
            return
        a, aoff = it.address(parts[3], env)
        desc = env[parts[6]]
        lld = desc[8]
        full = np.asarray(a[aoff:aoff + lld * (ja - 1 + n)]).reshape((ja - 1 + n, lld)).T
        for i in range(k1, k2 + 1):
            p_ = int(env[parts[9]][i - 1])
            if p_ != i:
                full[[i - 1, p_ - 1], ja - 1:ja - 1 + n] = full[[p_ - 1, i - 1], ja - 1:ja - 1 + n]

    def pdgemm(it, env, parts):
        ta, tb = ev(it, env, parts, 0)[0].upper(), ev(it, env, parts, 1)[0].upper()
        m, n, k, alpha, beta = (ev(it, env, parts, q) for q in (2, 3, 4, 5, 14))
        assert ta == "N" and tb == "N"
        if m > 0 and n > 0:
            c = window(it, env, parts, 15, m, n)
            prod = np.array(window(it, env, parts, 6, m, k)) @ np.array(window(it, env, parts, 10, k, n)) if k > 0 else 0.0
            c[:, :] = alpha * prod + beta * np.array(c)

    def pdlacpy_any(it, env, parts):                                                    # SRC/pdlacpy.f: 'L' lower trapezoid, 'U' upper, else all
        uplo = ev(it, env, parts, 0)[0].upper()
        m, n = ev(it, env, parts, 1), ev(it, env, parts, 2)
        if m <= 0 or n <= 0:
            return
        src, dst = window(it, env, parts, 3, m, n), window(it, env, parts, 7, m, n)
        mask = np.tril(np.ones((m, n), bool)) if uplo == "L" else (np.triu(np.ones((m, n), bool)) if uplo == "U" else np.ones((m, n), bool))
        dst[mask] = np.array(src)[mask]

    def pdlaset(it, env, parts):                                                        # SRC/pdlaset.f: off-diagonal part := ALPHA, diagonal := BETA
        uplo = ev(it, env, parts, 0)[0].upper()
        m, n, alpha, beta = ev(it, env, parts, 1), ev(it, env, parts, 2), ev(it, env, parts, 3), ev(it, env, parts, 4)
        if m <= 0 or n <= 0:
            return
        a = window(it, env, parts, 5, m, n)
        i, j = np.indices((m, n))
        a[(i > j) if uplo == "L" else ((i < j) if uplo == "U" else (i != j))] = alpha
        a[i == j] = beta

    def pdlapiv_cols(it, env, parts):                                                   # SRC/pdlapiv.f, ROWCOL = 'C', PIVROC = 'C'
        direc, rowcol, pivroc = (ev(it, env, parts, k)[0].upper() for k in range(3))
        m, n = ev(it, env, parts, 3), ev(it, env, parts, 4)
        assert pivroc == "C"                                                             # IPIV distributed like the rows of A
        ip = ev(it, env, parts, 10)                                                      # IPIV(IP:IP+N-1): global ROW indices of A (pdgetrf.f:118-121)
        piv = env[parts[9]]
        if rowcol == "R":                                                               # rows of sub(A) = A(IA:IA+M-1, JA:JA+N-1) interchanged
            arr, aoff = it.address(parts[5], env)
            ia, ja, desc = ev(it, env, parts, 6), ev(it, env, parts, 7), env[parts[8]]
            full = np.asarray(arr[aoff:aoff + desc[8] * (ja - 1 + n)]).reshape((ja - 1 + n, desc[8])).T
            for i in (range(m) if direc == "F" else range(m - 1, -1, -1)):
                r0, r1 = ia - 1 + i, int(piv[ip - 1 + i]) - 1
                if r0 != r1:
                    full[[r0, r1], ja - 1:ja - 1 + n] = full[[r1, r0], ja - 1:ja - 1 + n]
            return
        a = window(it, env, parts, 5, m, n)
        for j in (range(n) if direc == "F" else range(n - 1, -1, -1)):
            p_ = int(piv[ip - 1 + j]) - ip
            if p_ != j:
                a[:, [j, p_]] = a[:, [p_, j]]

    def dtrmv(it, env, parts):                                                          # BLAS DTRMV: x := A x
        uplo, trans, diag = (ev(it, env, parts, k)[0].upper() for k in range(3))
        n, lda = ev(it, env, parts, 3), ev(it, env, parts, 5)
        a, ao = it.address(parts[4], env); x, xo = it.address(parts[6], env)
        assert trans == "N" and ev(it, env, parts, 7) == 1
        if n > 0:
            t = tri_of(np.asarray(a[ao:ao + lda * (n - 1) + n]).copy() if False else np.array([[a[ao + i + j * lda] for j in range(n)] for i in range(n)]), uplo, diag)
            x[xo:xo + n] = t @ np.array(x[xo:xo + n])

    def dscal(it, env, parts):
        n, alpha = ev(it, env, parts, 0), ev(it, env, parts, 1)
        x, xo = it.address(parts[2], env)
        assert ev(it, env, parts, 3) == 1
        if n > 0:
            x[xo:xo + n] = alpha * np.array(x[xo:xo + n])

    def idamax(n, x, inc):
        return int(np.argmax(np.abs(np.array(x[:n])))) + 1 if n > 0 else 0


    cbs = {"DMATADD": dmatadd,
           "PDGETF2": pdgetf2, "PDLASWP": pdlaswp, "IGAMN2D": nop,
           "PDTRMM": pdtrmm, "PDTRSM": pdtrsm, "PDGEMM": pdgemm, "PDLASET": pdlaset, "PDLAPIV": pdlapiv_cols, "DTRMV": dtrmv, "DSCAL": dscal,
           "BLACS_ABORT": nop,
           "PDGETRF": pdgetrf, "PDLACPY": pdlacpy_any, "DLASSQ": dlassq, "DCOMBSSQ": dcombssq, "IDAMAX": idamax, "PDTREECOMB": nop, "DGSUM2D": nop,
           "DGAMN2D": nop, "IGAMX2D": nop,
           "PDGEMV": pdgemv, "PDAGEMV": pdagemv, "PDCOPY": pdcopy, "PDAXPY": pdaxpy, "PDGETRS": pdgetrs, "DGAMX2D": nop,
           "BLACS_GRIDINFO": gridinfo, "PXERBLA": pxerbla, "PB_TOPGET": topget, "PB_TOPSET": nop, "PCHK1MAT": nop, "PCHK2MAT": nop, "DGEBS2D": nop,
           "DGEBR2D": nop, "IGSUM2D": nop, "PDLABAD": nop, "PDTRSV": pdtrsv, "PDASUM": pdasum, "PDAMAX": pdamax, "PDELGET": pdelget, "DCOPY": dcopy,
           "PDSCAL": pdscal, "PDLAMCH": _pdlamch}
    it = F.Interp(units, cbs)
    it.wrap32 = bool(matgen)
    it.log = log
    return it


def pdgecon(it, norm, lu, anorm, nb, ia=1, ja=1, n=None):
    """lu: the factors of PDGETRF as a global matrix (float64); the reference's PDGECON on a 1 x 1 grid with MB = NB = nb.
    Returns (rcond, info)."""
    M, N = lu.shape
    n = M if n is None else n
    a = np.asfortranarray(lu).reshape(-1, order="F").copy()
    desc = [1, 0, M, N, nb, nb, 0, 0, max(1, M)]
    work, iwork = np.zeros(1), np.zeros(1, np.int64)
    q = it.call("PDGECON", norm, n, a, ia, ja, desc, float(anorm), 0.0, work, -1, iwork, -1, 0)
    assert q["INFO"] == 0, q["INFO"]
    lw, liw = int(work[0]), int(iwork[0])
    work, iwork = np.zeros(lw + 8), np.zeros(liw + 8, np.int64)
    out = it.call("PDGECON", norm, n, a, ia, ja, desc, float(anorm), 0.0, work, lw, iwork, liw, 0)
    return out["RCOND"], out["INFO"]


def pdgerfs(it, trans, a, lu, ipiv, b, x, nb):
    """The reference's PDGERFS on a 1 x 1 grid (MB = NB = nb for A / AF, nb x nb blocks for B / X): refines x in place.
    a, lu: n x n; ipiv: PDGETRF's pivots (1-based); b, x: n x nrhs.  Returns (ferr, berr, info)."""
    n, nrhs = b.shape
    desca = [1, 0, n, n, nb, nb, 0, 0, max(1, n)]
    descb = [1, 0, n, nrhs, nb, nb, 0, 0, max(1, n)]
    af_, a_, b_ = (np.asfortranarray(m_).reshape(-1, order="F").copy() for m_ in (lu, a, b))
    x_ = np.asfortranarray(x).reshape(-1, order="F").copy()
    ip = np.concatenate([np.asarray(ipiv, np.int64), np.zeros(nb, np.int64)])
    ferr, berr = np.zeros(nrhs + 1), np.zeros(nrhs + 1)
    args = lambda work, lw, iwork, liw: (trans, n, nrhs, a_, 1, 1, desca, af_, 1, 1, desca, ip, b_, 1, 1, descb, x_, 1, 1, descb, ferr, berr,  # noqa: E731
                                         work, lw, iwork, liw, 0)
    work, iwork = np.zeros(4), np.zeros(4, np.int64)
    q = it.call("PDGERFS", *args(work, -1, iwork, -1))
    assert q["INFO"] == 0, q["INFO"]
    lw, liw = int(work[0]), int(iwork[0])
    out = it.call("PDGERFS", *args(np.zeros(lw + 8), lw, np.zeros(liw + 8, np.int64), liw))
    x[...] = x_.reshape((n, nrhs), order="F")
    return ferr[:nrhs].copy(), berr[:nrhs].copy(), out["INFO"]


SVX_UNITS = (("SRC", "pdgerfs"), ("SRC", "pdgesvx"), ("SRC", "pdgeequ"), ("SRC", "pdlaqge"), ("SRC", "pdlange"), ("TOOLS", "ilcm"))
LU_UNITS = (("SRC", "pdgetrf"),)        # pdgetrf.f executed on flat arrays (its checks, its blocked loop); PDGETF2 / PDLASWP / PDTRSM / PDGEMM are leaves


def pdgesvx(it, fact, trans, a, af, ipiv, equed, r, c, b, nb):
    """The reference's PDGESVX on a 1 x 1 grid.  a (n x n), b (n x nrhs) are overwritten as the routine overwrites them; af, ipiv, r, c are
    inputs for FACT = 'F' and outputs otherwise.  Returns dict(equed, rcond, info, x, ferr, berr)."""
    n, nrhs = b.shape
    desca = [1, 0, n, n, nb, nb, 0, 0, max(1, n)]
    descb = [1, 0, n, nrhs, nb, nb, 0, 0, max(1, n)]
    flat = lambda m_: np.asfortranarray(m_).reshape(-1, order="F").copy()  # noqa: E731
    a_, af_, b_, x_ = flat(a), flat(af), flat(b), np.zeros(n * nrhs)
    ip = np.concatenate([np.asarray(ipiv, np.int64), np.zeros(nb, np.int64)])
    r_, c_ = np.array(r, float), np.array(c, float)
    ferr, berr = np.zeros(nrhs + 1), np.zeros(nrhs + 1)
    args = lambda work, lw, iwork, liw: (fact, trans, n, nrhs, a_, 1, 1, desca, af_, 1, 1, desca, ip, equed, r_, c_, b_, 1, 1, descb, x_, 1, 1, descb,  # noqa: E731
                                         0.0, ferr, berr, work, lw, iwork, liw, 0)
    work, iwork = np.zeros(4), np.zeros(4, np.int64)
    q = it.call("PDGESVX", *args(work, -1, iwork, -1))
    if q["INFO"] != 0:
        return dict(equed=q["EQUED"], rcond=q["RCOND"], info=q["INFO"])
    lw, liw = int(work[0]), int(iwork[0])
    out = it.call("PDGESVX", *args(np.zeros(lw + 8), lw, np.zeros(liw + 8, np.int64), liw))
    a[...] = a_.reshape((n, n), order="F"); af[...] = af_.reshape((n, n), order="F"); b[...] = b_.reshape((n, nrhs), order="F")
    ipiv[...] = ip[:n]; r[...] = r_; c[...] = c_
    return dict(equed=out["EQUED"], rcond=out["RCOND"], info=out["INFO"], x=x_.reshape((n, nrhs), order="F"), ferr=ferr[:nrhs].copy(),
                berr=berr[:nrhs].copy(), lwork=lw, liwork=liw)


TRI_UNITS = (("SRC", "pdgetri"), ("SRC", "pdtrtri"), ("SRC", "pdtrti2"), ("TOOLS", "ilcm"))


def pdgetri(it, lu, ipiv, nb, ia=1, ja=1, n=None):
    """The reference's PDGETRI (with PDTRTRI / PDTRTI2) on a 1 x 1 grid: lu (global matrix holding the factors of sub(A)) is overwritten by
    the inverse; ipiv: PDGETRF's pivots for the rows of sub(A), as global row indices of A.  Returns INFO."""
    M, N = lu.shape
    n = M - ia + 1 if n is None else n
    a = np.asfortranarray(lu).reshape(-1, order="F").copy()
    desc = [1, 0, M, N, nb, nb, 0, 0, max(1, M)]
    ip = np.zeros(M + nb, np.int64); ip[ia - 1:ia - 1 + n] = ipiv
    work, iwork = np.zeros(4), np.zeros(4, np.int64)
    q = it.call("PDGETRI", n, a, ia, ja, desc, ip, work, -1, iwork, -1, 0)
    assert q["INFO"] == 0, q["INFO"]
    lw, liw = int(work[0]), int(iwork[0])
    out = it.call("PDGETRI", n, a, ia, ja, desc, ip, np.zeros(lw + 8), lw, np.zeros(liw + 8, np.int64), liw, 0)
    lu[...] = a.reshape((M, N), order="F")
    return out["INFO"], lw, liw


CHK_UNITS = (("TESTING/traditional/LIN", "pdgetrrv"), ("TESTING/traditional/LIN", "pdlafchk"), ("SRC", "pdlange"))


def fresid(it, lu, ipiv, nb, aseed):
    """FRESID as the reference's LU driver computes it (pdludriver.f:540-556) on a 1 x 1 grid: PDGETRRV rebuilds P L U from the factors
    in place, PDLAFCHK subtracts the regenerated A = PDMATGEN(aseed) and scales the infinity norm of the difference."""
    m, n = lu.shape
    desc = [1, 0, m, n, nb, nb, 0, 0, max(1, m)]
    a0 = np.zeros(m * n)
    it.call("PDMATGEN", 0, "N", "N", m, n, nb, nb, a0, max(1, m), 0, 0, aseed, 0, m, 0, n, 0, 0, 1, 1)
    work = np.zeros(m * nb + nb * n + m * n + 64)
    anorm = it.call("PDLANGE", "I", m, n, a0, 1, 1, desc, work)["__result__"]
    a = np.asfortranarray(lu).reshape(-1, order="F").copy()
    ip = np.concatenate([np.asarray(ipiv, np.int64), np.zeros(nb, np.int64)])
    it.call("PDGETRRV", m, n, a, 1, 1, desc, ip, work)
    out = it.call("PDLAFCHK", "N", "N", m, n, a, 1, 1, desc, aseed, anorm, 0.0, work)
    return out["FRESID"], anorm
