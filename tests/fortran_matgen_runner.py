"""Runs the reference's OWN test-matrix generator -- TESTING/traditional/LIN/pdmatgen.f, pzmatgen.f and the 31-bit linear congruential
arithmetic of pmatgeninc.f (LADD, LMUL, XJUMPM, SETRAN, JUMPIT, PDRAND), read from /root/reference -- with the mini interpreter of
tests/fortran77_mini.py, once per process (MYROW, MYCOL) of an emulated NPROW x NPCOL grid.  Nothing is stood in for: INTEGER arithmetic
wraps at 32 bits as the reference's compilers make it (LMUL corrects for the overflow, pmatgeninc.f:67-77), COMMON /RANCOM/ is shared storage, and
NUMROC / ICEIL are the reference's too.  TEST INFRASTRUCTURE: pins the closed-form jump-ahead generator of oracle/oracle.c
(orc_pdmatgen_*) -- and through it every input matrix of the parity tests -- against the reference's source text."""
import os

import numpy as np

import fortran77_mini as F


def make(ref_root="/root/reference"):
    lin = os.path.join(ref_root, "TESTING", "traditional", "LIN")
    units = []
    for f in ("pmatgeninc.f", "pdmatgen.f", "pzmatgen.f"):
        units += F.parse_file(open(os.path.join(lin, f)).read())
    units += [F.parse(open(os.path.join(ref_root, "TOOLS", f + ".f")).read()) for f in ("numroc", "iceil")]
    log = []
    it = F.Interp(units, {"PXERBLA": lambda it_, env, parts: log.append(("PXERBLA", it_.eval(parts[1], env), it_.eval(parts[2], env)))})
    it.wrap32 = True
    it.log = log
    return it


def _numroc(it, n, nb, iproc, isrc, nprocs):
    return it.call("NUMROC", n, nb, iproc, isrc, nprocs)["__result__"]


def local(it, m, n, mb, nb, myrow, mycol, nprow, npcol, iseed=100, iarow=0, iacol=0, aform="N", diag="N", complex_=False):
    """The local piece process (myrow, mycol) generates: PxMATGEN called as TESTING/traditional/LIN/pdludriver.f:417-420 calls it
    (IROFF = ICOFF = 0, IRNUM = LOCr(M), ICNUM = LOCc(N))."""
    mp, nq = _numroc(it, m, mb, myrow, iarow, nprow), _numroc(it, n, nb, mycol, iacol, npcol)
    a = np.zeros((max(1, mp), max(1, nq)), dtype=np.complex128 if complex_ else np.float64, order="F")
    it.call("PZMATGEN" if complex_ else "PDMATGEN", 0, aform, diag, m, n, mb, nb, a, max(1, mp), iarow, iacol, iseed, 0, mp, 0, nq,
            myrow, mycol, nprow, npcol)
    return a[:mp, :nq]


def global_(it, m, n, mb, nb, nprow, npcol, iseed=100, iarow=0, iacol=0, aform="N", diag="N", complex_=False):
    """The global matrix assembled from every process's local piece (2D block-cyclic, first block on (iarow, iacol))."""
    g = np.zeros((m, n), dtype=np.complex128 if complex_ else np.float64, order="F")
    for pr in range(nprow):
        rows = [i for i in range(m) if ((i // mb) + iarow) % nprow == pr]
        for pc in range(npcol):
            cols = [j for j in range(n) if ((j // nb) + iacol) % npcol == pc]
            loc = local(it, m, n, mb, nb, pr, pc, nprow, npcol, iseed, iarow, iacol, aform, diag, complex_)
            assert loc.shape == (len(rows), len(cols))
            if rows and cols:
                g[np.ix_(rows, cols)] = loc
    return g


def make_checks(ref_root="/root/reference"):
    """+ TESTING/traditional/LIN/pdlaschk.f (the solve residual every bench line and parity test reports as SRESID), executed with the
    executed generator underneath; leaves: PBDTRAN (local transpose of a block of X), DGEMM, DLASET, IDAMAX."""
    lin = os.path.join(ref_root, "TESTING", "traditional", "LIN")
    units = []
    for f in ("pmatgeninc.f", "pdmatgen.f"):
        units += F.parse_file(open(os.path.join(lin, f)).read())
    units += [F.parse(open(os.path.join(lin, "pdlaschk.f")).read())]
    units += [F.parse(open(os.path.join(ref_root, "TOOLS", f + ".f")).read()) for f in ("numroc", "iceil", "infog2l", "indxg2p", "indxg2l")]
    log = []

    def ev(it, env, parts, k):
        return it.eval(parts[k], env)

    def gridinfo(it, env, parts):
        for name, v in zip(parts[1:], (1, 1, 0, 0)):
            env[name] = v

    def nop(it, env, parts):
        pass

    def mat(it, env, part, m, n, ld):
        """m x n column-major block at the address `part` with leading dimension ld: index array into the flat storage"""
        arr, off = it.address(part, env)
        idx = off + np.arange(m)[:, None] + ld * np.arange(n)[None, :]
        return arr, idx

    # PBLAS/SRC/PBBLAS/pbdtran.f, ADIST = 'Column', TRANS = 'T', one process: C (N x M, ldc) := A' (A: M x N, lda) + beta C
    def pbdtran(it, env, parts):
        m, n = ev(it, env, parts, 3), ev(it, env, parts, 4)
        a, ia = mat(it, env, parts[6], m, n, ev(it, env, parts, 7))
        beta = ev(it, env, parts, 8)
        c, ic = mat(it, env, parts[9], n, m, ev(it, env, parts, 10))
        c[ic] = a[ia].T + (beta * c[ic] if beta != 0.0 else 0.0)

    def dgemm(it, env, parts):
        ta, tb = ev(it, env, parts, 0)[0].upper(), ev(it, env, parts, 1)[0].upper()
        m, n, k, alpha, beta = (ev(it, env, parts, q) for q in (2, 3, 4, 5, 10))
        a, ia = mat(it, env, parts[6], *((m, k) if ta == "N" else (k, m)), ev(it, env, parts, 7))
        b, ib = mat(it, env, parts[8], *((k, n) if tb == "N" else (n, k)), ev(it, env, parts, 9))
        c, ic = mat(it, env, parts[11], m, n, ev(it, env, parts, 12))
        pa, pb = (a[ia] if ta == "N" else a[ia].T), (b[ib] if tb == "N" else b[ib].T)
        c[ic] = alpha * (pa @ pb) + (beta * c[ic] if beta != 0.0 else 0.0)

    def dlaset(it, env, parts):
        m, n, alpha, beta = ev(it, env, parts, 1), ev(it, env, parts, 2), ev(it, env, parts, 3), ev(it, env, parts, 4)
        a, ia = mat(it, env, parts[5], m, n, ev(it, env, parts, 6))
        a[ia] = alpha
        for i in range(min(m, n)):
            a[ia[i, i]] = beta

    def idamax(it, env, parts):                                    # by address: IDAMAX( N, X( k ), 1 )
        n = it.eval(parts[0], env)
        x, off = it.address(parts[1], env)
        return int(np.argmax(np.abs(np.asarray(x[off:off + n])))) + 1 if n > 0 else 0

    cbs = {"BLACS_GRIDINFO": gridinfo, "PXERBLA": lambda it_, env, parts: log.append(("PXERBLA", parts)), "PBDTRAN": pbdtran, "DGEMM": dgemm, "DLASET": dlaset,
           "DGSUM2D": nop, "DGAMX2D": nop, "DGEBS2D": nop, "DGEBR2D": nop, "DGESD2D": nop, "DGERV2D": nop,
           "PDLAMCH": lambda ictxt, cmach: {"E": 2.0 ** -53, "S": float(np.finfo(np.float64).tiny)}[str(cmach)[:1].upper()]}
    it = F.Interp(units, cbs)
    it.raw_functions = {"IDAMAX": idamax}
    it.wrap32 = True
    it.log = log
    return it


def pdlaschk(it, x, n, nrhs, nb, nbrhs, aseed, bseed, anorm):
    """The reference's PDLASCHK on a 1 x 1 grid: x (n x nrhs) is the computed solution of A x = b with A = PDMATGEN(aseed) (n x n, blocks
    nb x nb) and b = PDMATGEN(bseed) (n x nrhs, blocks nb x nbrhs); returns RESID."""
    desca = [1, 0, n, n, nb, nb, 0, 0, max(1, n)]
    descx = [1, 0, n, nrhs, nb, nbrhs, 0, 0, max(1, n)]
    xf = np.asfortranarray(x).reshape(-1, order="F").copy()
    work = np.zeros(n * nbrhs + n * nbrhs + max(nb, 1) * n + 4 * nbrhs + 64)
    out = it.call("PDLASCHK", "N", "N", n, nrhs, xf, 1, 1, descx, aseed, 1, 1, desca, bseed, float(anorm), 0.0, work)
    return out["RESID"]
