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
