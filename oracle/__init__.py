"""CPU oracle for the ScaLAPACK dense-LU path (TEST INFRASTRUCTURE ONLY).

ctypes front end of ``oracle/oracle.c`` -- a serial restatement of the
reference's PDGETRF/PDGETRS/PZGETRF algorithm, its test-matrix generators
(TESTING/traditional/LIN/pdmatgen.f) and its residual checks
(pdlafchk.f, pdlaschk.f).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this
package; the product (``scalapack_b200``) must never do so.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile oracle.c with gcc (no GPU, no reference sources needed)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle_next.c")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def _find_openblas() -> str | None:
    try:
        import scipy  # noqa: F401
        base = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
        hits = sorted(glob.glob(os.path.join(base, "libscipy_openblas*.so")))
        return hits[0] if hits else None
    except Exception:
        return None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_wtime.restype = C.c_double
        L.orc_dlange_inf.restype = C.c_double
        for f in ("orc_fresid", "orc_sresid", "orc_zfresid", "orc_zsresid", "orcn_dlange", "orcn_dgecon"):
            getattr(L, f).restype = C.c_double
        p = _find_openblas()
        if p is not None:
            L.orc_init_blas(p.encode())
        _LIB = L
    return _LIB


def have_blas() -> bool:
    return bool(lib().orc_have_blas())


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


def get_threads() -> int:
    return int(lib().orc_get_threads())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- index algebra
def numroc(n, nb, iproc, isrc, nprocs):
    return lib().orc_numroc(n, nb, iproc, isrc, nprocs)


def indxg2p(ig, nb, iproc, isrc, nprocs):
    return lib().orc_indxg2p(ig, nb, iproc, isrc, nprocs)


def indxg2l(ig, nb, iproc, isrc, nprocs):
    return lib().orc_indxg2l(ig, nb, iproc, isrc, nprocs)


def indxl2g(il, nb, iproc, isrc, nprocs):
    return lib().orc_indxl2g(il, nb, iproc, isrc, nprocs)


def infog2l(gr, gc, desc, nprow, npcol, myrow, mycol):
    d = (C.c_int * 9)(*desc)
    lr, lc, rs, cs = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib().orc_infog2l(gr, gc, d, nprow, npcol, myrow, mycol, C.byref(lr), C.byref(lc), C.byref(rs), C.byref(cs))
    return lr.value, lc.value, rs.value, cs.value


def descinit(m, n, mb, nb, irsrc, icsrc, ictxt, lld, nprow, npcol, myrow):
    d = (C.c_int * 9)()
    info = lib().orc_descinit(d, m, n, mb, nb, irsrc, icsrc, ictxt, lld, nprow, npcol, myrow)
    return list(d), info


def chk1mat(ma, mapos0, na, napos0, ia, ja, desc, descpos0, nprow, npcol, myrow, mycol, info=0):
    d = (C.c_int * 9)(*desc)
    inf = C.c_int(info)
    lib().orc_chk1mat(ma, mapos0, na, napos0, ia, ja, d, descpos0, nprow, npcol, myrow, mycol, C.byref(inf))
    return inf.value


# ---------------------------------------------------------------- generators
def pdmatgen(m, n, iseed=100):
    """Global M x N PDMATGEN matrix (Fortran order)."""
    a = np.empty((m, n), dtype=np.float64, order="F")
    lib().orc_pdmatgen_global(m, n, iseed, _p(a), C.c_int64(max(m, 1)))
    return a


def pdmatgen_local(m, n, mb, nb, myrow, mycol, nprow, npcol, iseed=100, iarow=0, iacol=0):
    mp = numroc(m, mb, myrow, iarow, nprow)
    nq = numroc(n, nb, mycol, iacol, npcol)
    a = np.zeros((max(mp, 1), max(nq, 1)), dtype=np.float64, order="F")
    lib().orc_pdmatgen_local(m, n, mb, nb, _p(a), max(mp, 1), iarow, iacol, iseed, myrow, mycol, nprow, npcol)
    return a[:mp, :nq]


def pzmatgen(m, n, iseed=100):
    a = np.empty((m, n), dtype=np.complex128, order="F")
    lib().orc_pzmatgen_global(m, n, iseed, _p(a), C.c_int64(max(m, 1)))
    return a


def matgen64_tile(m, seed, i0, mr, j0, nc, complex_=False):
    """Tile of the 64-bit HPL-style generator (not in the reference)."""
    a = np.empty((mr, nc), dtype=np.complex128 if complex_ else np.float64, order="F")
    f = lib().orc_zmatgen64_tile if complex_ else lib().orc_matgen64_tile
    f(C.c_int64(m), C.c_uint64(seed), C.c_int64(i0), C.c_int64(mr), C.c_int64(j0), C.c_int64(nc), _p(a),
      C.c_int64(max(mr, 1)))
    return a


# ---------------------------------------------------------------- distribution
def scatter(ag, mb, nb, nprow, npcol, myrow, mycol, rsrc=0, csrc=0, lld=None):
    m, n = ag.shape
    ag = np.asfortranarray(ag)
    mp = numroc(m, mb, myrow, rsrc, nprow)
    nq = numroc(n, nb, mycol, csrc, npcol)
    lld = max(1, mp) if lld is None else lld
    al = np.zeros((lld, max(nq, 1)), dtype=ag.dtype, order="F")
    lib().orc_scatter(m, n, mb, nb, rsrc, csrc, nprow, npcol, myrow, mycol, _p(ag), C.c_int64(max(m, 1)), _p(al),
                      C.c_int64(lld), ag.dtype.itemsize)
    return al


def gather_into(ag, al, mb, nb, nprow, npcol, myrow, mycol, rsrc=0, csrc=0):
    m, n = ag.shape
    assert ag.flags.f_contiguous and al.flags.f_contiguous
    lib().orc_gather(m, n, mb, nb, rsrc, csrc, nprow, npcol, myrow, mycol, _p(ag), C.c_int64(max(m, 1)), _p(al),
                     C.c_int64(al.shape[0]), ag.dtype.itemsize)


def ipiv_local(m, mn, mb, nprow, myrow, ipiv_g, nloc, rsrc=0, fill=-1):
    out = np.empty(nloc, dtype=np.int32)
    g = np.ascontiguousarray(ipiv_g, dtype=np.int32)
    lib().orc_ipiv_local(m, mn, mb, rsrc, nprow, myrow, _p(g), _p(out), nloc, fill)
    return out


# ---------------------------------------------------------------- LU / solve / residuals
def tie_grid(nprow=1, rsrc=0):
    """Exact ties in the pivot search are broken as the reference's PDAMAX breaks them on a grid with `nprow` process rows (the lowest
    absolute process row wins, pdamax_.c:436-458); nprow = 1 (default): the first global index."""
    lib().orc_set_tie_grid(int(nprow), int(rsrc))


def getrf(a, nb, phase_times=False):
    """In-place restated PDGETRF on the global matrix.  Returns (ipiv[1-based], info)."""
    assert a.flags.f_contiguous
    m, n = a.shape
    ipiv = np.zeros(max(min(m, n), 1), dtype=np.int32)
    if a.dtype == np.complex128:
        info = lib().orc_zgetrf(m, n, _p(a), C.c_int64(max(m, 1)), nb, _p(ipiv))
        return ipiv[:min(m, n)], info
    assert a.dtype == np.float64
    pt = np.zeros(4)
    info = lib().orc_dgetrf(m, n, _p(a), C.c_int64(max(m, 1)), nb, _p(ipiv), _p(pt))
    if phase_times:
        return ipiv[:min(m, n)], info, pt
    return ipiv[:min(m, n)], info


def getrf_steps(a, nb, nsteps):
    """First `nsteps` block steps only (bounded CPU-baseline sample). Returns (info, flops)."""
    m, n = a.shape
    ipiv = np.zeros(max(min(m, n), 1), dtype=np.int32)
    fl = C.c_double(0)
    info = lib().orc_dgetrf_steps(m, n, _p(a), C.c_int64(max(m, 1)), nb, _p(ipiv), nsteps, C.byref(fl))
    return info, fl.value


def getrs(lu, ipiv, b, trans="N"):
    assert lu.flags.f_contiguous and b.flags.f_contiguous
    n = lu.shape[0]
    nrhs = b.shape[1]
    ip = np.ascontiguousarray(ipiv, dtype=np.int32)
    if lu.dtype == np.complex128:
        lib().orc_zgetrs(C.c_char(trans.encode()), n, nrhs, _p(lu), C.c_int64(max(n, 1)), _p(ip), _p(b),
                         C.c_int64(b.shape[0]))
    else:
        lib().orc_dgetrs(C.c_char(trans.encode()), n, nrhs, _p(lu), C.c_int64(max(n, 1)), _p(ip), _p(b),
                         C.c_int64(b.shape[0]))
    return b


def fresid(lu, ipiv, a0):
    m, n = lu.shape
    lu = np.asfortranarray(lu)
    a0 = np.asfortranarray(a0)
    ip = np.ascontiguousarray(ipiv, dtype=np.int32)
    f = lib().orc_zfresid if lu.dtype == np.complex128 else lib().orc_fresid
    return f(m, n, _p(lu), C.c_int64(max(m, 1)), _p(ip), _p(a0), C.c_int64(max(m, 1)))


def sresid(a0, x, b0):
    n = a0.shape[0]
    a0 = np.asfortranarray(a0)
    x = np.asfortranarray(x)
    b0 = np.asfortranarray(b0)
    f = lib().orc_zsresid if a0.dtype == np.complex128 else lib().orc_sresid
    return f(n, x.shape[1], _p(a0), C.c_int64(max(n, 1)), _p(x), C.c_int64(x.shape[0]), _p(b0),
             C.c_int64(b0.shape[0]))


def lu_tolerance_ok(lu_test, lu_ref, a0):
    """LU-factor tolerance of SURVEY.md 8a(vi): with identical IPIV,
    max|LU_test - LU_ref| / (||A||_inf * N * eps) < 1."""
    n = max(a0.shape)
    anorm = np.abs(a0).sum(axis=1).max()
    err = np.abs(lu_test - lu_ref).max() / (anorm * n * 2.0 ** -53)
    return err, err < 1.0


# ---------------------------------------------------------------- SURVEY 8(f) rows (oracle_next.c)
def _f(a):
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and (a.flags.f_contiguous or a.ndim == 1)
    return a.ctypes.data_as(C.c_void_p)


def dlange(norm, a):
    """SRC/pdlange.f on the global matrix."""
    m, n = a.shape
    return float(lib().orcn_dlange(C.c_char(norm.encode()), m, n, _f(a), C.c_int64(a.strides[1] // 8)))


def dgeequ(a):
    """SRC/pdgeequ.f: returns (r, c, rowcnd, colcnd, amax, info)."""
    m, n = a.shape
    r, c = np.zeros(m), np.zeros(n)
    rc, cc, am = C.c_double(1.0), C.c_double(1.0), C.c_double(0.0)
    info = lib().orcn_dgeequ(m, n, _f(a), C.c_int64(a.strides[1] // 8), _f(r), _f(c), C.byref(rc), C.byref(cc), C.byref(am))
    return r, c, rc.value, cc.value, am.value, int(info)


def dlaqge(a, r, c, rowcnd, colcnd, amax):
    """SRC/pdlaqge.f: scales a in place, returns EQUED."""
    m, n = a.shape
    L = lib(); L.orcn_dlaqge.restype = C.c_char
    return L.orcn_dlaqge(m, n, _f(a), C.c_int64(a.strides[1] // 8), _f(r), _f(c), C.c_double(rowcnd), C.c_double(colcnd),
                         C.c_double(amax)).decode()


def lacon_keep_est(keep):
    """False (default): PDLACON as the reference's source behaves (EST reset on every call, pdlacon.f:188-189); True: LAPACK's DLACON."""
    lib().orcn_lacon_keep_est(1 if keep else 0)


def dgecon(norm, lu, anorm):
    """SRC/pdgecon.f + pdlacon.f on the factors of getrf(); returns RCOND."""
    n = lu.shape[0]
    return float(lib().orcn_dgecon(C.c_char(norm.encode()), n, _f(lu), C.c_int64(lu.strides[1] // 8), C.c_double(anorm)))


def dgerfs(trans, a, af, ipiv, b, x):
    """SRC/pdgerfs.f: refines x in place; returns (ferr, berr)."""
    n, nrhs = b.shape
    ferr, berr = np.zeros(nrhs), np.zeros(nrhs)
    ip = np.ascontiguousarray(ipiv, dtype=np.int32)
    lib().orcn_dgerfs(C.c_char(trans.encode()), n, nrhs, _f(a), C.c_int64(a.strides[1] // 8), _f(af), C.c_int64(af.strides[1] // 8), _p(ip),
                      _f(b), C.c_int64(b.strides[1] // 8), _f(x), C.c_int64(x.strides[1] // 8), _f(ferr), _f(berr))
    return ferr, berr


def dgesvx(fact, trans, a, af, ipiv, equed, r, c, b, x, nb=64):
    """SRC/pdgesvx.f (IA = JA = 1): a, af, ipiv, r, c, b, x in / out; returns (equed, rcond, ferr, berr, info)."""
    n, nrhs = b.shape
    ferr, berr = np.zeros(nrhs), np.zeros(nrhs)
    rcond = C.c_double(0.0)
    eq = C.c_char(equed.encode())
    assert ipiv.dtype == np.int32
    info = lib().orcn_dgesvx(C.c_char(fact.encode()), C.c_char(trans.encode()), n, nrhs, _f(a), C.c_int64(a.strides[1] // 8), _f(af),
                             C.c_int64(af.strides[1] // 8), _p(ipiv), C.byref(eq), _f(r), _f(c), _f(b), C.c_int64(b.strides[1] // 8), _f(x),
                             C.c_int64(x.strides[1] // 8), C.byref(rcond), _f(ferr), _f(berr), nb)
    return eq.value.decode(), rcond.value, ferr, berr, int(info)


def dpotrf(uplo, a, nb):
    """SRC/pdpotrf.f on the global matrix (only the UPLO triangle is referenced); returns INFO."""
    n = a.shape[0]
    return int(lib().orcn_dpotrf(C.c_char(uplo.encode()), n, _f(a), C.c_int64(a.strides[1] // 8), nb))


def dpotrs(uplo, a, b):
    """SRC/pdpotrs.f: b <- inv(A) b from the Cholesky factor."""
    n, nrhs = b.shape
    lib().orcn_dpotrs(C.c_char(uplo.encode()), n, nrhs, _f(a), C.c_int64(a.strides[1] // 8), _f(b), C.c_int64(b.strides[1] // 8))


def dgetri(lu, ipiv, nb):
    """SRC/pdgetri.f: the inverse in place from the factors of getrf(); returns INFO."""
    n = lu.shape[0]
    ip = np.ascontiguousarray(ipiv, dtype=np.int32)
    return int(lib().orcn_dgetri(n, _f(lu), C.c_int64(lu.strides[1] // 8), _p(ip), nb))


# ---------------------------------------------------------------- PBLAS definitions (numpy; the reference's Purpose blocks)
def dgemm(transa, transb, alpha, a, b, beta, c):
    """PBLAS/SRC/pdgemm_.c:36-50: C := alpha op(A) op(B) + beta C on the global sub-matrices."""
    opa = a.T if transa.upper() in "TC" else a
    opb = b.T if transb.upper() in "TC" else b
    return alpha * (opa @ opb) + (beta * c if beta != 0.0 else 0.0)


def dtrsm(side, uplo, transa, diag, alpha, a, b):
    """PBLAS/SRC/pdtrsm_.c:34-52: X with op(A) X = alpha B (side L) or X op(A) = alpha B (side R); only the UPLO triangle of A
    is referenced, DIAG = 'U' assumes a unit diagonal."""
    t = np.tril(a) if uplo.upper() == "L" else np.triu(a)
    if diag.upper() == "U":
        t = t - np.diag(np.diag(t)) + np.eye(a.shape[0])
    opt = t.T if transa.upper() in "TC" else t
    if side.upper() == "L":
        return np.linalg.solve(opt, alpha * b)
    return np.linalg.solve(opt.T, alpha * b.T).T


# ---------------------------------------------------------------- oracle/_ref: the reference's own PDGEMR2D (REDIST/SRC/pdgemr.c)
_REF = None


def ref_redist_lib():
    """oracle/_ref/libref_redist.so = the reference's REDIST sources compiled in place + oracle/ref_redist.c (thread-based mini-BLACS);
    None when it has not been built (no /root/reference at build time)."""
    global _REF
    if _REF is None:
        so = os.path.join(_HERE, "_ref", "libref_redist.so")
        if not os.path.exists(so):
            try:
                subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
            except Exception:
                pass
        _REF = C.CDLL(so) if os.path.exists(so) else False
    return _REF or None


def ref_pdgemr2d(ag, m, n, ia, ja, ib, jb, grid_a, blk_a, src_a, shape_b, grid_b, blk_b, src_b, fill=-9923.0):
    """The REFERENCE redistributes sub(A) of the global matrix ag (distributed on grid_a with blk_a blocks from process src_a) into
    sub(B) of a `fill`-initialised B (shape_b on grid_b, blk_b, src_b).  Returns the list of local B arrays (LOCr x LOCc, Fortran order;
    None for processes outside B's grid), indexed by the process number in the global 1 x np context (= row-major position in a grid)."""
    L = ref_redist_lib()
    assert L is not None
    (pa, qa), (pb, qb) = grid_a, grid_b
    np_ = max(pa * qa, pb * qb)
    z = np.iscomplexobj(ag)                                   # complex: the reference's PZGEMR2D (REDIST/SRC/pzgemr.c)
    dt = np.complex128 if z else np.float64
    ag = np.asfortranarray(ag, dtype=dt)
    ma, na = ag.shape
    mb, nb = shape_b
    w = 2 if z else 1
    stride = w * (mb + 1) * (nb + 1)
    bout = np.zeros(np_ * stride)
    dims = np.zeros(2 * np_, np.int32)
    rc = L.ref_pdgemr2d_run(np_, m, n, ia, ja, ib, jb, pa, qa, ma, na, blk_a[0], blk_a[1], src_a[0], src_a[1], pb, qb, mb, nb, blk_b[0], blk_b[1],
                            src_b[0], src_b[1], ag.ctypes.data_as(C.c_void_p), C.c_double(fill), _f(bout), C.c_int64(stride), _p(dims), 8 * w)
    assert rc == 0
    out = []
    for r in range(np_):
        ml, nl = int(dims[2 * r]), int(dims[2 * r + 1])
        if r >= pb * qb:
            out.append(None)
        else:
            lld = max(1, ml)
            loc = bout[r * stride:r * stride + w * lld * max(1, nl)].view(dt)
            out.append(np.asfortranarray(loc.reshape((lld, max(1, nl)), order="F")[:ml, :nl]))
    return out
