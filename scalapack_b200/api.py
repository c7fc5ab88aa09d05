"""ctypes mirror of the reference's Fortran interface for the LU path.

Every function keeps the reference's name, argument order and meaning (SRC/pdgetrf.f:1, pdgetrs.f:1-2,
pdgesv.f:1-2, TOOLS/descinit.f:1-2, BLACS/SRC/blacs_*.c); scalars are passed by value here and by
reference underneath, INFO is returned.  Matrix arguments may be

* numpy arrays in Fortran order (HOST memory: staged through HBM by the library), or
* torch CUDA tensors / raw integer device pointers (DEVICE memory: factored in place).

Index arguments (IA, JA, ...) and IPIV values are 1-based like the reference.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

DTYPE_, CTXT_, M_, N_, MB_, NB_, RSRC_, CSRC_, LLD_ = range(9)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "lib", "libscalapack_b200.so")
_lib = None


def have_library() -> bool:
    return os.path.exists(_SO)


def lib():
    """The C-ABI library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(f"{_SO} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(_SO, mode=C.RTLD_GLOBAL)
        for f in ("slb200_last_factor_ms", "slb200_last_solve_ms", "slb200_last_update_ms", "slb200_last_update_flops",
                  "slb200_pdlaschk", "slb200_test_gemm", "slb200_test_panel", "slb200_bench_dmma_tflops",
                  "slb200_bench_dfma_tflops", "slb200_bench_copy_gbs"):
            getattr(L, f).restype = C.c_double
        L.slb200_get_counter.restype = C.c_int64
        L.slb200_last_update_launches.restype = C.c_int64
        L.slb200_version.restype = C.c_char_p
        _lib = L
    return _lib


def has_cuda() -> bool:
    return bool(lib().slb200_has_cuda())


def device() -> int:
    """Index of the GPU this BLACS process drives (LOCAL_RANK modulo the device count); device-resident operands
    must live there."""
    return int(lib().slb200_device())


def _i(v):
    return C.byref(C.c_int(int(v)))


def _ptr(a):
    """(void*) of a numpy array, torch tensor, ctypes pointer or integer address."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return a


def _desc(d):
    return (C.c_int * 9)(*[int(x) for x in d])


# ------------------------------------------------------------------ BLACS
def blacs_pinfo():
    me, n = C.c_int(), C.c_int()
    lib().blacs_pinfo_(C.byref(me), C.byref(n))
    return me.value, n.value


def blacs_get(ictxt=-1, what=0):
    v = C.c_int()
    lib().blacs_get_(_i(ictxt), _i(what), C.byref(v))
    return v.value


def blacs_gridinit(ictxt, order, nprow, npcol):
    c = C.c_int(ictxt)
    lib().blacs_gridinit_(C.byref(c), order.encode(), _i(nprow), _i(npcol))
    return c.value


def blacs_gridinfo(ictxt):
    p, q, r, c = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib().blacs_gridinfo_(_i(ictxt), C.byref(p), C.byref(q), C.byref(r), C.byref(c))
    return p.value, q.value, r.value, c.value


def blacs_gridexit(ictxt):
    lib().blacs_gridexit_(_i(ictxt))


def blacs_exit(notdone=0):
    lib().blacs_exit_(_i(notdone))


def blacs_barrier(ictxt, scope="All"):
    lib().blacs_barrier_(_i(ictxt), scope.encode())


def blacs_pnum(ictxt, prow, pcol):
    return lib().blacs_pnum_(_i(ictxt), _i(prow), _i(pcol))


def blacs_pcoord(ictxt, pnum):
    r, c = C.c_int(), C.c_int()
    lib().blacs_pcoord_(_i(ictxt), _i(pnum), C.byref(r), C.byref(c))
    return r.value, c.value


def sl_init(nprow, npcol):
    c = C.c_int()
    lib().sl_init_(C.byref(c), _i(nprow), _i(npcol))
    return c.value


# ------------------------------------------------------------------ TOOLS
def numroc(n, nb, iproc, isrcproc, nprocs):
    return lib().numroc_(_i(n), _i(nb), _i(iproc), _i(isrcproc), _i(nprocs))


def indxg2p(ig, nb, iproc, isrc, nprocs):
    return lib().indxg2p_(_i(ig), _i(nb), _i(iproc), _i(isrc), _i(nprocs))


def indxg2l(ig, nb, iproc, isrc, nprocs):
    return lib().indxg2l_(_i(ig), _i(nb), _i(iproc), _i(isrc), _i(nprocs))


def indxl2g(il, nb, iproc, isrc, nprocs):
    return lib().indxl2g_(_i(il), _i(nb), _i(iproc), _i(isrc), _i(nprocs))


def infog2l(gr, gc, desc, nprow, npcol, myrow, mycol):
    lr, lc, rs, cs = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib().infog2l_(_i(gr), _i(gc), _desc(desc), _i(nprow), _i(npcol), _i(myrow), _i(mycol), C.byref(lr), C.byref(lc),
                   C.byref(rs), C.byref(cs))
    return lr.value, lc.value, rs.value, cs.value


def iceil(a, b):
    return lib().iceil_(_i(a), _i(b))


def ilcm(a, b):
    return lib().ilcm_(_i(a), _i(b))


def descinit(m, n, mb, nb, irsrc, icsrc, ictxt, lld):
    d = (C.c_int * 9)()
    info = C.c_int()
    lib().descinit_(d, _i(m), _i(n), _i(mb), _i(nb), _i(irsrc), _i(icsrc), _i(ictxt), _i(lld), C.byref(info))
    return list(d), info.value


def chk1mat(ma, mapos0, na, napos0, ia, ja, desca, descapos0, info=0):
    inf = C.c_int(info)
    lib().chk1mat_(_i(ma), _i(mapos0), _i(na), _i(napos0), _i(ia), _i(ja), _desc(desca), _i(descapos0), C.byref(inf))
    return inf.value


# ------------------------------------------------------------------ LU factor / solve
def _ipiv_ptr(ipiv):
    assert isinstance(ipiv, np.ndarray) and ipiv.dtype == np.int32 and ipiv.flags.c_contiguous, \
        "IPIV must be a contiguous numpy int32 array of LOCr(M_A)+MB_A entries"
    return ipiv.ctypes.data_as(C.c_void_p)


def pdgetrf(m, n, a, ia, ja, desca, ipiv):
    info = C.c_int()
    lib().pdgetrf_(_i(m), _i(n), _ptr(a), _i(ia), _i(ja), _desc(desca), _ipiv_ptr(ipiv), C.byref(info))
    return info.value


def pzgetrf(m, n, a, ia, ja, desca, ipiv):
    info = C.c_int()
    lib().pzgetrf_(_i(m), _i(n), _ptr(a), _i(ia), _i(ja), _desc(desca), _ipiv_ptr(ipiv), C.byref(info))
    return info.value


def pdgetrs(trans, n, nrhs, a, ia, ja, desca, ipiv, b, ib, jb, descb):
    info = C.c_int()
    lib().pdgetrs_(trans.encode(), _i(n), _i(nrhs), _ptr(a), _i(ia), _i(ja), _desc(desca), _ipiv_ptr(ipiv), _ptr(b),
                   _i(ib), _i(jb), _desc(descb), C.byref(info))
    return info.value


def pzgetrs(trans, n, nrhs, a, ia, ja, desca, ipiv, b, ib, jb, descb):
    info = C.c_int()
    lib().pzgetrs_(trans.encode(), _i(n), _i(nrhs), _ptr(a), _i(ia), _i(ja), _desc(desca), _ipiv_ptr(ipiv), _ptr(b),
                   _i(ib), _i(jb), _desc(descb), C.byref(info))
    return info.value


def pdgesv(n, nrhs, a, ia, ja, desca, ipiv, b, ib, jb, descb):
    info = C.c_int()
    lib().pdgesv_(_i(n), _i(nrhs), _ptr(a), _i(ia), _i(ja), _desc(desca), _ipiv_ptr(ipiv), _ptr(b), _i(ib), _i(jb),
                  _desc(descb), C.byref(info))
    return info.value


def pzgesv(n, nrhs, a, ia, ja, desca, ipiv, b, ib, jb, descb):
    info = C.c_int()
    lib().pzgesv_(_i(n), _i(nrhs), _ptr(a), _i(ia), _i(ja), _desc(desca), _ipiv_ptr(ipiv), _ptr(b), _i(ib), _i(jb),
                  _desc(descb), C.byref(info))
    return info.value


# ------------------------------------------------------------------ consumers of the factors
def _d(v):
    return C.byref(C.c_double(float(v)))


def _dptr(a):
    assert isinstance(a, np.ndarray) and a.dtype == np.float64
    return a.ctypes.data_as(C.c_void_p)


def pdlange(norm, m, n, a, ia, ja, desca):
    """SRC/pdlange.f: 'M', '1' / 'O', 'I', 'F' / 'E' norm of sub(A)."""
    L = lib()
    L.pdlange_.restype = C.c_double
    return float(L.pdlange_(norm.encode(), _i(m), _i(n), _ptr(a), _i(ia), _i(ja), _desc(desca), C.c_void_p(0)))


def pdgeequ(m, n, a, ia, ja, desca, r, c):
    """SRC/pdgeequ.f: fills the local arrays R (LOCr(M_A)) and C (LOCc(N_A)); returns (rowcnd, colcnd, amax, info)."""
    rc, cc, am, info = C.c_double(), C.c_double(), C.c_double(), C.c_int()
    lib().pdgeequ_(_i(m), _i(n), _ptr(a), _i(ia), _i(ja), _desc(desca), _dptr(r), _dptr(c), C.byref(rc), C.byref(cc), C.byref(am),
                   C.byref(info))
    return rc.value, cc.value, am.value, info.value


def pdlaqge(m, n, a, ia, ja, desca, r, c, rowcnd, colcnd, amax):
    """SRC/pdlaqge.f: scales sub(A) in place; returns EQUED."""
    eq = C.create_string_buffer(b"N", 2)
    lib().pdlaqge_(_i(m), _i(n), _ptr(a), _i(ia), _i(ja), _desc(desca), _dptr(r), _dptr(c), _d(rowcnd), _d(colcnd), _d(amax), eq)
    return eq.value[:1].decode()


def pdgecon(norm, n, a, ia, ja, desca, anorm, lwork=None, liwork=None):
    """SRC/pdgecon.f: returns (rcond, info); lwork / liwork default to the sizes a workspace query reports."""
    rcond, info = C.c_double(), C.c_int()
    w1, iw1 = np.zeros(1), np.zeros(1, np.int32)
    lib().pdgecon_(norm.encode(), _i(n), _ptr(a), _i(ia), _i(ja), _desc(desca), _d(anorm), C.byref(rcond), _dptr(w1), _i(-1),
                   _ipiv_ptr(iw1), _i(-1), C.byref(info))
    if info.value != 0:
        return rcond.value, info.value
    lw = int(w1[0]) if lwork is None else lwork
    liw = int(iw1[0]) if liwork is None else liwork
    work, iwork = np.zeros(max(1, lw)), np.zeros(max(1, liw), np.int32)
    lib().pdgecon_(norm.encode(), _i(n), _ptr(a), _i(ia), _i(ja), _desc(desca), _d(anorm), C.byref(rcond), _dptr(work), _i(lw),
                   _ipiv_ptr(iwork), _i(liw), C.byref(info))
    return rcond.value, info.value


def pdgerfs(trans, n, nrhs, a, ia, ja, desca, af, iaf, jaf, descaf, ipiv, b, ib, jb, descb, x, ix, jx, descx, ferr, berr):
    """SRC/pdgerfs.f: refines X in place, fills the local arrays FERR / BERR (LOCc(N_B)); returns info."""
    info = C.c_int()
    args = lambda work, lw, iwork, liw: (trans.encode(), _i(n), _i(nrhs), _ptr(a), _i(ia), _i(ja), _desc(desca), _ptr(af), _i(iaf),  # noqa: E731
                                         _i(jaf), _desc(descaf), _ipiv_ptr(ipiv), _ptr(b), _i(ib), _i(jb), _desc(descb), _ptr(x), _i(ix),
                                         _i(jx), _desc(descx), _dptr(ferr), _dptr(berr), _dptr(work), _i(lw), _ipiv_ptr(iwork), _i(liw),
                                         C.byref(info))
    w1, iw1 = np.zeros(1), np.zeros(1, np.int32)
    lib().pdgerfs_(*args(w1, -1, iw1, -1))
    if info.value != 0:
        return info.value
    work, iwork = np.zeros(max(1, int(w1[0]))), np.zeros(max(1, int(iw1[0])), np.int32)
    lib().pdgerfs_(*args(work, work.size, iwork, iwork.size))
    return info.value


def pdgesvx(fact, trans, n, nrhs, a, ia, ja, desca, af, iaf, jaf, descaf, ipiv, equed, r, c, b, ib, jb, descb, x, ix, jx, descx,
            ferr, berr):
    """SRC/pdgesvx.f: returns (equed, rcond, info)."""
    rcond, info = C.c_double(), C.c_int()
    eq = C.create_string_buffer(equed.encode()[:1], 2)
    args = lambda work, lw, iwork, liw: (fact.encode(), trans.encode(), _i(n), _i(nrhs), _ptr(a), _i(ia), _i(ja), _desc(desca), _ptr(af),  # noqa: E731
                                         _i(iaf), _i(jaf), _desc(descaf), _ipiv_ptr(ipiv), eq, _dptr(r), _dptr(c), _ptr(b), _i(ib), _i(jb),
                                         _desc(descb), _ptr(x), _i(ix), _i(jx), _desc(descx), C.byref(rcond), _dptr(ferr), _dptr(berr),
                                         _dptr(work), _i(lw), _ipiv_ptr(iwork), _i(liw), C.byref(info))
    w1, iw1 = np.zeros(1), np.zeros(1, np.int32)
    lib().pdgesvx_(*args(w1, -1, iw1, -1))
    if info.value != 0:
        return eq.value[:1].decode(), rcond.value, info.value
    work, iwork = np.zeros(max(1, int(w1[0]))), np.zeros(max(1, int(iw1[0])), np.int32)
    lib().pdgesvx_(*args(work, work.size, iwork, iwork.size))
    return eq.value[:1].decode(), rcond.value, info.value


# ------------------------------------------------------------------ redistribution
def pdgemr2d(m, n, a, ia, ja, desca, b, ib, jb, descb, ictxt):
    """REDIST/SRC/pdgemr.c: sub(B) <- sub(A) across layouts / grids; every process of context ictxt calls."""
    lib().pdgemr2d_(_i(m), _i(n), _ptr(a), _i(ia), _i(ja), _desc(desca), _ptr(b), _i(ib), _i(jb), _desc(descb), _i(ictxt))


def pzgemr2d(m, n, a, ia, ja, desca, b, ib, jb, descb, ictxt):
    lib().pzgemr2d_(_i(m), _i(n), _ptr(a), _i(ia), _i(ja), _desc(desca), _ptr(b), _i(ib), _i(jb), _desc(descb), _i(ictxt))


# ------------------------------------------------------------------ Cholesky
def pdpotrf(uplo, n, a, ia, ja, desca):
    info = C.c_int()
    lib().pdpotrf_(uplo.encode(), _i(n), _ptr(a), _i(ia), _i(ja), _desc(desca), C.byref(info))
    return info.value


def pdpotrs(uplo, n, nrhs, a, ia, ja, desca, b, ib, jb, descb):
    info = C.c_int()
    lib().pdpotrs_(uplo.encode(), _i(n), _i(nrhs), _ptr(a), _i(ia), _i(ja), _desc(desca), _ptr(b), _i(ib), _i(jb), _desc(descb), C.byref(info))
    return info.value


def pdposv(uplo, n, nrhs, a, ia, ja, desca, b, ib, jb, descb):
    info = C.c_int()
    lib().pdposv_(uplo.encode(), _i(n), _i(nrhs), _ptr(a), _i(ia), _i(ja), _desc(desca), _ptr(b), _i(ib), _i(jb), _desc(descb), C.byref(info))
    return info.value


# ------------------------------------------------------------------ inverse
def pdgetri(n, a, ia, ja, desca, ipiv, lwork=None, liwork=None):
    """SRC/pdgetri.f: sub(A) <- inv(sub(A)) from the factors of PDGETRF; returns INFO."""
    info = C.c_int()
    w1, iw1 = np.zeros(1), np.zeros(1, np.int32)
    lib().pdgetri_(_i(n), _ptr(a), _i(ia), _i(ja), _desc(desca), _ipiv_ptr(ipiv), _dptr(w1), _i(-1), _ipiv_ptr(iw1), _i(-1), C.byref(info))
    if info.value != 0:
        return info.value
    lw = int(w1[0]) if lwork is None else lwork
    liw = int(iw1[0]) if liwork is None else liwork
    work, iwork = np.zeros(max(1, lw)), np.zeros(max(1, liw), np.int32)
    lib().pdgetri_(_i(n), _ptr(a), _i(ia), _i(ja), _desc(desca), _ipiv_ptr(ipiv), _dptr(work), _i(lw), _ipiv_ptr(iwork), _i(liw), C.byref(info))
    return info.value


# ------------------------------------------------------------------ PBLAS entry points
def pdgemm(transa, transb, m, n, k, alpha, a, ia, ja, desca, b, ib, jb, descb, beta, c, ic, jc, descc):
    lib().pdgemm_(transa.encode(), transb.encode(), _i(m), _i(n), _i(k), _d(alpha), _ptr(a), _i(ia), _i(ja), _desc(desca), _ptr(b), _i(ib),
                  _i(jb), _desc(descb), _d(beta), _ptr(c), _i(ic), _i(jc), _desc(descc))


def pdtrsm(side, uplo, transa, diag, m, n, alpha, a, ia, ja, desca, b, ib, jb, descb):
    lib().pdtrsm_(side.encode(), uplo.encode(), transa.encode(), diag.encode(), _i(m), _i(n), _d(alpha), _ptr(a), _i(ia), _i(ja),
                  _desc(desca), _ptr(b), _i(ib), _i(jb), _desc(descb))


def pdtran(m, n, alpha, a, ia, ja, desca, beta, c, ic, jc, descc):
    lib().pdtran_(_i(m), _i(n), _d(alpha), _ptr(a), _i(ia), _i(ja), _desc(desca), _d(beta), _ptr(c), _i(ic), _i(jc), _desc(descc))


def pdgetrs_l3(trans, n, nrhs, a, ia, ja, desca, ipiv, b, ib, jb, descb):
    """PDGETRS forced through the level-3 (many right-hand sides) path."""
    info = C.c_int()
    lib().slb200_pdgetrs_l3(trans.encode(), _i(n), _i(nrhs), _ptr(a), _i(ia), _i(ja), _desc(desca), _ipiv_ptr(ipiv), _ptr(b), _i(ib), _i(jb),
                            _desc(descb), C.byref(info))
    return info.value


# ------------------------------------------------------------------ test-driver helpers
def pdmatgen(ictxt, m, n, mb, nb, a, lda, iarow=0, iacol=0, iseed=100):
    lib().slb200_pdmatgen(_i(ictxt), _i(m), _i(n), _i(mb), _i(nb), _ptr(a), _i(lda), _i(iarow), _i(iacol), _i(iseed))


def matgen64(ictxt, m, n, mb, nb, a, lda, seed, iarow=0, iacol=0):
    lib().slb200_matgen64(_i(ictxt), C.byref(C.c_int64(m)), C.byref(C.c_int64(n)), _i(mb), _i(nb), _ptr(a),
                          C.byref(C.c_int64(lda)), _i(iarow), _i(iacol), C.byref(C.c_uint64(seed)))


def zmatgen64(ictxt, m, n, mb, nb, a, lda, seed, iarow=0, iacol=0):
    lib().slb200_zmatgen64(_i(ictxt), C.byref(C.c_int64(m)), C.byref(C.c_int64(n)), _i(mb), _i(nb), _ptr(a),
                           C.byref(C.c_int64(lda)), _i(iarow), _i(iacol), C.byref(C.c_uint64(seed)))


def pdlaschk(ictxt, n, nrhs, x, descx, desca, aseed, bseed, gen=64):
    return lib().slb200_pdlaschk(_i(ictxt), _i(n), _i(nrhs), _ptr(x), _desc(descx), _desc(desca),
                                 C.byref(C.c_uint64(aseed)), C.byref(C.c_uint64(bseed)), _i(gen))


# ------------------------------------------------------------------ runtime controls
def set_option(key, value):
    lib().slb200_set_option(key.encode(), C.c_int64(int(value)))


def get_counter(key):
    return int(lib().slb200_get_counter(key.encode()))


def reset_counters():
    lib().slb200_reset_counters()


def last_factor_ms():
    return float(lib().slb200_last_factor_ms())


def last_solve_ms():
    return float(lib().slb200_last_solve_ms())


def last_update():
    """(ms, flops, launches) of the trailing-update kernel summed over the last PDGETRF on this rank."""
    L = lib()
    return float(L.slb200_last_update_ms()), float(L.slb200_last_update_flops()), int(L.slb200_last_update_launches())
