! scalapack_b200_iface.f90 -- ISO_C_BINDING view of libscalapack_b200.so for Fortran callers.
!
! An UNMODIFIED Fortran-77 ScaLAPACK program needs none of this: the library exports the classic
! mangled symbols (pdgetrf_, blacs_gridinit_, descinit_, numroc_, ...), so `CALL PDGETRF(...)` links
! directly.  This module is for callers that want explicit interfaces, and for the two extras the
! reference does not have (device-resident operands and the on-device test generators).
! NOTE: no Fortran compiler exists in the build image; this file is syntax-reviewed only.
module scalapack_b200_iface
  use iso_c_binding
  implicit none
  interface
     subroutine pdgetrf(m, n, a, ia, ja, desca, ipiv, info) bind(C, name="pdgetrf_")   ! SRC/pdgetrf.f:1
       import :: c_int, c_double
       integer(c_int), intent(in) :: m, n, ia, ja, desca(9)
       real(c_double), intent(inout) :: a(*)
       integer(c_int), intent(out) :: ipiv(*), info
     end subroutine
     subroutine pdgetrs(trans, n, nrhs, a, ia, ja, desca, ipiv, b, ib, jb, descb, info) bind(C, name="pdgetrs_")  ! SRC/pdgetrs.f:1
       import :: c_int, c_double, c_char
       character(kind=c_char), intent(in) :: trans
       integer(c_int), intent(in) :: n, nrhs, ia, ja, desca(9), ipiv(*), ib, jb, descb(9)
       real(c_double), intent(in) :: a(*)
       real(c_double), intent(inout) :: b(*)
       integer(c_int), intent(out) :: info
     end subroutine
     subroutine pdgesv(n, nrhs, a, ia, ja, desca, ipiv, b, ib, jb, descb, info) bind(C, name="pdgesv_")           ! SRC/pdgesv.f:1
       import :: c_int, c_double
       integer(c_int), intent(in) :: n, nrhs, ia, ja, desca(9), ib, jb, descb(9)
       real(c_double), intent(inout) :: a(*), b(*)
       integer(c_int), intent(out) :: ipiv(*), info
     end subroutine
     subroutine pzgetrf(m, n, a, ia, ja, desca, ipiv, info) bind(C, name="pzgetrf_")   ! SRC/pzgetrf.f:1
       import :: c_int, c_double_complex
       integer(c_int), intent(in) :: m, n, ia, ja, desca(9)
       complex(c_double_complex), intent(inout) :: a(*)
       integer(c_int), intent(out) :: ipiv(*), info
     end subroutine
     subroutine blacs_pinfo(mypnum, nprocs) bind(C, name="blacs_pinfo_")
       import :: c_int
       integer(c_int), intent(out) :: mypnum, nprocs
     end subroutine
     subroutine blacs_get(ictxt, what, val) bind(C, name="blacs_get_")
       import :: c_int
       integer(c_int), intent(in) :: ictxt, what
       integer(c_int), intent(out) :: val
     end subroutine
     subroutine blacs_gridinit(ictxt, order, nprow, npcol) bind(C, name="blacs_gridinit_")
       import :: c_int, c_char
       integer(c_int), intent(inout) :: ictxt
       character(kind=c_char), intent(in) :: order
       integer(c_int), intent(in) :: nprow, npcol
     end subroutine
     subroutine blacs_gridinfo(ictxt, nprow, npcol, myrow, mycol) bind(C, name="blacs_gridinfo_")
       import :: c_int
       integer(c_int), intent(in) :: ictxt
       integer(c_int), intent(out) :: nprow, npcol, myrow, mycol
     end subroutine
     subroutine blacs_gridexit(ictxt) bind(C, name="blacs_gridexit_")
       import :: c_int
       integer(c_int), intent(in) :: ictxt
     end subroutine
     subroutine blacs_exit(notdone) bind(C, name="blacs_exit_")
       import :: c_int
       integer(c_int), intent(in) :: notdone
     end subroutine
     subroutine descinit(desc, m, n, mb, nb, irsrc, icsrc, ictxt, lld, info) bind(C, name="descinit_")   ! TOOLS/descinit.f:1
       import :: c_int
       integer(c_int), intent(out) :: desc(9), info
       integer(c_int), intent(in) :: m, n, mb, nb, irsrc, icsrc, ictxt, lld
     end subroutine
     integer(c_int) function numroc(n, nb, iproc, isrcproc, nprocs) bind(C, name="numroc_")             ! TOOLS/numroc.f
       import :: c_int
       integer(c_int), intent(in) :: n, nb, iproc, isrcproc, nprocs
     end function
     ! extras: device-side 64-bit test generator and solve-residual check (TESTING/traditional/LIN analogues)
     subroutine slb200_matgen64(ictxt, m, n, mb, nb, a, lda, iarow, iacol, seed) bind(C, name="slb200_matgen64")
       import :: c_int, c_int64_t, c_double
       integer(c_int), intent(in) :: ictxt, mb, nb, iarow, iacol
       integer(c_int64_t), intent(in) :: m, n, lda, seed
       real(c_double), intent(out) :: a(*)
     end subroutine
     real(c_double) function slb200_last_factor_ms() bind(C, name="slb200_last_factor_ms")
       import :: c_double
     end function
     ! ---- around the factors (SURVEY 8f row 1) ----
     function pdlange(norm, m, n, a, ia, ja, desca, work) bind(C, name="pdlange_") result(v)          ! SRC/pdlange.f:1
       import :: c_int, c_double, c_char
       character(kind=c_char), intent(in) :: norm
       integer(c_int), intent(in) :: m, n, ia, ja, desca(9)
       real(c_double), intent(in) :: a(*)
       real(c_double) :: work(*), v
     end function
     subroutine pdgecon(norm, n, a, ia, ja, desca, anorm, rcond, work, lwork, iwork, liwork, info) bind(C, name="pdgecon_")   ! SRC/pdgecon.f:1
       import :: c_int, c_double, c_char
       character(kind=c_char), intent(in) :: norm
       integer(c_int), intent(in) :: n, ia, ja, desca(9), lwork, liwork
       real(c_double), intent(in) :: a(*), anorm
       real(c_double), intent(out) :: rcond
       real(c_double) :: work(*)
       integer(c_int) :: iwork(*)
       integer(c_int), intent(out) :: info
     end subroutine
     subroutine pdgerfs(trans, n, nrhs, a, ia, ja, desca, af, iaf, jaf, descaf, ipiv, b, ib, jb, descb, x, ix, jx, descx, &
                        ferr, berr, work, lwork, iwork, liwork, info) bind(C, name="pdgerfs_")                               ! SRC/pdgerfs.f:1
       import :: c_int, c_double, c_char
       character(kind=c_char), intent(in) :: trans
       integer(c_int), intent(in) :: n, nrhs, ia, ja, desca(9), iaf, jaf, descaf(9), ipiv(*), ib, jb, descb(9), ix, jx, descx(9), lwork, liwork
       real(c_double), intent(in) :: a(*), af(*), b(*)
       real(c_double), intent(inout) :: x(*)
       real(c_double), intent(out) :: ferr(*), berr(*)
       real(c_double) :: work(*)
       integer(c_int) :: iwork(*)
       integer(c_int), intent(out) :: info
     end subroutine
     subroutine pdgesvx(fact, trans, n, nrhs, a, ia, ja, desca, af, iaf, jaf, descaf, ipiv, equed, r, c, b, ib, jb, descb, &
                        x, ix, jx, descx, rcond, ferr, berr, work, lwork, iwork, liwork, info) bind(C, name="pdgesvx_")      ! SRC/pdgesvx.f:1
       import :: c_int, c_double, c_char
       character(kind=c_char), intent(in) :: fact, trans
       character(kind=c_char), intent(inout) :: equed
       integer(c_int), intent(in) :: n, nrhs, ia, ja, desca(9), iaf, jaf, descaf(9), ib, jb, descb(9), ix, jx, descx(9), lwork, liwork
       real(c_double), intent(inout) :: a(*), af(*), r(*), c(*), b(*), x(*)
       integer(c_int), intent(inout) :: ipiv(*)
       real(c_double), intent(out) :: rcond, ferr(*), berr(*)
       real(c_double) :: work(*)
       integer(c_int) :: iwork(*)
       integer(c_int), intent(out) :: info
     end subroutine
     ! ---- redistribution, Cholesky, inverse, PBLAS entry points (SURVEY 8f rows 2-4) ----
     subroutine pdgemr2d(m, n, a, ia, ja, desca, b, ib, jb, descb, ictxt) bind(C, name="pdgemr2d_")    ! REDIST/SRC/pdgemr.c:253
       import :: c_int, c_double
       integer(c_int), intent(in) :: m, n, ia, ja, desca(9), ib, jb, descb(9), ictxt
       real(c_double), intent(in) :: a(*)
       real(c_double), intent(inout) :: b(*)
     end subroutine
     subroutine pdpotrf(uplo, n, a, ia, ja, desca, info) bind(C, name="pdpotrf_")                      ! SRC/pdpotrf.f:1
       import :: c_int, c_double, c_char
       character(kind=c_char), intent(in) :: uplo
       integer(c_int), intent(in) :: n, ia, ja, desca(9)
       real(c_double), intent(inout) :: a(*)
       integer(c_int), intent(out) :: info
     end subroutine
     subroutine pdpotrs(uplo, n, nrhs, a, ia, ja, desca, b, ib, jb, descb, info) bind(C, name="pdpotrs_")   ! SRC/pdpotrs.f:1
       import :: c_int, c_double, c_char
       character(kind=c_char), intent(in) :: uplo
       integer(c_int), intent(in) :: n, nrhs, ia, ja, desca(9), ib, jb, descb(9)
       real(c_double), intent(in) :: a(*)
       real(c_double), intent(inout) :: b(*)
       integer(c_int), intent(out) :: info
     end subroutine
     subroutine pdgetri(n, a, ia, ja, desca, ipiv, work, lwork, iwork, liwork, info) bind(C, name="pdgetri_")   ! SRC/pdgetri.f:1
       import :: c_int, c_double
       integer(c_int), intent(in) :: n, ia, ja, desca(9), ipiv(*), lwork, liwork
       real(c_double), intent(inout) :: a(*)
       real(c_double) :: work(*)
       integer(c_int) :: iwork(*)
       integer(c_int), intent(out) :: info
     end subroutine
     subroutine pdgemm(transa, transb, m, n, k, alpha, a, ia, ja, desca, b, ib, jb, descb, beta, c, ic, jc, descc) bind(C, name="pdgemm_")   ! PBLAS/SRC/pdgemm_.c:21
       import :: c_int, c_double, c_char
       character(kind=c_char), intent(in) :: transa, transb
       integer(c_int), intent(in) :: m, n, k, ia, ja, desca(9), ib, jb, descb(9), ic, jc, descc(9)
       real(c_double), intent(in) :: alpha, beta, a(*), b(*)
       real(c_double), intent(inout) :: c(*)
     end subroutine
     subroutine pdtrsm(side, uplo, transa, diag, m, n, alpha, a, ia, ja, desca, b, ib, jb, descb) bind(C, name="pdtrsm_")   ! PBLAS/SRC/pdtrsm_.c:21
       import :: c_int, c_double, c_char
       character(kind=c_char), intent(in) :: side, uplo, transa, diag
       integer(c_int), intent(in) :: m, n, ia, ja, desca(9), ib, jb, descb(9)
       real(c_double), intent(in) :: alpha, a(*)
       real(c_double), intent(inout) :: b(*)
     end subroutine
  end interface
end module scalapack_b200_iface
