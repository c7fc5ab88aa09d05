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
  end interface
end module scalapack_b200_iface
