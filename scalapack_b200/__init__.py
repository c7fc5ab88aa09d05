"""scalapack_b200 -- B200-native distributed dense LU behind the ScaLAPACK calling convention.

The product is the C-ABI shared library ``scalapack_b200/lib/libscalapack_b200.so`` (hand-written sm_100a
CUDA + NCCL; declared in ``include/scalapack_b200.h``).  This package is the thin Python mirror of the
reference interface (same routine names, argument order and INFO conventions as the Fortran entry points
``PDGETRF / PDGETRS / PDGESV / PZ*`` and the BLACS / TOOLS setup calls) used by the tests and the bench.
There is no CPU fallback: compute entry points abort if the CUDA library or a GPU is missing.
"""
from .api import (  # noqa: F401
    lib, have_library, has_cuda, device,
    blacs_pinfo, blacs_get, blacs_gridinit, blacs_gridinfo, blacs_gridexit, blacs_exit, blacs_barrier,
    blacs_pnum, blacs_pcoord, sl_init,
    numroc, indxg2p, indxg2l, indxl2g, infog2l, descinit, iceil, ilcm, chk1mat,
    pdgetrf, pdgetrs, pdgesv, pzgetrf, pzgetrs, pzgesv,
    pdlange, pdgeequ, pdlaqge, pdgecon, pdgerfs, pdgesvx, pdgemr2d, pzgemr2d,
    pdpotrf, pdpotrs, pdposv, pdgetri, pdgemm, pdtrsm, pdtran, pdgetrs_l3,
    pdmatgen, matgen64, zmatgen64, pdlaschk,
    set_option, get_counter, reset_counters, last_factor_ms, last_solve_ms, last_update,
    DTYPE_, CTXT_, M_, N_, MB_, NB_, RSRC_, CSRC_, LLD_,
)

__version__ = "0.1.0"
