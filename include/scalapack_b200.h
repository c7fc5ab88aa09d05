/*
 * scalapack_b200.h -- C-ABI of the B200-native distributed dense LU path.
 *
 * Every entry point below is a drop-in for the reference ScaLAPACK symbol of
 * the same name (Fortran-77 calling convention: lower case + trailing
 * underscore, all arguments by reference, INTEGER = 32-bit int, CHARACTER as
 * char* whose first letter is significant; hidden string lengths are ignored
 * exactly as the reference's C code ignores them).  Reference file:line of the
 * interface each symbol replaces is cited next to it (paths relative to the
 * reference source root).
 *
 * One process drives one GPU (device = $LOCAL_RANK, else rank % device count).
 * Process bootstrap replaces MPI_Init: rank/size come from RANK/WORLD_SIZE
 * (torchrun), OMPI_COMM_WORLD_*, or PMI_*; rendezvous is TCP on
 * MASTER_ADDR : MASTER_PORT + SLB200_PORT_OFFSET (default 23).
 *
 * Matrix arguments (A, B) may be HOST pointers (staged through device memory,
 * the drop-in case for an unmodified Fortran caller) or DEVICE pointers
 * (factored in place in HBM).  IPIV / descriptors / scalars are host memory.
 */
#ifndef SCALAPACK_B200_H
#define SCALAPACK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } slb200_z;   /* COMPLEX*16 */

/* ---- BLACS setup API (BLACS/SRC/<name>.c) ------------------------------------ */
void blacs_pinfo_(int *mypnum, int *nprocs);                                  /* BLACS/SRC/blacs_pinfo_.c:3-28 */
void blacs_get_(const int *ictxt, const int *what, int *val);                 /* BLACS/SRC/blacs_get_.c */
void blacs_set_(const int *ictxt, const int *what, const int *val);           /* BLACS/SRC/blacs_set_.c (accepted, ignored) */
void blacs_gridinit_(int *ictxt, const char *order, const int *nprow, const int *npcol);   /* BLACS/SRC/blacs_init_.c:3-40 */
void blacs_gridmap_(int *ictxt, const int *usermap, const int *ldumap, const int *nprow, const int *npcol); /* blacs_map_.c:84-141 */
void blacs_gridinfo_(const int *ictxt, int *nprow, int *npcol, int *myrow, int *mycol);    /* BLACS/SRC/blacs_info_.c */
void blacs_gridexit_(const int *ictxt);                                       /* BLACS/SRC/blacs_grid_.c */
void blacs_exit_(const int *notdone);                                         /* BLACS/SRC/blacs_exit_.c */
void blacs_abort_(const int *ictxt, const int *errnum);                       /* BLACS/SRC/blacs_abort_.c */
void blacs_barrier_(const int *ictxt, const char *scope);                     /* BLACS/SRC/blacs_barr_.c:16-26 */
int  blacs_pnum_(const int *ictxt, const int *prow, const int *pcol);         /* BLACS/SRC/blacs_pnum_.c */
void blacs_pcoord_(const int *ictxt, const int *pnum, int *prow, int *pcol);  /* BLACS/SRC/blacs_pcoord_.c */
/* C twins (BLACS/SRC/<name>.c compile both bindings from one file, dgebs2d_.c:3-8) */
void Cblacs_pinfo(int *mypnum, int *nprocs);
void Cblacs_get(int ictxt, int what, int *val);
void Cblacs_gridinit(int *ictxt, const char *order, int nprow, int npcol);
void Cblacs_gridinfo(int ictxt, int *nprow, int *npcol, int *myrow, int *mycol);
void Cblacs_gridexit(int ictxt);
void Cblacs_exit(int notdone);
void Cblacs_barrier(int ictxt, const char *scope);
int  Cblacs_pnum(int ictxt, int prow, int pcol);
void Cblacs_pcoord(int ictxt, int pnum, int *prow, int *pcol);
/* the two integer combines the LU path itself calls directly */
void igamn2d_(const int *ictxt, const char *scope, const char *top, const int *m, const int *n, int *a,
              const int *lda, int *ra, int *ca, const int *rcflag, const int *rdest, const int *cdest); /* BLACS/SRC/igamn2d_.c */
void igamx2d_(const int *ictxt, const char *scope, const char *top, const int *m, const int *n, int *a,
              const int *lda, int *ra, int *ca, const int *rcflag, const int *rdest, const int *cdest); /* BLACS/SRC/igamx2d_.c */

/* ---- TOOLS (TOOLS/<name>.f, SL_init.f) ---------------------------------------- */
void sl_init_(int *ictxt, const int *nprow, const int *npcol);                /* TOOLS/SL_init.f */
void descinit_(int *desc, const int *m, const int *n, const int *mb, const int *nb, const int *irsrc,
               const int *icsrc, const int *ictxt, const int *lld, int *info);                  /* TOOLS/descinit.f:1-2 */
void descset_(int *desc, const int *m, const int *n, const int *mb, const int *nb, const int *irsrc,
              const int *icsrc, const int *ictxt, const int *lld);                               /* TOOLS/descset.f */
int  numroc_(const int *n, const int *nb, const int *iproc, const int *isrcproc, const int *nprocs);   /* TOOLS/numroc.f */
int  indxg2p_(const int *indxglob, const int *nb, const int *iproc, const int *isrcproc, const int *nprocs); /* TOOLS/indxg2p.f */
int  indxg2l_(const int *indxglob, const int *nb, const int *iproc, const int *isrcproc, const int *nprocs); /* TOOLS/indxg2l.f */
int  indxl2g_(const int *indxloc, const int *nb, const int *iproc, const int *isrcproc, const int *nprocs);  /* TOOLS/indxl2g.f */
void infog2l_(const int *grindx, const int *gcindx, const int *desc, const int *nprow, const int *npcol,
              const int *myrow, const int *mycol, int *lrindx, int *lcindx, int *rsrc, int *csrc);  /* TOOLS/infog2l.f */
int  iceil_(const int *inum, const int *idenom);                              /* TOOLS/iceil.f */
int  ilcm_(const int *m, const int *n);                                       /* TOOLS/ilcm.f */
void chk1mat_(const int *ma, const int *mapos0, const int *na, const int *napos0, const int *ia, const int *ja,
              const int *desca, const int *descapos0, int *info);             /* TOOLS/chk1mat.f:1 */
void pchk1mat_(const int *ma, const int *mapos0, const int *na, const int *napos0, const int *ia, const int *ja,
               const int *desca, const int *descapos0, const int *nextra, const int *ex, const int *expos,
               int *info);                                                     /* TOOLS/pchkxmat.f:1 */
void pchk2mat_(const int *ma, const int *mapos0, const int *na, const int *napos0, const int *ia, const int *ja,
               const int *desca, const int *descapos0, const int *mb, const int *mbpos0, const int *nb,
               const int *nbpos0, const int *ib, const int *jb, const int *descb, const int *descbpos0,
               const int *nextra, const int *ex, const int *expos, int *info); /* TOOLS/pchkxmat.f:173 */
void pxerbla_(const int *ictxt, const char *srname, const int *info);         /* PBLAS/SRC/PTZBLAS/pxerbla.f:53-58 */
void pb_topget_(const int *ictxt, const char *op, const char *scope, char *top);   /* PBLAS/SRC/PTOOLS/PB_Ctop.c:76-141 */
void pb_topset_(const int *ictxt, const char *op, const char *scope, const char *top);

/* ---- the hot path: distributed LU factor / solve ------------------------- */
void pdgetrf_(const int *m, const int *n, double *a, const int *ia, const int *ja, const int *desca,
              int *ipiv, int *info);                                          /* SRC/pdgetrf.f:1 */
void pdgetrs_(const char *trans, const int *n, const int *nrhs, const double *a, const int *ia, const int *ja,
              const int *desca, const int *ipiv, double *b, const int *ib, const int *jb, const int *descb,
              int *info);                                                     /* SRC/pdgetrs.f:1-2 */
void pdgesv_(const int *n, const int *nrhs, double *a, const int *ia, const int *ja, const int *desca,
             int *ipiv, double *b, const int *ib, const int *jb, const int *descb, int *info);   /* SRC/pdgesv.f:1-2 */
void pzgetrf_(const int *m, const int *n, slb200_z *a, const int *ia, const int *ja, const int *desca,
              int *ipiv, int *info);                                          /* SRC/pzgetrf.f:1 */
void pzgetrs_(const char *trans, const int *n, const int *nrhs, const slb200_z *a, const int *ia, const int *ja,
              const int *desca, const int *ipiv, slb200_z *b, const int *ib, const int *jb, const int *descb,
              int *info);                                                     /* SRC/pzgetrs.f:1-2 */
void pzgesv_(const int *n, const int *nrhs, slb200_z *a, const int *ia, const int *ja, const int *desca,
             int *ipiv, slb200_z *b, const int *ib, const int *jb, const int *descb, int *info); /* SRC/pzgesv.f:1-2 */

/* ---- around the factors: norms, equilibration, condition estimate, refinement, expert driver ----
 * WORK / IWORK keep the reference's workspace protocol (LWORK = -1 queries the minimal sizes into WORK(1) / IWORK(1));
 * the library allocates its own device scratch and does not touch them otherwise. */
double pdlange_(const char *norm, const int *m, const int *n, const double *a, const int *ia, const int *ja,
                const int *desca, double *work);                              /* SRC/pdlange.f:1-2 */
void pdgeequ_(const int *m, const int *n, const double *a, const int *ia, const int *ja, const int *desca, double *r,
              double *c, double *rowcnd, double *colcnd, double *amax, int *info);             /* SRC/pdgeequ.f:1-2 */
void pdlaqge_(const int *m, const int *n, double *a, const int *ia, const int *ja, const int *desca, const double *r,
              const double *c, const double *rowcnd, const double *colcnd, const double *amax, char *equed); /* SRC/pdlaqge.f:1-2 */
void pdgecon_(const char *norm, const int *n, const double *a, const int *ia, const int *ja, const int *desca,
              const double *anorm, double *rcond, double *work, const int *lwork, int *iwork, const int *liwork,
              int *info);                                                     /* SRC/pdgecon.f:1-2 */
void pdgerfs_(const char *trans, const int *n, const int *nrhs, const double *a, const int *ia, const int *ja,
              const int *desca, const double *af, const int *iaf, const int *jaf, const int *descaf, const int *ipiv,
              const double *b, const int *ib, const int *jb, const int *descb, double *x, const int *ix, const int *jx,
              const int *descx, double *ferr, double *berr, double *work, const int *lwork, int *iwork,
              const int *liwork, int *info);                                  /* SRC/pdgerfs.f:1-4 */
void pdgesvx_(const char *fact, const char *trans, const int *n, const int *nrhs, double *a, const int *ia, const int *ja,
              const int *desca, double *af, const int *iaf, const int *jaf, const int *descaf, int *ipiv, char *equed,
              double *r, double *c, double *b, const int *ib, const int *jb, const int *descb, double *x, const int *ix,
              const int *jx, const int *descx, double *rcond, double *ferr, double *berr, double *work, const int *lwork,
              int *iwork, const int *liwork, int *info);                      /* SRC/pdgesvx.f:1-5 */

/* ---- redistribution between block-cyclic layouts / grids.  ictxt: a context containing every process of
 * both grids; all of its processes call; a process outside A's (B's) grid passes DESCA(CTXT_) (DESCB(CTXT_)) = -1. */
void pdgemr2d_(const int *m, const int *n, const double *a, const int *ia, const int *ja, const int *desca, double *b,
               const int *ib, const int *jb, const int *descb, const int *ictxt);             /* REDIST/SRC/pdgemr.c:253-260 */
void pzgemr2d_(const int *m, const int *n, const slb200_z *a, const int *ia, const int *ja, const int *desca, slb200_z *b,
               const int *ib, const int *jb, const int *descb, const int *ictxt);             /* REDIST/SRC/pzgemr.c */
void Cpdgemr2d(int m, int n, const double *a, int ia, int ja, const int *desca, double *b, int ib, int jb,
               const int *descb, int gcontext);                                               /* REDIST/SRC/pdgemr.c:286-296 */
void Cpzgemr2d(int m, int n, const slb200_z *a, int ia, int ja, const int *desca, slb200_z *b, int ib, int jb,
               const int *descb, int gcontext);

/* ---- Cholesky: only the UPLO triangle of sub(A) is referenced / written ---- */
void pdpotrf_(const char *uplo, const int *n, double *a, const int *ia, const int *ja, const int *desca, int *info);   /* SRC/pdpotrf.f:1 */
void pdpotrs_(const char *uplo, const int *n, const int *nrhs, const double *a, const int *ia, const int *ja, const int *desca,
              double *b, const int *ib, const int *jb, const int *descb, int *info);                                /* SRC/pdpotrs.f:1-2 */
void pdposv_(const char *uplo, const int *n, const int *nrhs, double *a, const int *ia, const int *ja, const int *desca,
             double *b, const int *ib, const int *jb, const int *descb, int *info);                                 /* SRC/pdposv.f:1-2 */

/* ---- inverse from the factors of PDGETRF; WORK / IWORK: the reference's workspace protocol ---- */
void pdgetri_(const int *n, double *a, const int *ia, const int *ja, const int *desca, const int *ipiv, double *work,
              const int *lwork, int *iwork, const int *liwork, int *info);                                         /* SRC/pdgetri.f:1-2 */

/* ---- standalone PBLAS entry points over the LU's kernels: any alignment / blocking / transposition ---- */
void pdgemm_(const char *transa, const char *transb, const int *m, const int *n, const int *k, const double *alpha,
             const double *a, const int *ia, const int *ja, const int *desca, const double *b, const int *ib, const int *jb,
             const int *descb, const double *beta, double *c, const int *ic, const int *jc, const int *descc);      /* PBLAS/SRC/pdgemm_.c:21-33 */
void pdtrsm_(const char *side, const char *uplo, const char *transa, const char *diag, const int *m, const int *n,
             const double *alpha, const double *a, const int *ia, const int *ja, const int *desca, double *b,
             const int *ib, const int *jb, const int *descb);                                                       /* PBLAS/SRC/pdtrsm_.c:21-31 */
void pdtran_(const int *m, const int *n, const double *alpha, const double *a, const int *ia, const int *ja,
             const int *desca, const double *beta, double *c, const int *ic, const int *jc, const int *descc);      /* PBLAS/SRC/pdtran_.c:21-29 */

/* ---- test-driver helpers (TESTING/traditional/LIN, run on the device) ---- */
/* PDMATGEN 'N','N' closed form into a local block-cyclic array (pdmatgen.f:448-510);
 * a may be host or device. */
void slb200_pdmatgen(const int *ictxt, const int *m, const int *n, const int *mb, const int *nb, double *a,
                     const int *lda, const int *iarow, const int *iacol, const int *iseed);
/* 64-bit LCG test matrix for N beyond PDMATGEN's 2^31 period (not in the reference). */
void slb200_matgen64(const int *ictxt, const int64_t *m, const int64_t *n, const int *mb, const int *nb, double *a,
                     const int64_t *lda, const int *iarow, const int *iacol, const uint64_t *seed);
void slb200_zmatgen64(const int *ictxt, const int64_t *m, const int64_t *n, const int *mb, const int *nb, slb200_z *a,
                      const int64_t *lda, const int *iarow, const int *iacol, const uint64_t *seed);
/* Solve residual of pdlaschk.f:187,296 with A and B regenerated on the device from the 64-bit
 * generator (gen=64) or PDMATGEN (gen=31).  x: local block-cyclic solution (host or device). */
double slb200_pdlaschk(const int *ictxt, const int *n, const int *nrhs, const double *x, const int *descx,
                       const int *desca, const uint64_t *aseed, const uint64_t *bseed, const int *gen);

/* PDGETRS through the level-3 distributed path (block-cyclic copy of sub(B), tensor-core sweeps) whatever NRHS is; pdgetrs_ itself
 * switches to it when NRHS exceeds the option "solve_l3_min_nrhs" (default 64).  Real; TRANS = 'N' / 'T'. */
void slb200_pdgetrs_l3(const char *trans, const int *n, const int *nrhs, const double *a, const int *ia, const int *ja,
                       const int *desca, const int *ipiv, double *b, const int *ib, const int *jb, const int *descb, int *info);

/* ---- runtime controls / introspection (not in the reference) ------------- */
int  slb200_device(void);                  /* CUDA device this process drives, -1 if none */
int  slb200_has_cuda(void);                /* 1 when a usable sm_100 device is present     */
const char *slb200_version(void);
void slb200_set_option(const char *key, int64_t value);   /* "lookahead", "verbose", "panel_width", ...;
                                                             "lacon_keep_estimate" = 1: PDGECON / PDGERFS / PDGESVX use LAPACK's estimator instead of
                                                             what SRC/pdlacon.f:188-189 makes the reference return */
int64_t slb200_get_counter(const char *key);             /* "kernel_launches", "h2d_bytes", "d2h_bytes", ... */
void slb200_reset_counters(void);
/* Device-time (ms) of the last pdgetrf_/pdgetrs_ call on this rank, measured with CUDA events
 * on the library's own streams (host staging excluded). */
double slb200_last_factor_ms(void);
double slb200_last_solve_ms(void);
/* per-kernel event timing of the dominant kernel (trailing update) in the last pdgetrf_ */
double slb200_last_update_ms(void);
double slb200_last_update_flops(void);
int64_t slb200_last_update_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* SCALAPACK_B200_H */
