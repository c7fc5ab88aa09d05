// kernels.cuh -- launchers of the hand-written sm_100a kernels of the LU path.
// All matrices are column-major with explicit leading dimensions in ELEMENTS, 64-bit offsets.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace slb {

typedef double2 zcomplex;   // COMPLEX*16 as (re, im)

// ---- trailing update (replaces PDGEMM 'N','N', alpha=-1, beta=1; PBLAS/SRC/PTOOLS/PB_CpgemmAB.c:345) ----
// C[M x N] -= A[M x K] * B[K x N].  FP64 tensor cores (DMMA mma.sync m16n8k4), cp.async 4-stage pipeline.
// chunk > 0: each CTA processes `chunk` tiles and retires (lets a higher-priority stream interleave); 0: persistent.
// flags: GEMM_MAIN = the big trailing update on the main stream: may take the packed-operand kernel (gemm_packed.cu,
// one pack workspace, so one stream only); GEMM_REUSE_A = the previous GEMM_MAIN call packed this very A (same step).
enum { GEMM_MAIN = 1, GEMM_REUSE_A = 2 };
void launch_dgemm_minus(int64_t M, int64_t N, int K, const double *A, int64_t lda, const double *B, int64_t ldb,
                        double *C, int64_t ldc, cudaStream_t s, int chunk = 0, int flags = 0);
bool dgemm_takes_packed(int64_t M, int K, int flags);   // the decision launch_dgemm_minus makes (depends on M, K only)
void launch_dgemm_minus_packed(int64_t M, int64_t N, int K, const double *A, int64_t lda, const double *B, int64_t ldb,
                               double *C, int64_t ldc, cudaStream_t s, int chunk, bool reuse_a);
// complex update: flags = GEMM_MAIN routes large products through the packed real kernel (K doubled, interleaved C;
// gemm_packed.cu), everything else through the 64 x 64 complex DMMA kernel of gemm.cu
void launch_zgemm_minus(int64_t M, int64_t N, int K, const zcomplex *A, int64_t lda, const zcomplex *B, int64_t ldb,
                        zcomplex *C, int64_t ldc, cudaStream_t s, int chunk = 0, int flags = 0);
bool zgemm_takes_packed(int64_t M, int K, int flags);
void launch_zgemm_minus_packed(int64_t M, int64_t N, int K, const zcomplex *A, int64_t lda, const zcomplex *B, int64_t ldb,
                               zcomplex *C, int64_t ldc, cudaStream_t s, int chunk, bool reuse_a);

// ---- U12 triangular solve (replaces PDTRSM 'L','L','N','U'; PBLAS/SRC/PTOOLS/PB_CptrsmAB.c:359-416) ----
// B[jb x n] <- unit_lower(L[jb x jb])^-1 B, blocked: 64-row diagonal solves + DMMA updates.
void launch_dtrsm_llnu(int jb, int64_t n, const double *L, int64_t ldl, double *B, int64_t ldb, cudaStream_t s);
void launch_ztrsm_llnu(int jb, int64_t n, const zcomplex *L, int64_t ldl, zcomplex *B, int64_t ldb, cudaStream_t s);

// ---- panel factorisation (replaces PDGETF2 = PDAMAX+PDSWAP+PDSCAL+PDGER; SRC/pdgetf2.f:207-237) ----
// Describes the m x jb panel as up to 8 row segments (one per process row when the panel of a process
// column has been gathered onto one GPU); virtual row v of segment s is local row seg_lr0[s] + (v -
// seg_v0[s]) of process row seg_prow[s]; global row = block-cyclic map with nb, nprow, rsrc.
struct PanelRowMap {
    int g0;               // global row (0-based) of panel row 0; panel rows are consecutive global rows
    int nb, nprow, rsrc;  // block-cyclic row distribution: only used to order ties like the reference's combine
};
// W: m x jb panel (ld = ldw) in virtual row order.  ipiv_out[j] (j < jb) = 1-based GLOBAL row index chosen
// for panel column j (reference IPIV semantics, SRC/pdgetrf.f:118-121).  *info_out receives the first zero
// pivot column (1-based, relative to the panel) if it was 0 on entry.  work: >= panel_work_bytes().
size_t panel_work_bytes(int jb);
// gmax > 0 caps the number of CTAs (SMs) the leaf kernels use and launches them non-cooperatively: look-ahead mode,
// the panel shares the GPU with the trailing update of the previous step.
void launch_dpanel(int m, int jb, double *W, int64_t ldw, const PanelRowMap &map, int *ipiv_out, int *info_out,
                   int info_offset, void *work, cudaStream_t s, int gmax = 0);
void launch_zpanel(int m, int jb, zcomplex *W, int64_t ldw, const PanelRowMap &map, int *ipiv_out, int *info_out,
                   int info_offset, void *work, cudaStream_t s, int gmax = 0);

// ---- row interchanges (replaces PDLASWP / PDSWAP; SRC/pdlaswp.f:163-182, PBLAS/SRC/pdswap_.c:448-534) ----
// Plan of one block of jb sequential interchanges rows (j0+t) <-> ipiv[t]-1, t = 0..jb-1 (global, 0-based j0):
//   top_src[t]  = global row whose ORIGINAL content ends in top row j0+t
//   out_dst[t]  = global row outside the top block that receives new content (or -1)
//   out_src[t]  = index t' such that original top row j0+t' ends in out_dst[t]
struct SwapPlan { int *top_src, *out_dst, *out_src; };
void launch_swap_plan(int j0, int jb, const int *ipiv_blk, SwapPlan plan, cudaStream_t s);

// Row ownership of the local array: global row g lives on process row (rsrc + g/nb) % nprow at local row
// nb*(g/(nb*nprow)) + g%nb.
struct RowDist { int nb, nprow, myrow, rsrc; int shift; };   // local row = block-cyclic local row - shift

// pack: for local columns [c0, c1): Ubuf[t + (c-c0)*ldu] = A[lrow(top_src[t]) + c*lda] when I own top_src[t]
// (else left untouched), and, on the process row owning the top block, Obuf[t + (c-c0)*ldo] =
// A[lrow(j0 + out_src[t]) + c*lda] for valid out_dst[t].
template <typename T>
void launch_swap_pack(int jb, int j0, SwapPlan plan, RowDist rd, const T *A, int64_t lda, int64_t c0, int64_t c1,
                      T *Ubuf, int64_t ldu, T *Obuf, int64_t ldo, cudaStream_t s);
// unpack: A[lrow(out_dst[t]) + c*lda] = Obuf[t + (c-c0)*ldo] for out_dst[t] that I own.
template <typename T>
void launch_swap_unpack_out(int jb, SwapPlan plan, RowDist rd, T *A, int64_t lda, int64_t c0, int64_t c1,
                            const T *Obuf, int64_t ldo, cudaStream_t s);
// the interchange / copy kernels cap their grids (option swap_grid) so that they can run UNDER the trailing update; a caller
// that runs them alone on a stream lifts the cap for its launches (cap = 0) and restores it with -1
void swap_grid_override(int cap);
// select: U[t + c*ldu] = Cbuf_{owner(top_src[t])}[t + c*ldc] where Cbuf_p = Call + p*stride_p (all-gathered packs)
template <typename T>
void launch_swap_select(int jb, SwapPlan plan, RowDist rd, const T *Call, int64_t ldc, int64_t stride_p, int64_t ncols,
                        T *U, int64_t ldu, cudaStream_t s);
// block-cyclic <-> global row order for a gathered panel: local rows [l0, l0+rows) of process row prow_rel
// (relative to rsrc) <-> rows (global - gshift) of G.  to_global=1: G <- L, else L <- G.
template <typename T>
void launch_rows_bc(int64_t rows, int cols, T *L, int64_t ldl, int64_t l0, T *G, int64_t ldg, int64_t gshift, int nb, int nprow,
                    int prow_rel, int to_global, cudaStream_t s);
// copy a jb x ncols block: dst[i + c*ldd] = src[i + c*lds]
template <typename T>
void launch_copy2d(int64_t rows, int64_t cols, const T *src, int64_t lds, T *dst, int64_t ldd, cudaStream_t s);

// ---- test matrices / checks on the device (TESTING/traditional/LIN) ----
void launch_pdmatgen_local(int m, int n, int mb, int nb, double *a, int64_t lda, int iarow, int iacol, int iseed,
                           int myrow, int mycol, int nprow, int npcol, cudaStream_t s);
void launch_matgen64_local(int64_t m, int64_t n, int mb, int nb, double *a, int64_t lda, int iarow, int iacol,
                           uint64_t seed, int myrow, int mycol, int nprow, int npcol, int is_complex, cudaStream_t s);
// r[lrows] (+)= sum over local columns of A_gen(local) * x ; used by the solve-residual check
void launch_gen_matvec(int64_t n, int nb, uint64_t aseed, int gen, int myrow, int mycol, int nprow, int npcol,
                       const double *xrow /* x for my local columns, nq */, double *r /* np */, double *rowabs /* np */,
                       cudaStream_t s);

// ---- triangular solves of PDGETRS (SRC/pdgetrs.f:255-284) ----
// x[0:kb] = tri(op(Akk))^-1 x[0:kb], nrhs columns.  mode: TRSV_UPPER = the stored triangle is U (non-unit diagonal), else L
// (unit diagonal); TRSV_TRANS / TRSV_CONJ: op = transpose / conjugate transpose.
enum { TRSV_UPPER = 1, TRSV_TRANS = 2, TRSV_CONJ = 4, TRSV_NONUNIT_L = 8 };   // NONUNIT_L: the stored L has a general diagonal (Cholesky)
template <typename T>
void launch_trsv_block(int kb, const T *Akk, int64_t lda, T *X, int64_t ldx, int nrhs, int mode, cudaStream_t s);
// Y[rows] -= A[rows x kb] * X[kb] for nrhs columns (memory-bound GEMV-like)
template <typename T>
void launch_gemv_minus(int64_t rows, int kb, const T *A, int64_t lda, const T *X, int64_t ldx, T *Y, int64_t ldy, int nrhs,
                       cudaStream_t s);
// Y[ncols] -= op(A[kb x ncols])^T * X[kb] for nrhs columns (the transposed solves; conj: A^H)
template <typename T>
void launch_gemvt_minus(int kb, int64_t ncols, const T *A, int64_t lda, const T *X, int64_t ldx, T *Y, int64_t ldy, int nrhs,
                        bool conj, cudaStream_t s);

// ---- micro-benchmarks (roofline denominators) ----
double bench_dmma_peak_tflops(int iters);     // register-resident DMMA loop on all SMs
double bench_dfma_peak_tflops(int iters);     // plain FP64 FMA loop on all SMs
double bench_copy_gbs(size_t bytes);

}  // namespace slb
