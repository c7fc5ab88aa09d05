// pblas.cu -- SURVEY 8(f) row 4, second half: the standalone PBLAS entry points behind which the LU's kernels sit.
//   PDGEMM (PBLAS/SRC/pdgemm_.c): sub(C) <- alpha op(sub(A)) op(sub(B)) + beta sub(C)
//   PDTRSM (PBLAS/SRC/pdtrsm_.c): op(sub(A)) X = alpha sub(B)  or  X op(sub(A)) = alpha sub(B), sub(B) overwritten by X
//   PDTRAN (PBLAS/SRC/pdtran_.c): sub(C) <- beta sub(C) + alpha sub(A)'
//
// The PBLAS accept operands with any alignment, any blocking and any transposition; the kernels want one layout.  So every
// operand is first brought into a WORKING LAYOUT on the caller's grid -- square nb x nb blocks, first block on process (0, 0),
// transposition already applied -- by the redistribution engine of redist.cu (gemr2d_core: index lists on the host, one packed
// NCCL exchange), the arithmetic runs there on the FP64 tensor cores, and the result goes back the same way:
//   PDGEMM = SUMMA over the K blocks: the block column of op(A) along the process rows, the block row of op(B) down the process
//            columns, C -= (-alpha A_k) B_k with the LU's update kernel;
//   PDTRSM = one level-3 triangular sweep (tri_l3_sweep, inverse.cu); a right-hand side solve is the left-hand side solve of
//            the transposed system.
#include "common.h"
#include "dist.h"
#include "kernels.cuh"
#include "launch.h"
#include "lu.h"
#include "ncclw.h"
#include "worklayout.h"

namespace slb {

void tri_l3_sweep(Grid *g, bool upper, bool unit, int N, const double *A, int64_t lld, int nb, int rsrc, int csrc, double *X, int64_t ldx,
                  int64_t nlocx);                                                                                                            // inverse.cu
void getrs_l3_device(Grid *g, int N, const double *A, int64_t lld, int nb, int rsrc, int csrc, double *X, int64_t ldx, int64_t nlocx);      // inverse.cu

namespace {

// M <- M + f S (same shape)
__global__ void __launch_bounds__(256)
axpy_block_kernel(int64_t rows, int64_t cols, double *__restrict__ M, int64_t ld, const double *__restrict__ S, int64_t lds, double f)
{
    const int64_t total = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % rows, j = e / rows;
        M[i + j * ld] += f * S[i + j * lds];
    }
}

bool is_trans(char t) { return t == 'T' || t == 'C'; }
char upc(const char *c) { return (char)(c[0] & ~0x20); }

// C <- C - A B in the working layout (A: m x k, B: k x n, C: m x n, same nb): SUMMA over the block columns of A / block rows of B
void summa_minus(Grid *g, int nb, const Work &A, const Work &B, Work &C)
{
    cudaStream_t s = rt().s_main;
    const int P = g->nprow, Q = g->npcol, K = A.n;
    if (P * Q > 1 && !g->nccl) g->nccl = nccl_create(g);
    NcclComms *nc = g->nccl;
    double *Apan = (double *)workspace("pb_apan", (size_t)nb * (A.mloc > 0 ? A.mloc : 1) * sizeof(double));
    double *Bpan = (double *)workspace("pb_bpan", (size_t)nb * (B.nloc > 0 ? B.nloc : 1) * sizeof(double));
    for (int k0 = 0, kk = 0; k0 < K; k0 += nb, ++kk) {
        const int kw = K - k0 < nb ? K - k0 : nb;
        const int pc = kk % Q, pr = kk % P;
        const int64_t lc = numroc(k0, nb, g->mycol, 0, Q), lr = numroc(k0, nb, g->myrow, 0, P);
        const double *Aop = A.dev + lc * A.ld; int64_t lda = A.ld;
        if (Q > 1) {
            if (A.mloc > 0) {
                if (g->mycol == pc) launch_copy2d<double>(A.mloc, kw, A.dev + lc * A.ld, A.ld, Apan, A.mloc, s);
                nccl_bcast(nc->row, Apan, (size_t)A.mloc * kw, NT_F64, pc, s);
            }
            Aop = Apan; lda = A.mloc;
        }
        const double *Bop = B.dev + lr; int64_t ldb = B.ld;
        if (P > 1) {
            if (B.nloc > 0) {
                if (g->myrow == pr) launch_copy2d<double>(kw, B.nloc, B.dev + lr, B.ld, Bpan, kw, s);
                nccl_bcast(nc->col, Bpan, (size_t)kw * B.nloc, NT_F64, pr, s);
            }
            Bop = Bpan; ldb = kw;
        }
        if (C.mloc > 0 && C.nloc > 0) launch_dgemm_minus(C.mloc, C.nloc, kw, Aop, lda, Bop, ldb, C.dev, C.ld, s);
    }
    SLB_CUDA(cudaStreamSynchronize(s));
}

int working_nb(const int *desc) { const int nb = desc[MB_] < desc[NB_] ? desc[MB_] : desc[NB_]; return nb < 1 ? 1 : (nb > 512 ? 512 : nb); }

bool bad_sub(int m, int n, int i, int j, const int *desc) { return i < 1 || j < 1 || i + m - 1 > desc[M_] || j + n - 1 > desc[N_]; }

}  // namespace

// PDGETRS for MANY right-hand sides (the PB_CptrsmAB case of the reference's PDTRSM): instead of replicating sub(B) on every GPU,
// sub(B) is redistributed -- with the row interchanges applied on the way -- into a block-cyclic copy whose rows are aligned with the
// factors, both sweeps run at level 3 on the tensor cores (tri_l3_sweep), and the solution is redistributed back.
//   'N': X = P B;  X <- U^-1 L^-1 X.       'T': X = B;  X <- L^-T U^-T X on a transposed copy of the factors;  B = P' X.
// Adev: device window of the factors (ld, first block on (rsrc, csrc)); ipg: N pivots, 1-based, relative to sub(A).
void getrs_l3_entry(Grid *g, char trans, int n, int nrhs, const double *Adev, int64_t lda, int nb, int rsrc, int csrc, const int *ipg,
                    double *b, int ib, int jb, const int *descb)
{
    std::vector<int> perm((size_t)n), inv((size_t)n);
    for (int i = 0; i < n; ++i) perm[(size_t)i] = i;
    for (int i = 0; i < n; ++i) { const int p = ipg[i] - 1; if (p != i) { const int t = perm[(size_t)i]; perm[(size_t)i] = perm[(size_t)p]; perm[(size_t)p] = t; } }
    for (int i = 0; i < n; ++i) inv[(size_t)perm[(size_t)i]] = i;
    if (trans == 'N') {
        Work X("pb_C", g, n, nrhs, nb, rsrc);
        X.load(b, ib, jb, descb, false, perm.data());             // row i of P b is row perm[i] of b
        SLB_CUDA(cudaStreamSynchronize(rt().s_main));
        getrs_l3_device(g, n, Adev, lda, nb, rsrc, csrc, X.dev, X.ld, X.nloc);
        X.store(b, ib, jb, descb, false);
    } else {
        const int descf[9] = { 1, g->ctxt, n, n, nb, nb, rsrc, csrc, (int)lda };
        Work F("pb_A", g, n, n, nb), X("pb_C", g, n, nrhs, nb);
        F.load(Adev, 1, 1, descf, true);                           // F = (L \ U)': lower triangle U' (general diagonal), upper L' (unit)
        X.load(b, ib, jb, descb, false);
        SLB_CUDA(cudaStreamSynchronize(rt().s_main));
        tri_l3_sweep(g, false, false, n, F.dev, F.ld, nb, 0, 0, X.dev, X.ld, X.nloc);
        tri_l3_sweep(g, true, true, n, F.dev, F.ld, nb, 0, 0, X.dev, X.ld, X.nloc);
        X.store(b, ib, jb, descb, false, inv.data());             // x = P' z: row perm[i] of x is row i of z
    }
}

// PDPOTRS for many right-hand sides, the same way: A = L L' ('L') or U' U ('U'); the sweep with the stored triangle runs on the
// factor in place, the one with its transpose on a transposed copy.
void potrs_l3_entry(Grid *g, bool upper, int n, int nrhs, const double *Adev, int64_t lda, int nb, int rsrc, int csrc, double *b, int ib, int jb,
                    const int *descb)
{
    const int descf[9] = { 1, g->ctxt, n, n, nb, nb, rsrc, csrc, (int)lda };
    Work F("pb_A", g, n, n, nb);
    F.load(Adev, 1, 1, descf, true);                               // F = A': the other triangle of the factor
    if (!upper) {                                                  // L y = b on A (rows aligned with A), then L' x = y on F (rows from process 0)
        Work X("pb_C", g, n, nrhs, nb, rsrc);
        X.load(b, ib, jb, descb, false);
        SLB_CUDA(cudaStreamSynchronize(rt().s_main));
        tri_l3_sweep(g, false, false, n, Adev, lda, nb, rsrc, csrc, X.dev, X.ld, X.nloc);
        if (rsrc == 0) tri_l3_sweep(g, true, false, n, F.dev, F.ld, nb, 0, 0, X.dev, X.ld, X.nloc);
        else {                                                     // F's rows start on process row 0: move X there for the second sweep
            Work Y("pb_B", g, n, nrhs, nb, 0);
            Y.load(X.dev, 1, 1, X.desc, false);
            SLB_CUDA(cudaStreamSynchronize(rt().s_main));
            tri_l3_sweep(g, true, false, n, F.dev, F.ld, nb, 0, 0, Y.dev, Y.ld, Y.nloc);
            Y.store(b, ib, jb, descb, false);
            return;
        }
        X.store(b, ib, jb, descb, false);
    } else {                                                       // U' y = b on F, then U x = y on A
        Work Y("pb_B", g, n, nrhs, nb, 0);
        Y.load(b, ib, jb, descb, false);
        SLB_CUDA(cudaStreamSynchronize(rt().s_main));
        tri_l3_sweep(g, false, false, n, F.dev, F.ld, nb, 0, 0, Y.dev, Y.ld, Y.nloc);
        Work X("pb_C", g, n, nrhs, nb, rsrc);
        X.load(Y.dev, 1, 1, Y.desc, false);
        SLB_CUDA(cudaStreamSynchronize(rt().s_main));
        tri_l3_sweep(g, true, false, n, Adev, lda, nb, rsrc, csrc, X.dev, X.ld, X.nloc);
        X.store(b, ib, jb, descb, false);
    }
}

}  // namespace slb

using namespace slb;

// PDGETRS through the level-3 path whatever NRHS is (pdgetrs_ itself switches to it above the option solve_l3_min_nrhs); real, TRANS = N / T
extern "C" void slb200_pdgetrs_l3(const char *trans, const int *n, const int *nrhs, const double *a, const int *ia, const int *ja, const int *desca,
                                  const int *ipiv, double *b, const int *ib, const int *jb, const int *descb, int *info)
{
    const int ictxt = desca[CTXT_];
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    *info = 0;
    if (P == -1) { *info = -(700 + CTXT_ + 1); return; }
    if (*n == 0 || *nrhs == 0) return;
    const char t = (char)(trans[0] & ~0x20);
    if ((*ia - 1) % desca[MB_] || (*ja - 1) % desca[NB_] || desca[MB_] != desca[NB_] || (t != 'N' && t != 'T' && t != 'C')) { *info = -1; return; }
    Grid *g = grid_of(ictxt);
    const Window w = window(*n, *n, *ia, *ja, desca, P, Q, myrow, mycol);
    std::vector<int> ipg;
    gather_global_ipiv(g, *n, desca[NB_], w.rsrc, ipiv + w.loff_r, *ia - 1, ipg);
    StageMat<double> A("stage_A", a, desca[LLD_], w.loff_r, w.loff_c, w.mloc, w.nloc);
    getrs_l3_entry(g, t == 'N' ? 'N' : 'T', *n, *nrhs, A.dev, A.ld, desca[NB_], w.rsrc, w.csrc, ipg.data(), b, *ib, *jb, descb);
}

extern "C" {

void pdgemm_(const char *transa, const char *transb, const int *m_, const int *n_, const int *k_, const double *alpha_, const double *a,
             const int *ia, const int *ja, const int *desca, const double *b, const int *ib, const int *jb, const int *descb,
             const double *beta_, double *c, const int *ic, const int *jc, const int *descc)
{
    const int m = *m_, n = *n_, k = *k_, ictxt = descc[CTXT_];
    const double alpha = *alpha_, beta = *beta_;
    const char ta = upc(transa), tb = upc(transb);
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    // argument checks in the order of pdgemm_.c:254-287 (PB_Cchkmat); an error is reported through PXERBLA and the call returns
    int info = 0;
    if (P == -1) info = -(1900 + CTXT_ + 1);
    else if (ta != 'N' && !is_trans(ta)) info = -1;
    else if (tb != 'N' && !is_trans(tb)) info = -2;
    else if (m < 0) info = -3;
    else if (n < 0) info = -4;
    else if (k < 0) info = -5;
    else if (desca[CTXT_] != ictxt) info = -(1000 + CTXT_ + 1);
    else if (descb[CTXT_] != ictxt) info = -(1400 + CTXT_ + 1);
    else if (m && k && (is_trans(ta) ? bad_sub(k, m, *ia, *ja, desca) : bad_sub(m, k, *ia, *ja, desca))) info = -8;
    else if (k && n && (is_trans(tb) ? bad_sub(n, k, *ib, *jb, descb) : bad_sub(k, n, *ib, *jb, descb))) info = -12;
    else if (m && n && bad_sub(m, n, *ic, *jc, descc)) info = -17;
    if (info != 0) { xerbla(ictxt, "PDGEMM", info); return; }
    if (m == 0 || n == 0 || ((alpha == 0.0 || k == 0) && beta == 1.0)) return;         // pdgemm_.c:295-297
    Grid *g = grid_of(ictxt);
    const int nb = working_nb(descc);
    Work C("pb_C", g, m, n, nb);
    C.load(c, *ic, *jc, descc, false);
    C.scale(beta);
    if (alpha != 0.0 && k > 0) {
        Work A("pb_A", g, m, k, nb), B("pb_B", g, k, n, nb);
        A.load(a, *ia, *ja, desca, is_trans(ta));
        B.load(b, *ib, *jb, descb, is_trans(tb));
        A.scale(-alpha);                                                                // C -= (-alpha A) B
        summa_minus(g, nb, A, B, C);
    }
    SLB_CUDA(cudaStreamSynchronize(rt().s_main));
    C.store(c, *ic, *jc, descc, false);
}

void pdtrsm_(const char *side, const char *uplo, const char *transa, const char *diag, const int *m_, const int *n_, const double *alpha_,
             const double *a, const int *ia, const int *ja, const int *desca, double *b, const int *ib, const int *jb, const int *descb)
{
    const int m = *m_, n = *n_, ictxt = descb[CTXT_];
    const double alpha = *alpha_;
    const char sd = upc(side), ul = upc(uplo), ta = upc(transa), dg = upc(diag);
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    const int na = sd == 'L' ? m : n;                                                   // order of the triangular matrix
    int info = 0;                                                                       // pdtrsm_.c:251-289
    if (P == -1) info = -(1500 + CTXT_ + 1);
    else if (sd != 'L' && sd != 'R') info = -1;
    else if (ul != 'U' && ul != 'L') info = -2;
    else if (ta != 'N' && !is_trans(ta)) info = -3;
    else if (dg != 'U' && dg != 'N') info = -4;
    else if (m < 0) info = -5;
    else if (n < 0) info = -6;
    else if (desca[CTXT_] != ictxt) info = -(1100 + CTXT_ + 1);
    else if (na && bad_sub(na, na, *ia, *ja, desca)) info = -9;
    else if (m && n && bad_sub(m, n, *ib, *jb, descb)) info = -13;
    if (info != 0) { xerbla(ictxt, "PDTRSM", info); return; }
    if (m == 0 || n == 0) return;
    Grid *g = grid_of(ictxt);
    const int nb = working_nb(descb);
    const bool right = sd == 'R';
    // X op(A) = alpha B  <=>  op(A)' X' = alpha B': work on X' (n x m) with the triangle op(A)'
    Work X("pb_C", g, right ? n : m, right ? m : n, nb);
    X.load(b, *ib, *jb, descb, right);
    if (alpha == 0.0) { X.scale(0.0); SLB_CUDA(cudaStreamSynchronize(rt().s_main)); X.store(b, *ib, *jb, descb, right); return; }   // pdtrsm_.c:300-305
    X.scale(alpha);
    const bool tr_copy = is_trans(ta) != right;                                         // the working triangle is op(A) (left) or op(A)' (right)
    const bool upper = (ul == 'U') != tr_copy;
    Work T("pb_A", g, na, na, nb);
    T.load(a, *ia, *ja, desca, tr_copy);
    SLB_CUDA(cudaStreamSynchronize(rt().s_main));
    tri_l3_sweep(g, upper, dg == 'U', na, T.dev, T.ld, nb, 0, 0, X.dev, X.ld, X.nloc);
    X.store(b, *ib, *jb, descb, right);
}

void pdtran_(const int *m_, const int *n_, const double *alpha_, const double *a, const int *ia, const int *ja, const int *desca,
             const double *beta_, double *c, const int *ic, const int *jc, const int *descc)
{
    const int m = *m_, n = *n_, ictxt = descc[CTXT_];                                   // sub(C) is m x n, sub(A) is n x m
    const double alpha = *alpha_, beta = *beta_;
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    int info = 0;                                                                       // pdtran_.c:228-245
    if (P == -1) info = -(1200 + CTXT_ + 1);
    else if (m < 0) info = -1;
    else if (n < 0) info = -2;
    else if (desca[CTXT_] != ictxt) info = -(700 + CTXT_ + 1);
    else if (m && n && bad_sub(n, m, *ia, *ja, desca)) info = -5;
    else if (m && n && bad_sub(m, n, *ic, *jc, descc)) info = -10;
    if (info != 0) { xerbla(ictxt, "PDTRAN", info); return; }
    if (m == 0 || n == 0 || (alpha == 0.0 && beta == 1.0)) return;
    Grid *g = grid_of(ictxt);
    const int nb = working_nb(descc);
    Work C("pb_C", g, m, n, nb);
    C.load(c, *ic, *jc, descc, false);
    C.scale(beta);
    if (alpha != 0.0) {
        Work At("pb_A", g, m, n, nb);
        At.load(a, *ia, *ja, desca, true);
        if (C.mloc > 0 && C.nloc > 0) SLB_LAUNCH(axpy_block_kernel, grid1d(C.mloc * C.nloc), 256, rt().s_main, C.mloc, C.nloc, C.dev, C.ld, (const double *)At.dev, At.ld, alpha);
    }
    SLB_CUDA(cudaStreamSynchronize(rt().s_main));
    C.store(c, *ic, *jc, descc, false);
}

}  // extern "C"
