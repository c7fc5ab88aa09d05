// inverse.cu -- SURVEY 8(f) row 4: PDGETRI (SRC/pdgetri.f:186-379), the inverse from the factors of PDGETRF.
//
// The reference inverts U (PDTRTRI) and then solves inv(A) L = inv(U) block column by block column with PDGEMM / PDTRSM and a
// work array, finishing with column interchanges.  Here the inverse is the solution of  L U X = P  computed by the LEVEL-3
// distributed solve getrs_l3_device: X is distributed like A, the right-hand side is the permuted identity generated in place,
// and both sweeps are made of exactly the LU's own hot kernels -- the unit-lower DMMA TRSM on a block row of X and the trailing
// update kernel (C -= A B) on the rows below / above -- so every flop of the 2 N^3 runs on the FP64 tensor cores.  The upper
// (non-unit) diagonal blocks go through the same unit-lower TRSM after reversing their index order (J U J is lower triangular)
// and splitting off the diagonal.  The same routine is the large-NRHS form of the triangular solves (PB_CptrsmAB's role).
//
// Communication per block step on a P x Q grid (NCCL): the diagonal block along its process row, the solved block row of X
// down every process column, the panel of L (or U) along every process row.
#include "common.h"
#include "dist.h"
#include "kernels.cuh"
#include "launch.h"
#include "lu.h"
#include "ncclw.h"

namespace slb {

namespace {

// The jb x jb triangular block Tk as a UNIT lower triangle times a diagonal: F = Tk (flip == 0, lower) or F = J Tk J (flip == 1,
// upper; J reverses the index order: F(i, j) = Tk(jb-1-i, jb-1-j) is lower triangular).  d_j = 1 if unitdiag else F(j, j);
// Lt(i, j) = F(i, j) / d_j for i > j (0 elsewhere); dinv[i] = 1 / d_i.  Then F^-1 = diag(dinv) Lt^-1.
__global__ void __launch_bounds__(256)
unit_tri_kernel(int jb, const double *__restrict__ Tk, int64_t ldt, int flip, int unitdiag, double *__restrict__ Lt, double *__restrict__ dinv)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= jb * jb) return;
    const int i = e % jb, j = e / jb;
    const int si = flip ? jb - 1 - i : i, sj = flip ? jb - 1 - j : j;
    const double djj = unitdiag ? 1.0 : Tk[sj + (int64_t)sj * ldt];
    Lt[i + (int64_t)j * jb] = i > j ? Tk[si + (int64_t)sj * ldt] / djj : 0.0;
    if (i == j) dinv[i] = 1.0 / djj;
}

// M(k, j) *= scale[k] for the jb x n block at M (ld)
__global__ void __launch_bounds__(256)
scale_block_rows_kernel(int jb, int64_t n, double *__restrict__ M, int64_t ld, const double *__restrict__ scale)
{
    const int64_t total = (int64_t)jb * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = e % jb, j = e / jb;
        M[k + j * ld] *= scale[k];
    }
}

// dst(jb-1-i, j) = src(i, j) * (scale ? scale[i] : 1): reverse the row order of a jb x n block (optionally scaling source row i)
__global__ void __launch_bounds__(256)
flip_rows_kernel(int jb, int64_t n, const double *__restrict__ src, int64_t lds, double *__restrict__ dst, int64_t ldd,
                 const double *__restrict__ scale)
{
    const int64_t total = (int64_t)jb * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % jb, j = e / jb;
        const double v = src[i + j * lds];
        dst[(jb - 1 - i) + j * ldd] = scale ? v * scale[i] : v;
    }
}

// X(il, jl) = 1 where the global column of jl equals perm[global row of il], else 0: the rows of the identity interchanged like P b
__global__ void __launch_bounds__(256)
perm_identity_kernel(int64_t mloc, int64_t nloc, int nb, int P, int Q, int myr_rel, int myc_rel, const int *__restrict__ perm,
                     double *__restrict__ X, int64_t ldx)
{
    const int64_t total = mloc * nloc;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t il = e % mloc, jl = e / mloc;
        const int64_t ig = ((il / nb) * P + myr_rel) * nb + il % nb, jg = ((jl / nb) * Q + myc_rel) * nb + jl % nb;
        X[il + jl * ldx] = perm[ig] == (int)jg ? 1.0 : 0.0;
    }
}

// first exactly-zero diagonal entry of the jb x jb block at Akk (global index j0 + i + 1) into *info if it is still 0
__global__ void zero_diag_kernel(int jb, const double *__restrict__ Akk, int64_t ld, int j0, int *info)
{
    if (blockIdx.x != 0 || threadIdx.x != 0 || *info != 0) return;
    for (int i = 0; i < jb; ++i) if (Akk[i + (int64_t)i * ld] == 0.0) { *info = j0 + i + 1; return; }
}

}  // namespace

// X <- T^-1 X, level 3, T = the lower / upper triangle (unit: implicit unit diagonal) of the distributed N x N matrix at A (local
// window, first block on (rsrc, csrc)).  X: local array whose ROWS are distributed like the rows of A (same nb, rsrc) and which
// holds nlocx local columns on this process (any column distribution).  The other triangle of A is never read.
void tri_l3_sweep(Grid *g, bool upper, bool unit, int N, const double *A, int64_t lld, int nb, int rsrc, int csrc, double *X, int64_t ldx,
                  int64_t nlocx)
{
    Runtime &r = rt();
    cudaStream_t s = r.s_main;
    const int P = g->nprow, Q = g->npcol, myrow = g->myrow, mycol = g->mycol;
    if (P * Q > 1 && !g->nccl) g->nccl = nccl_create(g);
    NcclComms *nc = g->nccl;
    const int64_t mloc = numroc(N, nb, myrow, rsrc, P);
    // a column communicator moves jb x nlocx blocks of X (nlocx may differ between process columns, not inside one);
    // a row communicator only ever moves pieces of A
    double *Dk = (double *)workspace("l3_D", ((size_t)nb * nb * 2 + nb) * sizeof(double));
    double *Lt = Dk + (size_t)nb * nb, *dinv = Lt + (size_t)nb * nb;
    double *Pan = (double *)workspace("l3_pan", (size_t)nb * (mloc > 0 ? mloc : 1) * sizeof(double));          // my rows of the panel, contiguous
    double *Xk = (double *)workspace("l3_xk", (size_t)nb * (nlocx > 0 ? nlocx : 1) * 2 * sizeof(double));       // block row k of X (+ a flipped copy)
    double *Yk = Xk + (size_t)nb * (nlocx > 0 ? nlocx : 1);
    const int nblk = (N + nb - 1) / nb;
    const bool fwd = !upper;                                     // lower: top-down, upper: bottom-up
    for (int q = 0; q < nblk; ++q) {
        const int k = fwd ? q : nblk - 1 - q;
        const int j0 = k * nb, jb = N - j0 < nb ? N - j0 : nb;
        const int pr = (rsrc + k) % P, pc = (csrc + k) % Q;
        const int64_t lr0 = numroc(j0, nb, myrow, rsrc, P), lc0 = numroc(j0, nb, mycol, csrc, Q);
        // ---- block row k of X on process row pr: X_k <- T_kk^-1 X_k ----
        if (myrow == pr) {
            const double *Tkk = A + lr0 + lc0 * lld; int64_t ldt = lld;
            if (Q > 1) {
                if (mycol == pc) launch_copy2d<double>(jb, jb, Tkk, lld, Dk, jb, s);
                nccl_bcast(nc->row, Dk, (size_t)jb * jb, NT_F64, pc, s);
                Tkk = Dk; ldt = jb;
            }
            if (nlocx > 0) {
                double *Xrow = X + lr0;
                if (!upper && unit) launch_dtrsm_llnu(jb, nlocx, Tkk, ldt, Xrow, ldx, s);
                else {
                    SLB_LAUNCH(unit_tri_kernel, (unsigned)((jb * jb + 255) / 256), 256, s, jb, Tkk, ldt, upper ? 1 : 0, unit ? 1 : 0, Lt, dinv);
                    if (!upper) {
                        launch_dtrsm_llnu(jb, nlocx, Lt, jb, Xrow, ldx, s);
                        SLB_LAUNCH(scale_block_rows_kernel, grid1d((int64_t)jb * nlocx), 256, s, jb, nlocx, Xrow, ldx, (const double *)dinv);
                    } else {
                        SLB_LAUNCH(flip_rows_kernel, grid1d((int64_t)jb * nlocx), 256, s, jb, nlocx, (const double *)Xrow, ldx, Yk, (int64_t)jb, (const double *)nullptr);
                        launch_dtrsm_llnu(jb, nlocx, Lt, jb, Yk, jb, s);
                        SLB_LAUNCH(flip_rows_kernel, grid1d((int64_t)jb * nlocx), 256, s, jb, nlocx, (const double *)Yk, (int64_t)jb, Xrow, ldx, (const double *)dinv);
                    }
                }
            }
        }
        // ---- the rows still to be solved: below block k (lower) / above it (upper) ----
        const int64_t rbeg = fwd ? numroc(j0 + jb, nb, myrow, rsrc, P) : 0, rend = fwd ? mloc : lr0;
        const int64_t mr = rend - rbeg;
        // solved block row to every process row (each process column moves its own jb x nlocx block)
        const double *Bop = X + lr0; int64_t ldb = ldx;
        if (P > 1) {
            if (nlocx > 0) {
                if (myrow == pr) launch_copy2d<double>(jb, nlocx, X + lr0, ldx, Xk, jb, s);
                nccl_bcast(nc->col, Xk, (size_t)jb * nlocx, NT_F64, pr, s);
            }
            Bop = Xk; ldb = jb;
        }
        // my rows of the panel of block column k along my process row (mr depends on the process row only)
        const double *Aop = A + rbeg + lc0 * lld; int64_t lda = lld;
        if (Q > 1) {
            if (mr > 0) {
                if (mycol == pc) launch_copy2d<double>(mr, jb, A + rbeg + lc0 * lld, lld, Pan, mr, s);
                nccl_bcast(nc->row, Pan, (size_t)mr * jb, NT_F64, pc, s);
            }
            Aop = Pan; lda = mr;
        }
        if (mr > 0 && nlocx > 0) launch_dgemm_minus(mr, nlocx, jb, Aop, lda, Bop, ldb, X + rbeg, ldx, s);
    }
    SLB_CUDA(cudaStreamSynchronize(s));
}

// X <- U^-1 L^-1 X with the factors of PDGETRF at A
void getrs_l3_device(Grid *g, int N, const double *A, int64_t lld, int nb, int rsrc, int csrc, double *X, int64_t ldx, int64_t nlocx)
{
    tri_l3_sweep(g, false, true, N, A, lld, nb, rsrc, csrc, X, ldx, nlocx);
    tri_l3_sweep(g, true, false, N, A, lld, nb, rsrc, csrc, X, ldx, nlocx);
}

}  // namespace slb

using namespace slb;

extern "C" void pdgetri_(const int *n_, double *a, const int *ia_, const int *ja_, const int *desca, const int *ipiv, double *work,
                         const int *lwork_, int *iwork, const int *liwork_, int *info)
{
    const int n = *n_, ia = *ia_, ja = *ja_, lwork = *lwork_, liwork = *liwork_;
    const int ictxt = desca[CTXT_];
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    bool lquery = false;
    *info = 0;
    if (P == -1) *info = -(500 + CTXT_ + 1);
    else {
        chk1mat(n, 1, n, 1, ia, ja, desca, 5, info);
        if (*info == 0) {
            const int mb = desca[MB_], nb = desca[NB_];
            const int iroff = (ia - 1) % mb, icoff = (ja - 1) % nb, iarow = indxg2p(ia, mb, desca[RSRC_], P);
            const int np = numroc(n + iroff, mb, myrow, iarow, P);
            const int lwmin = np * nb;
            const int nq = numroc(desca[N_], nb, mycol, desca[CSRC_], Q);
            int liwmin;
            auto iceil = [](int a_, int b_) { return (a_ + b_ - 1) / b_; };
            if (P == Q) liwmin = nq + nb;
            else {
                int a_ = P, b_ = Q; while (b_) { int t = a_ % b_; a_ = b_; b_ = t; }
                const int lcm = P / a_ * Q;
                const int t1 = numroc(desca[M_] + mb * P + (ia - 1) % mb, nb, mycol, desca[CSRC_], Q);
                const int t2 = mb * iceil(iceil(numroc(desca[M_] + mb * P, mb, myrow, desca[RSRC_], P), mb), lcm / P);
                liwmin = t1 + (t2 > nb ? t2 : nb);
            }
            work[0] = (double)lwmin; iwork[0] = liwmin;
            lquery = lwork == -1 || liwork == -1;
            if (iroff != icoff || iroff != 0) *info = -4;
            else if (mb != nb) *info = -(500 + NB_ + 1);
            else if (lwork < lwmin && !lquery) *info = -8;
            else if (liwork < liwmin && !lquery) *info = -10;
        }
        int ex[2] = { lwork == -1 ? -1 : 1, liwork == -1 ? -1 : 1 }, expos[2] = { 8, 10 }, one = 1, two = 2, five = 5;
        pchk1mat_(&n, &one, &n, &one, &ia, &ja, desca, &five, &two, ex, expos, info);
    }
    if (*info != 0) { xerbla(ictxt, "PDGETRI", *info); return; }
    if (lquery) return;
    if (n == 0) return;
    Grid *g = grid_of(ictxt);
    cudaStream_t s = rt().s_main;
    const int nb = desca[NB_];
    const Window w = window(n, n, ia, ja, desca, P, Q, myrow, mycol);
    StageMat<double> A("stage_A", a, desca[LLD_], w.loff_r, w.loff_c, w.mloc, w.nloc);
    // ---- singular U: INFO = i of the first zero U(i, i), the inverse is not computed (PDTRTRI, pdgetri.f:306-309) ----
    int *info_dev = (int *)workspace("l3_info", 64);
    SLB_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), s));
    const int nblk = (n + nb - 1) / nb;
    for (int k = 0; k < nblk; ++k) {
        if ((w.rsrc + k) % P != myrow || (w.csrc + k) % Q != mycol) continue;
        const int j0 = k * nb, jb = n - j0 < nb ? n - j0 : nb;
        SLB_LAUNCH(zero_diag_kernel, 1, 1, s, jb, (const double *)(A.dev + numroc(j0, nb, myrow, w.rsrc, P) + (int64_t)numroc(j0, nb, mycol, w.csrc, Q) * A.ld), A.ld, j0, info_dev);
    }
    int myinfo = 0;
    SLB_CUDA(cudaMemcpyAsync(&myinfo, info_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
    SLB_CUDA(cudaStreamSynchronize(s));
    {   // the smallest positive report over the grid
        const int np = P * Q;
        std::vector<int> all((size_t)np, myinfo);
        if (np > 1) grid_allgather(g, 'A', &myinfo, all.data(), sizeof(int));
        int best = 0; for (int v : all) if (v > 0 && (best == 0 || v < best)) best = v;
        *info = best;
    }
    if (*info > 0) return;
    // ---- X = P (rows of the identity interchanged), then X <- U^-1 L^-1 X ----
    std::vector<int> ipg, perm((size_t)n);
    gather_global_ipiv(g, n, nb, w.rsrc, ipiv + w.loff_r, ia - 1, ipg);
    for (int i = 0; i < n; ++i) perm[(size_t)i] = i;
    for (int i = 0; i < n; ++i) { const int p = ipg[(size_t)i] - 1; if (p != i) { const int t = perm[(size_t)i]; perm[(size_t)i] = perm[(size_t)p]; perm[(size_t)p] = t; } }
    int *perm_dev = (int *)workspace("l3_perm", (size_t)n * sizeof(int));
    SLB_CUDA(cudaMemcpyAsync(perm_dev, perm.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
    const int64_t ldx = (w.mloc + 1) & ~(int64_t)1;
    double *X = (double *)workspace("l3_X", (size_t)(ldx > 0 ? ldx : 2) * (size_t)(w.nloc > 0 ? w.nloc : 1) * sizeof(double));
    if (w.mloc > 0 && w.nloc > 0)
        SLB_LAUNCH(perm_identity_kernel, grid1d(w.mloc * w.nloc), 256, s, w.mloc, w.nloc, nb, P, Q, (P + myrow - w.rsrc) % P, (Q + mycol - w.csrc) % Q,
                   (const int *)perm_dev, X, ldx);
    getrs_l3_device(g, n, A.dev, A.ld, nb, w.rsrc, w.csrc, X, ldx, w.nloc);
    if (w.mloc > 0 && w.nloc > 0) launch_copy2d<double>(w.mloc, w.nloc, X, ldx, A.dev, A.ld, s);
    SLB_CUDA(cudaStreamSynchronize(s));
    A.download();
}
