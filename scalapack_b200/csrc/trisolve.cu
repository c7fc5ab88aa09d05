// trisolve.cu -- distributed triangular sweeps on replicated right-hand sides, and the routines made of them:
//   PDPOTRS (SRC/pdpotrs.f:166-265): two PDTRSMs with the Cholesky factor; PDPOSV (SRC/pdposv.f) = PDPOTRF + PDPOTRS.
//
// tri_sweep_device is the loop of PDGETRS (solve.cu) with the triangle, the transposition and the kind of diagonal as
// parameters: the right-hand sides (few, N x NRHS) are replicated on every GPU in global row order, each block step sums the
// partial products of block k over the process row (column for a transposed sweep), solves the nb x nb diagonal block on its
// owner, broadcasts x_k and folds it into the partial products with one HBM-bound GEMV over the panel.  The triangle is read
// exactly once per sweep.
#include "common.h"
#include "dist.h"
#include "kernels.cuh"
#include "launch.h"
#include "lu.h"
#include "ncclw.h"

namespace slb {

namespace {

// out(i, c) = x(i, c) + part(i, c) for the kb rows of block k
__global__ void __launch_bounds__(128)
add_part_kernel(int kb, int nrhs, const double *__restrict__ x, int64_t ldx, const double *__restrict__ part, int64_t ldp,
                double *__restrict__ out, int64_t ldo)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= kb * nrhs) return;
    const int i = e % kb, c = e / kb;
    out[i + (int64_t)c * ldo] = x[i + (int64_t)c * ldx] + part[i + (int64_t)c * ldp];
}

}  // namespace

// Xw (device, N x nrhs, ld = N, global row order, identical on every process) <- op(T)^-1 Xw, T = the upper / lower triangle of
// the distributed N x N matrix at A (first block on (rsrc, csrc)); unit: T has an implicit unit diagonal.
void tri_sweep_device(Grid *g, bool upper, bool trans, bool unit, int N, int nrhs, const double *A, int64_t lld, int nb, int rsrc,
                      int csrc, double *Xw)
{
    Runtime &r = rt();
    cudaStream_t s = r.s_main;
    const int P = g->nprow, Q = g->npcol, myrow = g->myrow, mycol = g->mycol;
    const bool multi = P * Q > 1;
    if (multi && !g->nccl) g->nccl = nccl_create(g);
    NcclComms *nc = g->nccl;
    const int64_t mloc = numroc(N, nb, myrow, rsrc, P), nloc = numroc(N, nb, mycol, csrc, Q);
    const bool fwd = upper == trans;                             // op(T) is lower triangular: top-down
    const int64_t nacc = trans ? nloc : mloc;                    // partial products by local column (transposed) or local row
    const int64_t lda_acc = nacc > 0 ? nacc : 1;
    double *acc = (double *)workspace("ts_acc", (size_t)lda_acc * nrhs * sizeof(double));
    double *red = (double *)workspace("ts_red", (size_t)2 * nb * nrhs * sizeof(double));
    double *tmp = red + (size_t)nb * nrhs;
    int mode = (upper ? TRSV_UPPER : 0) | (trans ? TRSV_TRANS : 0);
    if (!upper && !unit) mode |= TRSV_NONUNIT_L;
    if (upper && unit) fatal("tri_sweep_device: a unit upper triangle is not supported");
    SLB_CUDA(cudaMemsetAsync(acc, 0, (size_t)lda_acc * nrhs * sizeof(double), s));
    const int nblk = (N + nb - 1) / nb;
    for (int q = 0; q < nblk; ++q) {
        const int k = fwd ? q : nblk - 1 - q;
        const int j0 = k * nb, jb = (N - j0) < nb ? (N - j0) : nb;
        const int pr = (rsrc + k) % P, pc = (csrc + k) % Q;
        const int64_t lr0 = numroc(j0, nb, myrow, rsrc, P), lc0 = numroc(j0, nb, mycol, csrc, Q);
        double *xk = Xw + j0;
        const bool holds = trans ? mycol == pc : myrow == pr;    // the process row / column that holds block k's partial products
        if (holds) {
            const double *part = acc + (trans ? lc0 : lr0);
            int64_t ldp = lda_acc;
            if ((trans ? P : Q) > 1) {
                launch_copy2d<double>(jb, nrhs, part, lda_acc, red, jb, s);
                nccl_allreduce_sum_f64(trans ? nc->col : nc->row, red, red, (size_t)jb * nrhs, s);
                part = red; ldp = jb;
            }
            if (myrow == pr && mycol == pc) {
                SLB_LAUNCH(add_part_kernel, (unsigned)((jb * nrhs + 127) / 128), 128, s, jb, nrhs, (const double *)xk, (int64_t)N, part, ldp, tmp, (int64_t)jb);
                launch_trsv_block<double>(jb, A + lr0 + lc0 * lld, lld, tmp, jb, nrhs, mode, s);
            }
        }
        if (multi) nccl_bcast(nc->all, tmp, (size_t)jb * nrhs * sizeof(double), NT_U8, pr * Q + pc, s);
        launch_copy2d<double>(jb, nrhs, tmp, jb, xk, N, s);
        if (!trans) {
            if (mycol == pc) {
                if (fwd) {                                       // rows below block k
                    const int64_t rbeg = numroc(j0 + jb, nb, myrow, rsrc, P);
                    launch_gemv_minus<double>(mloc - rbeg, jb, A + rbeg + lc0 * lld, lld, tmp, jb, acc + rbeg, lda_acc, nrhs, s);
                } else launch_gemv_minus<double>(lr0, jb, A + lc0 * lld, lld, tmp, jb, acc, lda_acc, nrhs, s);   // rows above
            }
        } else if (myrow == pr) {
            if (fwd) {                                           // columns right of block k
                const int64_t cbeg = numroc(j0 + jb, nb, mycol, csrc, Q);
                launch_gemvt_minus<double>(jb, nloc - cbeg, A + lr0 + cbeg * lld, lld, tmp, jb, acc + cbeg, lda_acc, nrhs, false, s);
            } else launch_gemvt_minus<double>(jb, lc0, A + lr0, lld, tmp, jb, acc, lda_acc, nrhs, false, s);     // columns left
        }
    }
    SLB_CUDA(cudaStreamSynchronize(s));
}

// x <- A^-1 x for A = L L' ('L') or U' U ('U') given its Cholesky factor (pdpotrs.f:249-263)
void potrs_device(Grid *g, bool upper, int N, int nrhs, const double *A, int64_t lld, int nb, int rsrc, int csrc, double *Xw)
{
    if (upper) {
        tri_sweep_device(g, true, true, false, N, nrhs, A, lld, nb, rsrc, csrc, Xw);
        tri_sweep_device(g, true, false, false, N, nrhs, A, lld, nb, rsrc, csrc, Xw);
    } else {
        tri_sweep_device(g, false, false, false, N, nrhs, A, lld, nb, rsrc, csrc, Xw);
        tri_sweep_device(g, false, true, false, N, nrhs, A, lld, nb, rsrc, csrc, Xw);
    }
}

void potrf_device(Grid *g, bool upper, int N, double *A, int64_t lld, int nb, int rsrc, int csrc, int *info_host);   // chol.cu
void potrs_l3_entry(Grid *g, bool upper, int n, int nrhs, const double *Adev, int64_t lda, int nb, int rsrc, int csrc, double *b, int ib, int jb,
                    const int *descb);                                                                               // pblas.cu

namespace {

void potrs_checks(const char *name, int DA, int DB, int mbcode, const char *uplo, int n, int nrhs, int ia, int ja, const int *desca, int ib, int jb,
                  const int *descb, int *info)
{
    const int ictxt = desca[CTXT_];
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    const char u = (char)(uplo[0] & ~0x20);
    *info = 0;
    if (P == -1) *info = -(DA * 100 + CTXT_ + 1);
    else {
        chk1mat(n, 2, n, 2, ia, ja, desca, DA, info);
        chk1mat(n, 2, nrhs, 3, ib, jb, descb, DB, info);
        if (*info == 0) {
            const int iarow = indxg2p(ia, desca[MB_], desca[RSRC_], P), ibrow = indxg2p(ib, descb[MB_], descb[RSRC_], P);
            const int iroffa = (ia - 1) % desca[MB_], iroffb = (ib - 1) % descb[MB_], icoffa = (ja - 1) % desca[NB_];
            if (u != 'U' && u != 'L') *info = -1;
            else if (iroffa != 0) *info = -5;
            else if (icoffa != 0) *info = -6;
            else if (desca[MB_] != desca[NB_]) *info = -(DA * 100 + NB_ + 1);
            else if (iroffb != 0 || ibrow != iarow) *info = -(DB - 2);         // IB: position 9 of PDPOTRS, 9 of PDPOSV
            else if (descb[MB_] != desca[NB_]) *info = -(mbcode + NB_ + 1);
        }
        int ex[1] = { u == 'U' ? 'U' : 'L' }, expos[1] = { 1 }, one = 1, two = 2, three = 3;
        pchk2mat_(&n, &two, &n, &two, &ia, &ja, desca, &DA, &n, &two, &nrhs, &three, &ib, &jb, descb, &DB, &one, ex, expos, info);
    }
    (void)name;
}

// sub(B) -> replicated device block, sweeps, -> sub(B)
void potrs_run(Grid *g, bool upper, int n, int nrhs, const double *Adev, int64_t lda, int nb, const Window &w, double *b, int ib, int jb,
               const int *descb)
{
    cudaStream_t s = rt().s_main;
    // many right-hand sides: block-cyclic copy of sub(B), level-3 sweeps on the tensor cores (option potrs_l3_min_nrhs, 0 = never)
    const int64_t l3min = opt("potrs_l3_min_nrhs", 64);
    if (l3min > 0 && nrhs > l3min) { potrs_l3_entry(g, upper, n, nrhs, Adev, lda, nb, w.rsrc, w.csrc, b, ib, jb, descb); return; }
    std::vector<double> xg;
    gather_small(g, n, nrhs, b, ib, jb, descb, xg);
    double *Xw = (double *)workspace("ts_X", xg.size() * sizeof(double));
    SLB_CUDA(cudaMemcpyAsync(Xw, xg.data(), xg.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    potrs_device(g, upper, n, nrhs, Adev, lda, nb, w.rsrc, w.csrc, Xw);
    SLB_CUDA(cudaMemcpyAsync(xg.data(), Xw, xg.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    SLB_CUDA(cudaStreamSynchronize(s));
    scatter_small(g, n, nrhs, b, ib, jb, descb, xg);
}

}  // namespace

}  // namespace slb

using namespace slb;

extern "C" void pdpotrf_(const char *uplo, const int *n, double *a, const int *ia, const int *ja, const int *desca, int *info);   // chol.cu

extern "C" {

void pdpotrs_(const char *uplo, const int *n, const int *nrhs, const double *a, const int *ia, const int *ja, const int *desca, double *b,
              const int *ib, const int *jb, const int *descb, int *info)
{
    const int ictxt = desca[CTXT_];
    potrs_checks("PDPOTRS", 7, 11, 1100, uplo, *n, *nrhs, *ia, *ja, desca, *ib, *jb, descb, info);
    if (*info != 0) { xerbla(ictxt, "PDPOTRS", *info); return; }
    if (*n == 0 || *nrhs == 0) return;
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    Grid *g = grid_of(ictxt);
    const Window w = window(*n, *n, *ia, *ja, desca, P, Q, myrow, mycol);
    StageMat<double> A("stage_A", a, desca[LLD_], w.loff_r, w.loff_c, w.mloc, w.nloc);
    potrs_run(g, (uplo[0] & ~0x20) == 'U', *n, *nrhs, A.dev, A.ld, desca[NB_], w, b, *ib, *jb, descb);
}

void pdposv_(const char *uplo, const int *n, const int *nrhs, double *a, const int *ia, const int *ja, const int *desca, double *b,
             const int *ib, const int *jb, const int *descb, int *info)
{
    const int ictxt = desca[CTXT_];
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    // pdposv.f:197-245 checks A and the ALIGNMENT of B only: B's descriptor, NRHS and JB are first looked at by the inner PDPOTRS,
    // after the factorisation -- with PDPOTRS's argument numbers (executing the source shows it; tests/golden/errors_reference.json)
    const char u = (char)(uplo[0] & ~0x20);
    *info = 0;
    if (P == -1) *info = -(700 + CTXT_ + 1);
    else {
        chk1mat(*n, 2, *n, 2, *ia, *ja, desca, 7, info);
        if (*info == 0) {
            const int mbb = descb[MB_] > 0 ? descb[MB_] : 1;                            // a zero MB_B is a division by zero in the reference
            const int iarow = indxg2p(*ia, desca[MB_], desca[RSRC_], P), ibrow = indxg2p(*ib, mbb, descb[RSRC_], P);
            if (u != 'U' && u != 'L') *info = -1;
            else if ((*ia - 1) % desca[MB_]) *info = -5;
            else if ((*ja - 1) % desca[NB_]) *info = -6;
            else if (desca[MB_] != desca[NB_]) *info = -(700 + NB_ + 1);
            else if ((*ib - 1) % mbb || ibrow != iarow) *info = -9;
            else if (descb[MB_] != desca[NB_]) *info = -(1000 + NB_ + 1);               // sic: DESCB is argument 11
        }
        int ex[1] = { u == 'U' ? 'U' : 'L' }, expos[1] = { 1 }, one = 1, two = 2, three = 3, p7 = 7, p11 = 11;
        pchk2mat_(n, &two, n, &two, ia, ja, desca, &p7, n, &two, nrhs, &three, ib, jb, descb, &p11, &one, ex, expos, info);
    }
    if (*info != 0) { xerbla(ictxt, "PDPOSV", *info); return; }
    pdpotrf_(uplo, n, a, ia, ja, desca, info);                                          // pdposv.f:255
    if (*info == 0) pdpotrs_(uplo, n, nrhs, a, ia, ja, desca, b, ib, jb, descb, info);  // pdposv.f:262-268
}

}  // extern "C"
