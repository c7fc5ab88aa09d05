// chol.cu -- SURVEY 8(f) row 3: PDPOTRF (SRC/pdpotrf.f:166-362) = PDPOTF2 on the diagonal block + PDTRSM on the panel +
// PDSYRK on the trailing matrix, for UPLO = 'L' and 'U', on P x Q grids.  PDPOTRS / PDPOSV are in trisolve.cu.
//
// The same skeleton as the LU minus pivoting, built on the LU's hot kernels:
//   * the trailing update runs on the FP64 tensor cores through the LU's update kernel (launch_dgemm_minus: C -= A B).
//     Both triangles are written as  C(I,J) -= sum_k T(k,I) T(k,J)  with ONE transposed panel T (jb x trailing, k contiguous):
//     T = L21^T for 'L', T = U12 for 'U'.  T restricted to my local columns IS the B operand of the update kernel; T restricted
//     to my local rows, transposed once, is its A operand.  The update walks my local block columns, each launch covering the
//     rows on the triangle's side of that block column; the one diagonal block a launch crosses is saved before and its other
//     triangle put back after (the reference never touches the other triangle, SRC/pdpotrf.f:57-64).
//   * the panel solve A21 L11^-T ('L') / U11^-T A12 ('U') runs through the LU's unit-lower DMMA TRSM (launch_dtrsm_llnu) on
//     the transposed panel: L11 = Lt D with Lt unit lower, so L11^-1 = D^-1 Lt^-1.
//   * the nb x nb diagonal block is factored in 32-column steps: a one-warp kernel on the 32 x 32 diagonal piece, a
//     thread-per-row solve below it and a thread-per-element rank-32 update of the rest of the block.
//   * communication per block step (NCCL): INFO to everyone, L11 along the panel's process column (row for 'U'), each process
//     row's piece of T along its row, and one all-gather of the pieces along the columns (rows for 'U').
#include "common.h"
#include "dist.h"
#include "kernels.cuh"
#include "launch.h"
#include "lu.h"
#include "ncclw.h"

namespace slb {

namespace {

// element (i, j), i >= j, of the lower-triangular factor stored in D: for UPPER the factor is U^T, i.e. D(j, i)
template <bool UPPER>
__device__ __forceinline__ double &tri(double *D, int64_t ld, int i, int j) { return UPPER ? D[j + (int64_t)i * ld] : D[i + (int64_t)j * ld]; }

// Cholesky of the w x w (w <= 32) diagonal piece at D, one thread per row, right-looking (the arithmetic of the unblocked
// PDPOTF2, SRC/pdpotf2.f:206-262, column by column).  *info != 0 on entry: do nothing.  A pivot <= 0 (or NaN): *info = col0 + j + 1.
template <bool UPPER>
__global__ void __launch_bounds__(32)
potf2_32_kernel(int w, double *D, int64_t ld, int *info, int col0)
{
    __shared__ double a[32 * 33];
    __shared__ int bad, entry;                                   // two flags: `bad` is written again before the next barrier
    const int i = threadIdx.x;
    if (i == 0) { entry = *info; bad = 0; }
    for (int j = 0; j < 32; ++j) a[i * 33 + j] = (i < w && j <= i && j < w) ? tri<UPPER>(D, ld, i, j) : 0.0;
    __syncthreads();
    if (entry != 0) return;
    for (int j = 0; j < w; ++j) {
        if (i == j) {
            const double ajj = a[j * 33 + j];
            if (!(ajj > 0.0)) bad = col0 + j + 1;            // pdpotf2.f:225-229 (AJJ <= ZERO; a NaN fails the test too)
            else a[j * 33 + j] = sqrt(ajj);
        }
        __syncthreads();
        if (bad != 0) break;
        const double rd = 1.0 / a[j * 33 + j];                   // DSCAL by ONE / AJJ (pdpotf2.f:240)
        if (i > j && i < w) a[i * 33 + j] = a[i * 33 + j] * rd;
        __syncthreads();
        if (i > j && i < w) {
            const double lij = a[i * 33 + j];
            for (int c = j + 1; c <= i; ++c) a[i * 33 + c] -= lij * a[c * 33 + j];
        }
        __syncthreads();
    }
    // columns finished before a failure are final; the failing diagonal keeps its non-positive value (pdpotf2.f:226)
    for (int j = 0; j < 32; ++j) if (i < w && j <= i) tri<UPPER>(D, ld, i, j) = a[i * 33 + j];
    if (i == 0 && bad != 0) *info = bad;
}

// rows below the 32-column diagonal piece, inside the diagonal block: X(r, :) <- X(r, :) Lss^-T, one thread per row
// (X(r, c) = element (s + w + r, s + c) of the factor; Lss = the w x w piece just factored)
template <bool UPPER>
__global__ void __launch_bounds__(128)
potf2_below_kernel(int rows, int w, double *D, int64_t ld, int s, const int *info)
{
    if (*info != 0) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int i = s + w + r;
    for (int c = 0; c < w; ++c) {
        double x = tri<UPPER>(D, ld, i, s + c);
        for (int k = 0; k < c; ++k) x -= tri<UPPER>(D, ld, i, s + k) * tri<UPPER>(D, ld, s + c, s + k);
        tri<UPPER>(D, ld, i, s + c) = x / tri<UPPER>(D, ld, s + c, s + c);
    }
}

// rest of the diagonal block: F(i, j) -= sum_k X(i, k) X(j, k) for i >= j >= s + w, one thread per element
template <bool UPPER>
__global__ void __launch_bounds__(256)
potf2_rank_kernel(int rem, int w, double *D, int64_t ld, int s, const int *info)
{
    if (*info != 0) return;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)rem * rem) return;
    const int ri = (int)(e % rem), rj = (int)(e / rem);
    if (ri < rj) return;
    const int i = s + w + ri, j = s + w + rj;
    double acc = tri<UPPER>(D, ld, i, j);
    for (int k = 0; k < w; ++k) acc -= tri<UPPER>(D, ld, i, s + k) * tri<UPPER>(D, ld, j, s + k);
    tri<UPPER>(D, ld, i, j) = acc;
}

// Lt = unit-lower scaling of the factor F in D: Lt(i, j) = F(i, j) / F(j, j) for i > j (0 elsewhere); dinv[j] = 1 / F(j, j)
template <bool UPPER>
__global__ void __launch_bounds__(256)
unit_lower_kernel(int jb, double *D, int64_t ld, double *Lt, double *dinv)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= jb * jb) return;
    const int i = e % jb, j = e / jb;
    const double djj = tri<UPPER>(D, ld, j, j);
    Lt[i + (int64_t)j * jb] = i > j ? tri<UPPER>(D, ld, i, j) / djj : 0.0;
    if (i == j) dinv[j] = 1.0 / djj;
}

// dst(c, r) = src(r, c) [* scale[c]]: rows x cols of src (ld lds) -> cols x rows of dst (ld ldd)
__global__ void __launch_bounds__(256)
transpose_kernel(int64_t rows, int64_t cols, const double *__restrict__ src, int64_t lds, double *__restrict__ dst, int64_t ldd,
                 const double *__restrict__ scale)
{
    const int64_t total = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e % rows, c = e / rows;
        const double v = src[r + c * lds];
        dst[c + r * ldd] = scale ? v * scale[c] : v;
    }
}

// M(k, j) *= scale[k] for the jb x n block at M (ld)
__global__ void __launch_bounds__(256)
scale_rows_kernel(int jb, int64_t n, double *__restrict__ M, int64_t ld, const double *__restrict__ scale)
{
    const int64_t total = (int64_t)jb * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = e % jb, j = e / jb;
        M[k + j * ld] *= scale[k];
    }
}

// dst(:, j) = src(:, idx[j]) for jb-row column-major blocks with ld = jb (idx: source COLUMN numbers)
__global__ void __launch_bounds__(256)
gather_cols_kernel(int jb, int64_t n, const int64_t *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst)
{
    const int64_t total = (int64_t)jb * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = e % jb, j = e / jb;
        dst[e] = src[idx[j] * jb + k];
    }
}

// put back the triangle of a w x w diagonal block that the factorisation must not touch: strictly upper for 'L', strictly lower for 'U'
template <bool UPPER>
__global__ void __launch_bounds__(256)
restore_tri_kernel(int w, const double *__restrict__ saved, int64_t lds, double *__restrict__ A, int64_t lda)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= w * w) return;
    const int i = e % w, j = e / w;
    if (UPPER ? i > j : i < j) A[i + (int64_t)j * lda] = saved[i + (int64_t)j * lds];
}

template <bool UPPER>
void diag_potf2(int jb, double *D, int64_t ld, int *info_dev, int col0, cudaStream_t s)
{
    for (int s0 = 0; s0 < jb; s0 += 32) {
        const int w = jb - s0 < 32 ? jb - s0 : 32, rem = jb - s0 - w;
        SLB_LAUNCH_SYNC((potf2_32_kernel<UPPER>), 1, 32, s, w, D + s0 + (int64_t)s0 * ld, ld, info_dev, col0 + s0);
        if (rem > 0) {
            SLB_LAUNCH((potf2_below_kernel<UPPER>), (unsigned)((rem + 127) / 128), 128, s, rem, w, D, ld, s0, info_dev);
            SLB_LAUNCH((potf2_rank_kernel<UPPER>), (unsigned)(((int64_t)rem * rem + 255) / 256), 256, s, rem, w, D, ld, s0, info_dev);
        }
    }
}

}  // namespace

// A: device pointer to the local window of sub(A) (lld x LOCc(N)), first block on process (rsrc, csrc).  *info_host: 0, or the
// order k of the leading minor that is not positive definite (the factorisation stops at that block step like pdpotrf.f:280-284).
template <bool UPPER>
static void potrf_sweep(Grid *g, int N, double *A, int64_t lld, int nb, int rsrc, int csrc, int *info_host)
{
    Runtime &r = rt();
    cudaStream_t s = r.s_main;
    const int P = g->nprow, Q = g->npcol, myrow = g->myrow, mycol = g->mycol;
    const bool multi = P * Q > 1;
    if (multi && !g->nccl) g->nccl = nccl_create(g);
    NcclComms *nc = g->nccl;
    const int64_t mloc = numroc(N, nb, myrow, rsrc, P), nloc = numroc(N, nb, mycol, csrc, Q);
    // 'L': T pieces live by process ROW (piece of row r = my local trailing rows); 'U': by process COLUMN.
    // "along": the dimension the pieces are indexed by (rows for 'L'), "across": the other one.
    const int np_along = UPPER ? Q : P;
    int64_t piece_max = 0;                                      // most trailing indices any process holds along that dimension
    for (int p = 0; p < np_along; ++p) {
        const int64_t v = UPPER ? numroc(N, nb, p, csrc, Q) : numroc(N, nb, p, rsrc, P);
        if (v > piece_max) piece_max = v;
    }
    if (piece_max < 1) piece_max = 1;
    const int64_t macross = UPPER ? mloc : nloc;                // my local extent in the other dimension
    double *Dbuf = (double *)workspace("ch_D", ((size_t)nb * nb * 3 + nb) * sizeof(double));
    double *Lt = Dbuf + (size_t)nb * nb, *Sv = Lt + (size_t)nb * nb, *dinv = Sv + (size_t)nb * nb;
    double *Tmine = (double *)workspace("ch_Tmine", (size_t)nb * piece_max * sizeof(double));            // my piece of T (jb x count)
    double *Tall = (double *)workspace("ch_Tall", (size_t)nb * piece_max * np_along * sizeof(double));   // every piece, padded
    double *Tsel = (double *)workspace("ch_Tsel", (size_t)nb * (macross > 0 ? macross : 1) * sizeof(double));   // T for my indices across
    double *Aop = (double *)workspace("ch_Aop", (size_t)nb * (mloc > 0 ? mloc : 1) * sizeof(double));    // A operand: (my trailing rows) x jb
    int64_t *idx_dev = (int64_t *)workspace("ch_idx", (size_t)(macross > 0 ? macross : 1) * sizeof(int64_t));
    int *info_dev = (int *)workspace("ch_info", 64);
    SLB_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), s));
    std::vector<int64_t> idx;
    int info = 0;

    const int nblk = (N + nb - 1) / nb;
    for (int k = 0; k < nblk; ++k) {
        const int j0 = k * nb, jb = N - j0 < nb ? N - j0 : nb, t0 = j0 + jb;
        const int pr = (rsrc + k) % P, pc = (csrc + k) % Q;
        const int64_t ldr = numroc(j0, nb, myrow, rsrc, P), ldc = numroc(j0, nb, mycol, csrc, Q);   // local position of block k
        const int64_t lr0 = numroc(t0, nb, myrow, rsrc, P), lc0 = numroc(t0, nb, mycol, csrc, Q);   // local start of the trailing part
        const int64_t mr = mloc - lr0, ncl = nloc - lc0;
        double *Akk = A + ldr + ldc * lld;

        // ---- diagonal block (PDPOTF2) on its owner; INFO to everyone ----
        if (myrow == pr && mycol == pc) diag_potf2<UPPER>(jb, Akk, lld, info_dev, j0, s);
        if (multi) nccl_bcast(nc->all, info_dev, 1, NT_I32, pr * Q + pc, s);
        SLB_CUDA(cudaMemcpyAsync(&info, info_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
        SLB_CUDA(cudaStreamSynchronize(s));
        if (info != 0 || t0 >= N) break;

        // ---- panel: T = L21^T ('L', process column pc) or U12 ('U', process row pr) ----
        const bool in_panel = UPPER ? myrow == pr : mycol == pc;
        const int64_t cnt = UPPER ? ncl : mr;                   // my piece: jb x cnt
        if (in_panel) {
            if (UPPER ? mycol == pc : myrow == pr) launch_copy2d<double>(jb, jb, Akk, lld, Dbuf, jb, s);
            if (UPPER ? Q > 1 : P > 1) nccl_bcast(UPPER ? nc->row : nc->col, Dbuf, (size_t)jb * jb, NT_F64, UPPER ? pc : pr, s);
            SLB_LAUNCH((unit_lower_kernel<UPPER>), (unsigned)((jb * jb + 255) / 256), 256, s, jb, Dbuf, (int64_t)jb, Lt, dinv);
            if (cnt > 0) {
                if (UPPER) {
                    double *A12 = A + ldr + lc0 * lld;          // jb x ncl, already k-contiguous
                    launch_dtrsm_llnu(jb, cnt, Lt, jb, A12, lld, s);
                    SLB_LAUNCH(scale_rows_kernel, grid1d((int64_t)jb * cnt), 256, s, jb, cnt, A12, lld, (const double *)dinv);
                    launch_copy2d<double>(jb, cnt, A12, lld, Tmine, jb, s);
                } else {
                    double *A21 = A + lr0 + ldc * lld;          // mr x jb
                    SLB_LAUNCH(transpose_kernel, grid1d(cnt * jb), 256, s, cnt, (int64_t)jb, (const double *)A21, lld, Tmine, (int64_t)jb, (const double *)nullptr);
                    launch_dtrsm_llnu(jb, cnt, Lt, jb, Tmine, jb, s);
                    SLB_LAUNCH(scale_rows_kernel, grid1d((int64_t)jb * cnt), 256, s, jb, cnt, Tmine, (int64_t)jb, (const double *)dinv);
                    SLB_LAUNCH(transpose_kernel, grid1d(cnt * jb), 256, s, (int64_t)jb, cnt, (const double *)Tmine, (int64_t)jb, A21, lld, (const double *)nullptr);
                }
            }
        }
        // my piece to the processes that share my index range along the panel (row r for 'L', column c for 'U') ...
        if (cnt > 0 && (UPPER ? P > 1 : Q > 1)) nccl_bcast(UPPER ? nc->col : nc->row, Tmine, (size_t)jb * cnt, NT_F64, UPPER ? pr : pc, s);
        // ... and every piece to everyone across (padded to the longest piece)
        const double *Tsrc = Tmine;
        if (np_along > 1) { nccl_allgather(UPPER ? nc->row : nc->col, Tmine, Tall, (size_t)jb * piece_max, NT_F64, s); Tsrc = Tall; }

        // ---- T for my indices in the other dimension: 'L': my trailing local columns, 'U': my trailing local rows ----
        const int64_t nsel = UPPER ? mr : ncl, sel0 = UPPER ? lr0 : lc0;
        if (nsel > 0) {
            idx.resize((size_t)nsel);
            for (int64_t l = 0; l < nsel; ++l) {
                // global index of my local index sel0 + l across; where it sits in the pieces along
                const int gidx = UPPER ? indxl2g((int)(sel0 + l) + 1, nb, myrow, rsrc, P) - 1 : indxl2g((int)(sel0 + l) + 1, nb, mycol, csrc, Q) - 1;
                const int owner = UPPER ? indxg2p(gidx + 1, nb, csrc, Q) : indxg2p(gidx + 1, nb, rsrc, P);
                const int64_t lpos = (int64_t)indxg2l(gidx + 1, nb, UPPER ? Q : P) - 1 - (UPPER ? numroc(t0, nb, owner, csrc, Q) : numroc(t0, nb, owner, rsrc, P));
                idx[(size_t)l] = (np_along > 1 ? (int64_t)owner * piece_max : 0) + lpos;
            }
            SLB_CUDA(cudaMemcpyAsync(idx_dev, idx.data(), (size_t)nsel * sizeof(int64_t), cudaMemcpyHostToDevice, s));
            SLB_LAUNCH(gather_cols_kernel, grid1d((int64_t)jb * nsel), 256, s, jb, nsel, (const int64_t *)idx_dev, Tsrc, Tsel);
            SLB_CUDA(cudaStreamSynchronize(s));                 // idx is reused by the next step
        }
        // operands of the update: B = T over my trailing local columns (jb x ncl), A = (T over my trailing local rows)^T (mr x jb)
        const double *Bop = UPPER ? Tmine : Tsel;
        const double *Trow = UPPER ? Tsel : Tmine;
        if (mr > 0 && ncl > 0) {
            SLB_LAUNCH(transpose_kernel, grid1d(mr * jb), 256, s, (int64_t)jb, mr, Trow, (int64_t)jb, Aop, mr, (const double *)nullptr);
            // ---- trailing update (PDSYRK), one local block column at a time ----
            for (int64_t c = lc0; c < nloc;) {
                const int64_t w = (nb - c % nb) < (nloc - c) ? (nb - c % nb) : (nloc - c);
                const int J0 = indxl2g((int)c + 1, nb, mycol, csrc, Q) - 1;                  // global column of local column c
                const int64_t rd0 = numroc(J0, nb, myrow, rsrc, P), rd1 = numroc(J0 + (int)w, nb, myrow, rsrc, P);   // my rows of the diagonal block
                const int64_t rbeg = UPPER ? lr0 : rd0, rend = UPPER ? rd1 : mloc;           // rows on the triangle's side
                if (rend > rbeg) {
                    const bool has_diag = rd1 > rd0;
                    if (has_diag) launch_copy2d<double>(rd1 - rd0, w, A + rd0 + c * lld, lld, Sv, nb, s);
                    launch_dgemm_minus(rend - rbeg, w, jb, Aop + (rbeg - lr0), mr, Bop + (c - lc0) * jb, jb, A + rbeg + c * lld, lld, s);
                    if (has_diag) SLB_LAUNCH((restore_tri_kernel<UPPER>), (unsigned)((w * w + 255) / 256), 256, s, (int)w, (const double *)Sv, (int64_t)nb, A + rd0 + c * lld, lld);
                }
                c += w;
            }
        }
    }
    SLB_CUDA(cudaStreamSynchronize(s));
    *info_host = info;
}

void potrf_device(Grid *g, bool upper, int N, double *A, int64_t lld, int nb, int rsrc, int csrc, int *info_host)
{
    if (upper) potrf_sweep<true>(g, N, A, lld, nb, rsrc, csrc, info_host);
    else potrf_sweep<false>(g, N, A, lld, nb, rsrc, csrc, info_host);
}

}  // namespace slb

using namespace slb;

extern "C" void pdpotrf_(const char *uplo, const int *n, double *a, const int *ia, const int *ja, const int *desca, int *info)
{
    const int ictxt = desca[CTXT_];
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    const char u = (char)(uplo[0] & ~0x20);
    const bool upper = u == 'U';
    *info = 0;
    if (P == -1) *info = -(600 + CTXT_ + 1);
    else {
        chk1mat(*n, 2, *n, 2, *ia, *ja, desca, 6, info);
        if (*info == 0) {
            if (!upper && u != 'L') *info = -1;
            else if ((*ia - 1) % desca[MB_] != 0) *info = -4;
            else if ((*ja - 1) % desca[NB_] != 0) *info = -5;
            else if (desca[MB_] != desca[NB_]) *info = -(600 + NB_ + 1);
        }
        int ex[1] = { upper ? 'U' : 'L' }, expos[1] = { 1 }, one = 1, two = 2, six = 6;
        pchk1mat_(n, &two, n, &two, ia, ja, desca, &six, &one, ex, expos, info);
    }
    if (*info != 0) { xerbla(ictxt, "PDPOTRF", *info); return; }
    if (*n == 0) return;
    Grid *g = grid_of(ictxt);
    const Window w = window(*n, *n, *ia, *ja, desca, P, Q, myrow, mycol);
    StageMat<double> A("stage_A", a, desca[LLD_], w.loff_r, w.loff_c, w.mloc, w.nloc);
    potrf_device(g, upper, *n, A.dev, A.ld, desca[NB_], w.rsrc, w.csrc, info);
    A.download();
}
