// refine.cu -- SURVEY 8(f) row 1: the consumers of the LU factors around PDGESV.
//   PDLANGE  (SRC/pdlange.f:175-339)   norm of a distributed matrix (M, 1/O, I, F/E)
//   PDGEEQU  (SRC/pdgeequ.f:196-368)   row / column equilibration factors
//   PDLAQGE  (SRC/pdlaqge.f:190-271)   apply them
//   PDGECON  (SRC/pdgecon.f:218-414)   reciprocal condition estimate from the factors (PDLACON, SRC/pdlacon.f)
//   PDGERFS  (SRC/pdgerfs.f:300-887)   iterative refinement + forward / backward error bounds
//   PDGESVX  (SRC/pdgesvx.f:440-845)   the expert driver that strings them together with PDGETRF / PDGETRS
//
// Split of the work: everything that touches the N x N matrices runs on the GPU -- the two sweeps of every solve are
// getrs_device (solve.cu / solve_fast.cu), the residual b - op(A) x and the bound |op(A)| |x| + |b| are ONE fused pass
// over A (matvec_reduce_kernel), norms / equilibration factors are the same pass with a max- or sum-reduction.  The
// O(N) vectors between those passes (the reverse-communication logic of PDLACON, the componentwise error quotients)
// are replicated on every process and handled on the host, in global order, exactly once per process.
#include "common.h"
#include "dist.h"
#include "kernels.cuh"
#include "lacon.h"
#include "launch.h"
#include "lu.h"
#include "worklayout.h"

#include <cfloat>
#include <cmath>

namespace slb {

namespace {

const double EPS_ = DBL_EPSILON * 0.5;      // PDLAMCH( 'Epsilon' ): relative machine epsilon, 2^-53
const double SAFMIN_ = DBL_MIN;             // PDLAMCH( 'Safe minimum' ): 2^-1022 (1 / DBL_MAX < DBL_MIN)
const double PREC_ = DBL_EPSILON;           // PDLAMCH( 'Precision' ) = eps * base

// ---- one pass over the local window of a matrix with a reduction along rows or columns ----------------------------
// ROWS: out[il] = REDUCE over the local columns c of f(A[il, c], x[c])   (thread per local row: coalesced)
// else: out[c]  = REDUCE over the local rows i    of f(A[i, c],  x[i])   (thread per local column; a thread walks a
//       contiguous column, neighbouring threads are lld apart -- the sectors are re-used through L1)
// f / REDUCE by MODE: DOTABS: two sums, a*x and |a||x|;  MAX: max |a||x|;  SSQ: sum (|a| x)^2.  x == nullptr: x = 1.
// The reduced dimension is cut into gridDim.y slices; slice s writes part[s * len + idx] (DOTABS: the |.| sums follow at
// part[(gridDim.y + s) * len + idx]); the host adds the slices in order (deterministic).
enum { RM_DOTABS = 0, RM_MAX = 1, RM_SSQ = 2 };
template <int MODE, bool ROWS>
__global__ void __launch_bounds__(256)
matvec_reduce_kernel(int64_t mloc, int64_t nloc, const double *__restrict__ A, int64_t lda, const double *__restrict__ x,
                     double *__restrict__ part)
{
    const int64_t len = ROWS ? mloc : nloc, red = ROWS ? nloc : mloc;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= len) return;
    const int64_t nsl = gridDim.y, sl = blockIdx.y;
    const int64_t chunk = (red + nsl - 1) / nsl;
    const int64_t k0 = sl * chunk, k1 = (k0 + chunk < red) ? k0 + chunk : red;
    const double *ap = ROWS ? A + idx : A + idx * lda;
    const int64_t step = ROWS ? lda : 1;
    // four independent chains (k, k+1, k+2, k+3 modulo 4): four loads in flight per thread, a fixed summation order
    double acc[4] = { 0.0, 0.0, 0.0, 0.0 }, acc2[4] = { 0.0, 0.0, 0.0, 0.0 };
    for (int64_t kb = k0; kb < k1; kb += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t k = kb + u;
            if (k >= k1) break;
            const double a = ap[k * step];
            const double xv = x ? x[k] : 1.0;
            if (MODE == RM_DOTABS) { acc[u] = fma(a, xv, acc[u]); acc2[u] = fma(fabs(a), fabs(xv), acc2[u]); }
            else if (MODE == RM_MAX) { const double t = fabs(a) * fabs(xv); if (t > acc[u] || t != t) acc[u] = t; }
            else { const double t = fabs(a) * xv; acc[u] = fma(t, t, acc[u]); }
        }
    }
    double r0, r1;
    if (MODE == RM_MAX) {
        r0 = acc[0];
        for (int u = 1; u < 4; ++u) if (acc[u] > r0 || acc[u] != acc[u]) r0 = acc[u];
        r1 = 0.0;
    } else { r0 = (acc[0] + acc[1]) + (acc[2] + acc[3]); r1 = (acc2[0] + acc2[1]) + (acc2[2] + acc2[3]); }
    part[sl * len + idx] = r0;
    if (MODE == RM_DOTABS) part[(nsl + sl) * len + idx] = r1;
}

// A[il, c] *= (r ? r[il] : 1) * (c ? cs[c] : 1): the three branches of PDLAQGE (pdlaqge.f:223-262; CJ*R(I) first, then * A)
__global__ void __launch_bounds__(256)
scale_rc_kernel(int64_t mloc, int64_t nloc, double *__restrict__ A, int64_t lda, const double *__restrict__ r, const double *__restrict__ cs)
{
    const int64_t total = mloc * nloc;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t il = e % mloc, c = e / mloc;
        double f;
        if (r && cs) f = cs[c] * r[il]; else if (r) f = r[il]; else f = cs[c];
        A[il + c * lda] = f * A[il + c * lda];
    }
}

// out (global order, replicated, length = the non-reduced dimension of sub(A)) of one reduction pass.
//   rows = true : out[i] over global rows  i of sub(A), x given in global COLUMN order (or null)
//   rows = false: out[j] over global columns j,         x given in global ROW order (or null)
// sums are combined over the grid with '+', maxima with 'M'.  out2 (DOTABS only): the |a||x| sums.
struct Reducer {
    Grid *g; AnyWindow w; const double *Adev; int64_t lda; int m, n;
    Reducer(Grid *g_, const AnyWindow &w_, const double *Adev_, int64_t lda_, int m_, int n_) : g(g_), w(w_), Adev(Adev_), lda(lda_), m(m_), n(n_) {}

    void run(int mode, bool rows, const double *xglob, std::vector<double> &out, std::vector<double> *out2 = nullptr)
    {
        cudaStream_t s = rt().s_main;
        const int64_t len = rows ? w.mloc : w.nloc, red = rows ? w.nloc : w.mloc;
        const int glen = rows ? m : n;
        out.assign((size_t)glen, 0.0);
        if (out2) out2->assign((size_t)glen, 0.0);
        if (len > 0 && red > 0) {
            // slices: enough CTAs to fill the GPU, at least 512 reduced elements each
            int64_t nsl = (148 * 8) / ((len + 255) / 256);
            if (nsl > (red + 511) / 512) nsl = (red + 511) / 512;
            if (nsl < 1) nsl = 1;
            if (nsl > 64) nsl = 64;
            const int nout = mode == RM_DOTABS ? 2 : 1;
            double *part = (double *)workspace("rf_part", (size_t)nout * nsl * len * sizeof(double));
            double *xdev = nullptr;
            if (xglob) {
                std::vector<double> xl((size_t)red);
                for (int64_t k = 0; k < red; ++k) xl[(size_t)k] = xglob[rows ? w.gcol(k) : w.grow(k)];
                xdev = (double *)workspace("rf_x", (size_t)red * sizeof(double));
                SLB_CUDA(cudaMemcpyAsync(xdev, xl.data(), (size_t)red * sizeof(double), cudaMemcpyHostToDevice, s));
                SLB_CUDA(cudaStreamSynchronize(s));       // xl leaves scope
            }
            dim3 grid((unsigned)((len + 255) / 256), (unsigned)nsl), block(256);
            if (mode == RM_DOTABS) {
                if (rows) SLB_LAUNCH((matvec_reduce_kernel<RM_DOTABS, true>), grid, block, s, w.mloc, w.nloc, Adev, lda, xdev, part);
                else SLB_LAUNCH((matvec_reduce_kernel<RM_DOTABS, false>), grid, block, s, w.mloc, w.nloc, Adev, lda, xdev, part);
            } else if (mode == RM_MAX) {
                if (rows) SLB_LAUNCH((matvec_reduce_kernel<RM_MAX, true>), grid, block, s, w.mloc, w.nloc, Adev, lda, xdev, part);
                else SLB_LAUNCH((matvec_reduce_kernel<RM_MAX, false>), grid, block, s, w.mloc, w.nloc, Adev, lda, xdev, part);
            } else {
                if (rows) SLB_LAUNCH((matvec_reduce_kernel<RM_SSQ, true>), grid, block, s, w.mloc, w.nloc, Adev, lda, xdev, part);
                else SLB_LAUNCH((matvec_reduce_kernel<RM_SSQ, false>), grid, block, s, w.mloc, w.nloc, Adev, lda, xdev, part);
            }
            std::vector<double> hp((size_t)nout * nsl * len);
            SLB_CUDA(cudaMemcpyAsync(hp.data(), part, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
            SLB_CUDA(cudaStreamSynchronize(s));
            for (int64_t l = 0; l < len; ++l) {
                const int64_t gi = rows ? w.grow(l) : w.gcol(l);
                double a = hp[(size_t)l], a2 = nout == 2 ? hp[(size_t)(nsl * len + l)] : 0.0;
                for (int64_t sl = 1; sl < nsl; ++sl) {
                    const double b = hp[(size_t)(sl * len + l)];
                    if (mode == RM_MAX) { if (b > a || b != b) a = b; } else a += b;
                    if (nout == 2) a2 += hp[(size_t)((nsl + sl) * len + l)];
                }
                out[(size_t)gi] = a;
                if (out2) (*out2)[(size_t)gi] = a2;
            }
        }
        grid_combine(g, 'A', out.data(), out.size(), mode == RM_MAX ? 'M' : '+');
        if (out2) grid_combine(g, 'A', out2->data(), out2->size(), '+');
    }
};

double vmax(const std::vector<double> &v) { double a = 0.0; for (double e : v) if (e > a || e != e) a = e; return a; }

char up(const char *c) { return (char)(c[0] & ~0x20); }

// ---- PDLANGE ---------------------------------------------------------------------------------------------------------
double lange_impl(char norm, int m, int n, const double *a, int ia, int ja, const int *desca)
{
    const int ictxt = desca[CTXT_];
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    if (P == -1) return 0.0;
    if (m == 0 || n == 0) return 0.0;                                              // pdlange.f:196
    Grid *g = grid_of(ictxt);
    const AnyWindow w = any_window(m, n, ia, ja, desca, P, Q, myrow, mycol);
    StageMat<double> A("stage_A", a, desca[LLD_], w.loff_r, w.loff_c, w.mloc, w.nloc);
    Reducer R(g, w, A.dev, A.ld, m, n);
    std::vector<double> v, v2;
    if (norm == 'M') { R.run(RM_MAX, true, nullptr, v); return vmax(v); }          // max |a_ij|
    if (norm == 'O' || norm == '1') { R.run(RM_DOTABS, false, nullptr, v, &v2); return vmax(v2); }   // max column sum
    if (norm == 'I') { R.run(RM_DOTABS, true, nullptr, v, &v2); return vmax(v2); }                   // max row sum
    if (norm == 'F' || norm == 'E') {
        // DLASSQ keeps (scale, sumsq) to stay clear of overflow; the same value comes from scaling by max |a_ij| first
        R.run(RM_MAX, true, nullptr, v);
        const double amax = vmax(v);
        if (amax == 0.0 || amax != amax) return amax;
        std::vector<double> inv((size_t)n, 1.0 / amax);
        R.run(RM_SSQ, true, inv.data(), v);
        double ss = 0.0; for (double e : v) ss += e;
        return amax * sqrt(ss);
    }
    return 0.0;
}

// R / C of PDGEEQU and PDLAQGE are "aligned with the distributed matrix A": R(LOCr(M_A)) holds the entries of my local rows
// (replicated across process columns), C(LOCc(N_A)) those of my local columns.
void fill_local(const AnyWindow &w, bool rows, const std::vector<double> &glob, double *loc)
{
    const int64_t len = rows ? w.mloc : w.nloc, off = rows ? w.loff_r : w.loff_c;
    for (int64_t l = 0; l < len; ++l) loc[off + l] = glob[(size_t)(rows ? w.grow(l) : w.gcol(l))];
}

void geequ_impl(int m, int n, const double *a, int ia, int ja, const int *desca, double *r, double *c, double *rowcnd,
                double *colcnd, double *amax, int *info)
{
    const int ictxt = desca[CTXT_];
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    *info = 0;
    if (P == -1) *info = -(600 + CTXT_ + 1);
    else {
        chk1mat(m, 1, n, 2, ia, ja, desca, 6, info);
        int zero = 0, one = 1, two = 2, six = 6, idum = 0;
        pchk1mat_(&m, &one, &n, &two, &ia, &ja, desca, &six, &zero, &idum, &idum, info);
    }
    if (*info != 0) { xerbla(ictxt, "PDGEEQU", *info); return; }
    if (m == 0 || n == 0) { *rowcnd = 1.0; *colcnd = 1.0; *amax = 0.0; return; }
    Grid *g = grid_of(ictxt);
    const double smlnum = SAFMIN_, bignum = 1.0 / smlnum;
    const AnyWindow w = any_window(m, n, ia, ja, desca, P, Q, myrow, mycol);
    StageMat<double> A("stage_A", a, desca[LLD_], w.loff_r, w.loff_c, w.mloc, w.nloc);
    Reducer R(g, w, A.dev, A.ld, m, n);
    std::vector<double> rg, cg;
    R.run(RM_MAX, true, nullptr, rg);                                              // R(i) = max_j |a_ij|   (pdgeequ.f:244-252)
    double rcmin = bignum, rcmax = 0.0;
    for (double e : rg) { if (e > rcmax) rcmax = e; if (e < rcmin) rcmin = e; }
    *amax = rcmax;
    if (rcmin == 0.0) {
        fill_local(w, true, rg, r);
        // pdgeequ.f:268-275: every process row reports its first zero row, the column-wise IGAMX2D keeps the largest report
        std::vector<int> first((size_t)P, 0);
        for (int i = m - 1; i >= 0; --i) if (rg[(size_t)i] == 0.0) first[(size_t)indxg2p(ia + i, desca[MB_], desca[RSRC_], P)] = i + 1;
        for (int p = 0; p < P; ++p) if (first[(size_t)p] > *info) *info = first[(size_t)p];
        return;
    }
    for (double &e : rg) e = 1.0 / fmin(fmax(e, smlnum), bignum);
    *rowcnd = fmax(rcmin, smlnum) / fmin(rcmax, bignum);
    fill_local(w, true, rg, r);
    R.run(RM_MAX, false, rg.data(), cg);                                           // C(j) = max_i |a_ij| R(i)   (pdgeequ.f:291-299)
    rcmin = bignum; rcmax = 0.0;
    for (double e : cg) { if (e > rcmax) rcmax = e; if (e < rcmin) rcmin = e; }
    if (rcmin == 0.0) {
        fill_local(w, false, cg, c);
        // pdgeequ.f:318-325: the first zero column among the columns of MY process column (the combine there is column-wise too)
        for (int j = 0; j < n; ++j)
            if (cg[(size_t)j] == 0.0 && indxg2p(ja + j, desca[NB_], desca[CSRC_], Q) == mycol) { *info = m + j + 1; break; }
        return;
    }
    for (double &e : cg) e = 1.0 / fmin(fmax(e, smlnum), bignum);
    *colcnd = fmax(rcmin, smlnum) / fmin(rcmax, bignum);
    fill_local(w, false, cg, c);
}

void laqge_impl(int m, int n, double *a, int ia, int ja, const int *desca, const double *r, const double *c, double rowcnd,
                double colcnd, double amax, char *equed)
{
    const double THRESH = 0.1;                                                     // pdlaqge.f:164
    if (m <= 0 || n <= 0) { *equed = 'N'; return; }
    const int ictxt = desca[CTXT_];
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    if (P == -1) return;
    const double small_ = SAFMIN_ / PREC_, large_ = 1.0 / small_;
    const bool rowscale = !(rowcnd >= THRESH && amax >= small_ && amax <= large_);
    const bool colscale = !(colcnd >= THRESH);
    *equed = rowscale ? (colscale ? 'B' : 'R') : (colscale ? 'C' : 'N');
    if (!rowscale && !colscale) return;
    const AnyWindow w = any_window(m, n, ia, ja, desca, P, Q, myrow, mycol);
    if (w.mloc <= 0 || w.nloc <= 0) return;
    cudaStream_t s = rt().s_main;
    StageMat<double> A("stage_A", a, desca[LLD_], w.loff_r, w.loff_c, w.mloc, w.nloc);
    double *rc = (double *)workspace("rf_rc", (size_t)(w.mloc + w.nloc) * sizeof(double));
    SLB_CUDA(cudaMemcpyAsync(rc, r + w.loff_r, (size_t)w.mloc * sizeof(double), is_device_ptr(r) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    SLB_CUDA(cudaMemcpyAsync(rc + w.mloc, c + w.loff_c, (size_t)w.nloc * sizeof(double), is_device_ptr(c) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    const unsigned grid = grid1d(w.mloc * w.nloc);
    const double *rdev = rowscale ? rc : nullptr, *cdev = colscale ? rc + w.mloc : nullptr;
    SLB_LAUNCH(scale_rc_kernel, grid, 256, s, w.mloc, w.nloc, A.dev, A.ld, rdev, cdev);
    SLB_CUDA(cudaStreamSynchronize(s));
    A.download();
}

// ---- the factors as an operator on replicated N-vectors ----------------------------------------------------------------
struct Factors {
    Grid *g; int n, nb; Window w; const double *dev; int64_t ld; std::vector<int> ipiv, ident; double *xdev;
    Factors(Grid *g_, int n_, int nb_, const Window &w_, const double *dev_, int64_t ld_) : g(g_), n(n_), nb(nb_), w(w_), dev(dev_), ld(ld_)
    {
        ident.resize((size_t)n); for (int i = 0; i < n; ++i) ident[(size_t)i] = i + 1;
        xdev = (double *)workspace("rf_xrep", (size_t)n * sizeof(double));
    }
    // x <- op(A)^-1 x with the interchanges (piv = true: PDGETRS) or op(L U)^-1 x without them (PDGECON's two PDLATRS calls)
    void solve(char trans, double *x, bool piv)
    {
        cudaStream_t s = rt().s_main;
        SLB_CUDA(cudaMemcpyAsync(xdev, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
        getrs_device<double>(g, trans, n, 1, dev, ld, nb, w.rsrc, w.csrc, piv ? ipiv.data() : ident.data(), nullptr, 1, nb, 0, 0, 0, xdev, xdev);
        SLB_CUDA(cudaMemcpyAsync(x, xdev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
        SLB_CUDA(cudaStreamSynchronize(s));
    }
};

// ---- PDGECON -----------------------------------------------------------------------------------------------------------
int gecon_lwmin(int n, int ia, int ja, const int *desca, int P, int Q, int myrow, int mycol, int *liwmin)
{
    const int mb = desca[MB_], nb = desca[NB_];
    const int iarow = indxg2p(ia, mb, desca[RSRC_], P), iacol = indxg2p(ja, nb, desca[CSRC_], Q);
    const int npmod = numroc(n + (ia - 1) % mb, mb, myrow, iarow, P), nqmod = numroc(n + (ja - 1) % nb, nb, mycol, iacol, Q);
    auto iceil = [](int a, int b) { return (a + b - 1) / b; };
    auto imax = [](int a, int b) { return a > b ? a : b; };
    *liwmin = imax(1, npmod);
    return 2 * npmod + 2 * nqmod + imax(2, imax(nb * imax(1, iceil(P - 1, Q)), nqmod + nb * imax(1, iceil(Q - 1, P))));
}

void gecon_core(Grid *g, bool onenrm, int n, int nb, const Window &w, const double *AFdev, int64_t ld, double anorm, double *rcond)
{
    Factors F(g, n, nb, w, AFdev, ld);
    // kase 1: x <- inv(U) inv(L) x; kase 2: x <- inv(L') inv(U') x; the infinity norm swaps them (pdgecon.f:333-373)
    const double ainvnm = lacon_estimate(n, [&](double *x, int kase) { F.solve((kase == 1) == onenrm ? 'N' : 'T', x, false); });
    if (ainvnm != 0.0) *rcond = (1.0 / ainvnm) / anorm;
}

void gecon_impl(const char *norm, int n, const double *a, int ia, int ja, const int *desca, double anorm, double *rcond,
                double *work, int lwork, int *iwork, int liwork, int *info)
{
    const int ictxt = desca[CTXT_];
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    *info = 0;
    bool onenrm = false, lquery = false;
    if (P == -1) *info = -(600 + CTXT_ + 1);
    else {
        chk1mat(n, 2, n, 2, ia, ja, desca, 6, info);
        if (*info == 0) {
            onenrm = norm[0] == '1' || up(norm) == 'O';
            int liwmin, lwmin = gecon_lwmin(n, ia, ja, desca, P, Q, myrow, mycol, &liwmin);
            work[0] = (double)lwmin; iwork[0] = liwmin;
            lquery = lwork == -1 || liwork == -1;
            if (!onenrm && up(norm) != 'I') *info = -1;
            else if (anorm < 0.0) *info = -7;
            else if (lwork < lwmin && !lquery) *info = -10;
            else if (liwork < liwmin && !lquery) *info = -12;
        }
        int ex[3] = { onenrm ? '1' : 'I', lwork == -1 ? -1 : 1, liwork == -1 ? -1 : 1 }, expos[3] = { 1, 10, 12 };
        int two = 2, six = 6, three = 3;
        pchk1mat_(&n, &two, &n, &two, &ia, &ja, desca, &six, &three, ex, expos, info);
    }
    if (*info != 0) { xerbla(ictxt, "PDGECON", *info); return; }
    if (lquery) return;
    *rcond = 0.0;
    if (n == 0) { *rcond = 1.0; return; }
    if (anorm == 0.0) return;
    if (n == 1) { *rcond = 1.0; return; }
    Grid *g = grid_of(ictxt);
    if ((ia - 1) % desca[MB_] || (ja - 1) % desca[NB_] || desca[MB_] != desca[NB_]) {
        // the reference's PDTRSV takes the factors at any alignment; the solves here want block-aligned square blocks: work on an
        // aligned copy of sub(A) (one redistribution)
        const int nbw = desca[NB_] < desca[MB_] ? desca[NB_] : desca[MB_];
        Work F("pb_A", g, n, n, nbw);
        F.load(a, ia, ja, desca, false);
        SLB_CUDA(cudaStreamSynchronize(rt().s_main));
        Window wf; wf.loff_r = wf.loff_c = 0; wf.mloc = F.mloc; wf.nloc = F.nloc; wf.rsrc = wf.csrc = 0;
        gecon_core(g, onenrm, n, nbw, wf, F.dev, F.ld, anorm, rcond);
        return;
    }
    const Window w = window(n, n, ia, ja, desca, P, Q, myrow, mycol);
    StageMat<double> AF("stage_A", a, desca[LLD_], w.loff_r, w.loff_c, w.mloc, w.nloc);
    gecon_core(g, onenrm, n, desca[NB_], w, AF.dev, AF.ld, anorm, rcond);
}

// ---- PDGERFS -----------------------------------------------------------------------------------------------------------
// One right-hand side at a time like the reference (pdgerfs.f:497-660): b, x replicated on the host in global order.
void gerfs_core(Grid *g, char trans, int n, int nrhs, const AnyWindow &wa, const double *Adev, int64_t lda, Factors &F,
                const std::vector<double> &bg, std::vector<double> &xg, std::vector<double> &ferr, std::vector<double> &berr)
{
    const int ITMAX = 5;                                                           // pdgerfs.f:262
    const bool notran = trans == 'N';
    const char transt = notran ? 'T' : 'N';
    const int nz = n + 1;
    const double safe1 = nz * SAFMIN_, safe2 = safe1 / EPS_;
    Reducer R(g, wa, Adev, lda, n, n);
    std::vector<double> ax, aax, r((size_t)n), wk((size_t)n);
    ferr.assign((size_t)nrhs, 0.0); berr.assign((size_t)nrhs, 0.0);
    for (int k = 0; k < nrhs; ++k) {
        const double *b = bg.data() + (size_t)k * n;
        double *x = xg.data() + (size_t)k * n;
        int count = 1;
        double lstres = 3.0, s = 0.0;
        for (;;) {
            // r = b - op(A) x,  wk = |op(A)| |x| + |b|   (PDGEMV + PDAGEMV, pdgerfs.f:503-517)
            R.run(RM_DOTABS, notran, x, ax, &aax);
            for (int i = 0; i < n; ++i) { r[(size_t)i] = b[i] - ax[(size_t)i]; wk[(size_t)i] = fabs(b[i]) + aax[(size_t)i]; }
            s = 0.0;
            for (int i = 0; i < n; ++i) {
                const double q = wk[(size_t)i] > safe2 ? fabs(r[(size_t)i]) / wk[(size_t)i] : (fabs(r[(size_t)i]) + safe1) / (wk[(size_t)i] + safe1);
                if (q > s || q != q) s = q;
            }
            berr[(size_t)k] = s;
            if (s > EPS_ && 2.0 * s <= lstres && count <= ITMAX) {                 // pdgerfs.f:544-554
                F.solve(trans, r.data(), true);
                for (int i = 0; i < n; ++i) x[i] += r[(size_t)i];
                lstres = s; ++count;
                continue;
            }
            break;
        }
        // forward error bound: || inv(op(A)) diag(wk) ||_inf by PDLACON on its transpose (pdgerfs.f:556-615)
        for (int i = 0; i < n; ++i)
            wk[(size_t)i] = wk[(size_t)i] > safe2 ? fabs(r[(size_t)i]) + nz * EPS_ * wk[(size_t)i] : fabs(r[(size_t)i]) + nz * EPS_ * wk[(size_t)i] + safe1;
        const double est = lacon_estimate(n, [&](double *v, int kase) {
            if (kase == 1) { F.solve(transt, v, true); for (int i = 0; i < n; ++i) v[i] = wk[(size_t)i] * v[i]; }
            else { for (int i = 0; i < n; ++i) v[i] = wk[(size_t)i] * v[i]; F.solve(trans, v, true); }
        });
        double xmax = 0.0;
        for (int i = 0; i < n; ++i) if (fabs(x[i]) > xmax) xmax = fabs(x[i]);
        if (xmax != 0.0) ferr[(size_t)k] = est / xmax;
    }
}

// FERR / BERR are local arrays over the local columns of B (pdgerfs.f:213-229): entry of global column JB+k on the process
// column that owns it, at that column's local index
void store_err(int nrhs, int jb, const int *descb, int Q, int mycol, const std::vector<double> &v, double *out, double scale = 1.0)
{
    for (int k = 0; k < nrhs; ++k)
        if (indxg2p(jb + k, descb[NB_], descb[CSRC_], Q) == mycol) out[indxg2l(jb + k, descb[NB_], Q) - 1] = v[(size_t)k] / scale;
}

// the argument checks of PDGERFS (pdgerfs.f:304-452): INFO (0 = fine); *lquery: a workspace query.  Also what PDGESVX ends up
// reporting for its B / X arguments: the expert driver checks little of them itself and returns the INFO of its last inner call.
void gerfs_checks(char trans, int n, int nrhs, int ia, int ja, const int *desca, int iaf, int jaf, const int *descaf, int ib, int jb,
                  const int *descb, int ix, int jx, const int *descx, double *work, int lwork, int *iwork, int liwork, bool *lquery_out, int *info)
{
    const int ictxt = desca[CTXT_];
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    const bool notran = trans == 'N';
    bool lquery = false;
    *info = 0;
    if (P == -1) *info = -(700 + CTXT_ + 1);
    else {
        chk1mat(n, 2, n, 2, ia, ja, desca, 7, info);
        chk1mat(n, 2, n, 2, iaf, jaf, descaf, 11, info);
        chk1mat(n, 2, nrhs, 3, ib, jb, descb, 16, info);
        chk1mat(n, 2, nrhs, 3, ix, jx, descx, 20, info);
        if (*info == 0) {
            const int iroffa = (ia - 1) % desca[MB_], icoffa = (ja - 1) % desca[NB_], iroffaf = (iaf - 1) % descaf[MB_],
                      icoffaf = (jaf - 1) % descaf[NB_], iroffb = (ib - 1) % descb[MB_], icoffb = (jb - 1) % descb[NB_],
                      iroffx = (ix - 1) % descx[MB_], icoffx = (jx - 1) % descx[NB_];
            const int iarow = indxg2p(ia, desca[MB_], desca[RSRC_], P), iafcol = indxg2p(jaf, descaf[NB_], descaf[CSRC_], Q),
                      iafrow = indxg2p(iaf, descaf[MB_], descaf[RSRC_], P), iacol = indxg2p(ja, desca[NB_], desca[CSRC_], Q),
                      ixbrow = indxg2p(ib, descb[MB_], descb[RSRC_], P), ixbcol = indxg2p(jb, descb[NB_], descb[CSRC_], Q),
                      ixrow = indxg2p(ix, descx[MB_], descx[RSRC_], P), ixcol = indxg2p(jx, descx[NB_], descx[CSRC_], Q);
            const int npmod = numroc(n + iroffa, desca[MB_], myrow, iarow, P);
            const int lwmin = 3 * npmod, liwmin = npmod;
            work[0] = (double)lwmin; iwork[0] = liwmin;
            lquery = lwork == -1 || liwork == -1;
            if (!notran && trans != 'T' && trans != 'C') *info = -1;
            else if (n < 0) *info = -2;
            else if (nrhs < 0) *info = -3;
            else if (iroffa != 0) *info = -5;
            else if (icoffa != 0) *info = -6;
            else if (desca[MB_] != desca[NB_]) *info = -(700 + NB_ + 1);
            else if (desca[MB_] != descaf[MB_]) *info = -(1100 + MB_ + 1);
            else if (iroffaf != 0 || iarow != iafrow) *info = -9;
            else if (desca[NB_] != descaf[NB_]) *info = -(1100 + NB_ + 1);
            else if (icoffaf != 0 || iacol != iafcol) *info = -10;
            else if (ictxt != descaf[CTXT_]) *info = -(1100 + CTXT_ + 1);
            else if (iroffa != iroffb || iarow != ixbrow) *info = -14;
            else if (desca[MB_] != descb[MB_]) *info = -(1600 + MB_ + 1);
            else if (ictxt != descb[CTXT_]) *info = -(1600 + CTXT_ + 1);
            else if (descb[MB_] != descx[MB_]) *info = -(2000 + MB_ + 1);
            else if (iroffx != 0 || ixbrow != ixrow) *info = -18;
            else if (descb[NB_] != descx[NB_]) *info = -(2000 + NB_ + 1);
            else if (icoffb != icoffx || ixbcol != ixcol) *info = -19;
            else if (ictxt != descx[CTXT_]) *info = -(2000 + CTXT_ + 1);
            else if (lwork < lwmin && !lquery) *info = -24;
            else if (liwork < liwmin && !lquery) *info = -26;
        }
        int ex[5] = { notran ? 'N' : (trans == 'T' ? 'T' : 'C'), n, nrhs, lwork == -1 ? -1 : 1, liwork == -1 ? -1 : 1 };
        int expos[5] = { 1, 2, 3, 24, 26 };
        int two = 2, three = 3, five = 5, p7 = 7, p11 = 11, p16 = 16, p20 = 20;
        pchk2mat_(&n, &two, &n, &two, &ia, &ja, desca, &p7, &n, &two, &n, &two, &iaf, &jaf, descaf, &p11, &five, ex, expos, info);
        pchk2mat_(&n, &two, &nrhs, &three, &ib, &jb, descb, &p16, &n, &two, &nrhs, &three, &ix, &jx, descx, &p20, &five, ex, expos, info);
    }
    *lquery_out = lquery;
}

void gerfs_impl(const char *trans_, int n, int nrhs, const double *a, int ia, int ja, const int *desca, const double *af, int iaf,
                int jaf, const int *descaf, const int *ipiv, const double *b, int ib, int jb, const int *descb, double *x, int ix,
                int jx, const int *descx, double *ferr, double *berr, double *work, int lwork, int *iwork, int liwork, int *info)
{
    const int ictxt = desca[CTXT_];
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    const char trans = up(trans_);
    bool lquery = false;
    gerfs_checks(trans, n, nrhs, ia, ja, desca, iaf, jaf, descaf, ib, jb, descb, ix, jx, descx, work, lwork, iwork, liwork, &lquery, info);
    if (*info != 0) { xerbla(ictxt, "PDGERFS", *info); return; }
    if (lquery) return;
    if (n <= 1 || nrhs == 0) {                                                     // pdgerfs.f:457-463
        const int jjfbe = numroc(jb - 1, descb[NB_], mycol, descb[CSRC_], Q), myrhs = numroc(jb + nrhs - 1, descb[NB_], mycol, descb[CSRC_], Q);
        for (int jj = jjfbe; jj < myrhs; ++jj) { ferr[jj] = 0.0; berr[jj] = 0.0; }
        return;
    }
    Grid *g = grid_of(ictxt);
    const Window waf = window(n, n, iaf, jaf, descaf, P, Q, myrow, mycol);
    const AnyWindow wa = any_window(n, n, ia, ja, desca, P, Q, myrow, mycol);
    StageMat<double> A("stage_A2", a, desca[LLD_], wa.loff_r, wa.loff_c, wa.mloc, wa.nloc);
    StageMat<double> AF("stage_A", af, descaf[LLD_], waf.loff_r, waf.loff_c, waf.mloc, waf.nloc);
    Factors F(g, n, desca[NB_], waf, AF.dev, AF.ld);
    gather_global_ipiv(g, n, descaf[NB_], waf.rsrc, ipiv + waf.loff_r, iaf - 1, F.ipiv);
    std::vector<double> bg, xg, fe, be;
    gather_small(g, n, nrhs, b, ib, jb, descb, bg);
    gather_small(g, n, nrhs, x, ix, jx, descx, xg);
    gerfs_core(g, trans == 'C' ? 'T' : trans, n, nrhs, wa, A.dev, A.ld, F, bg, xg, fe, be);
    scatter_small(g, n, nrhs, x, ix, jx, descx, xg);
    store_err(nrhs, jb, descb, Q, mycol, fe, ferr);
    store_err(nrhs, jb, descb, Q, mycol, be, berr);
}

// ---- PDGESVX -----------------------------------------------------------------------------------------------------------
void gesvx_impl(const char *fact_, const char *trans_, int n, int nrhs, double *a, int ia, int ja, const int *desca, double *af,
                int iaf, int jaf, const int *descaf, int *ipiv, char *equed, double *r, double *c, double *b, int ib, int jb,
                const int *descb, double *x, int ix, int jx, const int *descx, double *rcond, double *ferr, double *berr,
                double *work, int lwork, int *iwork, int liwork, int *info)
{
    const int ictxt = desca[CTXT_];
    int P, Q, myrow, mycol; blacs_gridinfo_(&ictxt, &P, &Q, &myrow, &mycol);
    const char fact = up(fact_), trans = up(trans_);
    const bool nofact = fact == 'N', equil = fact == 'E', notran = trans == 'N';
    bool rowequ = false, colequ = false, lquery = false;
    double rowcnd = 1.0, colcnd = 1.0, amax = 0.0;
    const double smlnum = SAFMIN_, bignum = 1.0 / smlnum;
    int lwmin = 0, liwmin = 0;
    *info = 0;
    if (P == -1) *info = -(800 + CTXT_ + 1);
    else {
        chk1mat(n, 3, n, 3, ia, ja, desca, 8, info);
        if (fact == 'F') chk1mat(n, 3, n, 3, iaf, jaf, descaf, 12, info);
        chk1mat(n, 3, nrhs, 4, ib, jb, descb, 20, info);
        chk1mat(n, 3, nrhs, 4, ix, jx, descx, 24, info);
        if (nofact || equil) *equed = 'N';
        else { const char e = up(equed); rowequ = e == 'R' || e == 'B'; colequ = e == 'C' || e == 'B'; }
        int ibrow = 0, iarow = 0, ixrow = 0;
        if (*info == 0) {
            iarow = indxg2p(ia, desca[MB_], desca[RSRC_], P);
            // AF's descriptor is only checked under FACT = 'F' (pdgesvx.f:453-455); a zero block size there is a division by zero in
            // the reference (pdgesvx.f:468-470).  Here it falls through to the inner PDGETRF's own check below.
            const int mbaf = descaf[MB_] > 0 ? descaf[MB_] : 1;
            const int iafrow = descaf[MB_] > 0 ? indxg2p(iaf, mbaf, descaf[RSRC_], P) : iarow;
            ibrow = indxg2p(ib, descb[MB_], descb[RSRC_], P); ixrow = indxg2p(ix, descx[MB_], descx[RSRC_], P);
            const int iroffa = (ia - 1) % desca[MB_], iroffaf = descaf[MB_] > 0 ? (iaf - 1) % mbaf : 0, icoffa = (ja - 1) % desca[NB_];
            const int iacol = indxg2p(ja, desca[NB_], desca[CSRC_], Q);
            int np = numroc(n + iroffa, desca[MB_], myrow, iarow, P); if (myrow == iarow) np -= iroffa;
            int nq = numroc(n + icoffa, desca[NB_], mycol, iacol, Q); if (mycol == iacol) nq -= icoffa;
            auto iceil = [](int a_, int b_) { return (a_ + b_ - 1) / b_; };
            auto imax = [](int a_, int b_) { return a_ > b_ ? a_ : b_; };
            const int nqb = iceil(n + iroffa, desca[NB_] * Q);
            int lcm = P; { int a_ = P, b_ = Q; while (b_) { int t = a_ % b_; a_ = b_; b_ = t; } lcm = P / a_ * Q; }
            const int lcmq = lcm / Q;
            const int conwrk = 2 * np + 2 * nq + imax(2, imax(desca[NB_] * imax(1, iceil(P - 1, Q)), nq + desca[NB_] * imax(1, iceil(Q - 1, P))));
            int rfswrk = 3 * np;
            if (trans == 'N') rfswrk += np + nq + iceil(nqb, lcmq) * desca[NB_];
            else if (trans == 'T' || trans == 'C') rfswrk += np + nq;
            lwmin = imax(conwrk, rfswrk); liwmin = np;
            work[0] = (double)lwmin; iwork[0] = liwmin;
            if (!nofact && !equil && fact != 'F') *info = -1;
            else if (!notran && trans != 'T' && trans != 'C') *info = -2;
            else if (iroffa != 0) *info = -6;
            else if (icoffa != 0 || iroffa != icoffa) *info = -7;
            else if (desca[MB_] != desca[NB_]) *info = -(800 + NB_ + 1);
            else if (iafrow != iarow) *info = -10;
            else if (iroffaf != 0) *info = -10;
            else if (ictxt != descaf[CTXT_]) *info = -(1200 + CTXT_ + 1);
            else if (fact == 'F' && !(rowequ || colequ || up(equed) == 'N')) *info = -13;
            else {
                const Window w = window(n, n, ia, ja, desca, P, Q, myrow, mycol);
                if (rowequ) {                                                     // pdgesvx.f:535-553
                    double v[2] = { bignum, 0.0 };
                    for (int64_t j = 0; j < w.mloc; ++j) { v[0] = fmin(v[0], r[w.loff_r + j]); v[1] = fmax(v[1], r[w.loff_r + j]); }
                    Grid *g = grid_of(ictxt);
                    grid_combine(g, 'C', &v[0], 1, 'm'); grid_combine(g, 'C', &v[1], 1, 'M');
                    if (v[0] <= 0.0) *info = -14;
                    else rowcnd = n > 0 ? fmax(v[0], smlnum) / fmin(v[1], bignum) : 1.0;
                }
                if (colequ && *info == 0) {                                       // pdgesvx.f:554-573
                    double v[2] = { bignum, 0.0 };
                    for (int64_t j = 0; j < w.nloc; ++j) { v[0] = fmin(v[0], c[w.loff_c + j]); v[1] = fmax(v[1], c[w.loff_c + j]); }
                    Grid *g = grid_of(ictxt);
                    grid_combine(g, 'R', &v[0], 1, 'm'); grid_combine(g, 'R', &v[1], 1, 'M');
                    if (v[0] <= 0.0) *info = -15;
                    else colcnd = n > 0 ? fmax(v[0], smlnum) / fmin(v[1], bignum) : 1.0;
                }
            }
        }
        lquery = lwork == -1 || liwork == -1;
        if (*info == 0) {
            if (ibrow != iarow) *info = -18;
            else if (ixrow != ibrow) *info = -22;
            else if (descb[MB_] != desca[NB_]) *info = -(2000 + NB_ + 1);
            else if (ictxt != descb[CTXT_]) *info = -(2000 + CTXT_ + 1);
            else if (descx[MB_] != desca[NB_]) *info = -(2400 + NB_ + 1);
            else if (ictxt != descx[CTXT_]) *info = -(2400 + CTXT_ + 1);
            else if (lwork < lwmin && !lquery) *info = -29;
            else if (liwork < liwmin && !lquery) *info = -31;
            int ex[5], expos[5], nex;
            ex[0] = fact_[0]; expos[0] = 1; ex[1] = trans_[0]; expos[1] = 2;
            if (fact == 'F') { ex[2] = equed[0]; expos[2] = 14; ex[3] = lwork == -1 ? -1 : 1; expos[3] = 29; ex[4] = liwork == -1 ? -1 : 1; expos[4] = 31; nex = 5; }
            else { ex[2] = lwork == -1 ? -1 : 1; expos[2] = 29; ex[3] = liwork == -1 ? -1 : 1; expos[3] = 31; nex = 4; }
            int three = 3, four = 4, p8 = 8, p20 = 20;
            pchk2mat_(&n, &three, &n, &three, &ia, &ja, desca, &p8, &n, &three, &nrhs, &four, &ib, &jb, descb, &p20, &nex, ex, expos, info);
        }
    }
    if (*info != 0) { xerbla(ictxt, "PDGESVX", *info); return; }
    if (lquery) return;
    // what the reference's inner calls would report (PDGETRF on AF: pdgetrf.f:180-185; PDGETRS on X: pdgetrs.f:211-222)
    {
        int i2 = 0;
        if (nofact || equil) chk1mat(n, 1, n, 2, iaf, jaf, descaf, 6, &i2);              // pdgetrf.f:169
        if (i2 == 0) {
            if ((jaf - 1) % descaf[NB_]) i2 = -5;
            else if (descaf[MB_] != descaf[NB_] || descaf[NB_] != desca[NB_]) i2 = -(600 + NB_ + 1);
        }
        if (i2 != 0) { *info = i2; xerbla(ictxt, "PDGETRF", *info); return; }
    }
    {
        // PDGESVX does not look at INFO between PDGETRS and PDGERFS (pdgesvx.f:749-757): what the caller gets for a misplaced B or X is
        // what PDGERFS, the last inner call, reports -- e.g. -14 for a B that starts inside a block, -19 for JB / JX inside one
        std::vector<double> w1(1); std::vector<int> iw1(1); bool q = false; int i3 = 0;
        gerfs_checks(trans, n, nrhs, ia, ja, desca, iaf, jaf, descaf, ib, jb, descb, ix, jx, descx, w1.data(), lwork, iw1.data(), liwork, &q, &i3);
        if (i3 != 0) { *info = i3; xerbla(ictxt, "PDGERFS", *info); return; }
    }

    Grid *g = grid_of(ictxt);
    const int nb = desca[NB_];
    if (equil) {                                                                   // pdgesvx.f:660-672
        int infequ = 0;
        geequ_impl(n, n, a, ia, ja, desca, r, c, &rowcnd, &colcnd, &amax, &infequ);
        if (infequ == 0) {
            laqge_impl(n, n, a, ia, ja, desca, r, c, rowcnd, colcnd, amax, equed);
            rowequ = *equed == 'R' || *equed == 'B'; colequ = *equed == 'C' || *equed == 'B';
        }
    }
    // scale factors replicated in global order (the reference re-distributes R / C with PDCOPY + a row broadcast)
    const AnyWindow wa = any_window(n, n, ia, ja, desca, P, Q, myrow, mycol);
    std::vector<double> rg((size_t)n, 0.0), cg((size_t)n, 0.0);
    if (rowequ) { for (int64_t l = 0; l < wa.mloc; ++l) rg[(size_t)wa.grow(l)] = r[wa.loff_r + l]; grid_combine(g, 'C', rg.data(), rg.size(), 'M'); }
    if (colequ) { for (int64_t l = 0; l < wa.nloc; ++l) cg[(size_t)wa.gcol(l)] = c[wa.loff_c + l]; grid_combine(g, 'R', cg.data(), cg.size(), 'M'); }
    // scale the right-hand sides (pdgesvx.f:683-711)
    std::vector<double> bg;
    gather_small(g, n, nrhs, b, ib, jb, descb, bg);
    const std::vector<double> *bs = notran ? (rowequ ? &rg : nullptr) : (colequ ? &cg : nullptr);
    if (bs) {
        for (int k = 0; k < nrhs; ++k) for (int i = 0; i < n; ++i) bg[(size_t)i + (size_t)k * n] = (*bs)[(size_t)i] * bg[(size_t)i + (size_t)k * n];
        scatter_small(g, n, nrhs, b, ib, jb, descb, bg);
    }
    const Window w = window(n, n, ia, ja, desca, P, Q, myrow, mycol);
    const Window waf = window(n, n, iaf, jaf, descaf, P, Q, myrow, mycol);
    StageMat<double> A("stage_A2", a, desca[LLD_], w.loff_r, w.loff_c, w.mloc, w.nloc);
    StageMat<double> AF("stage_A", af, descaf[LLD_], waf.loff_r, waf.loff_c, waf.mloc, waf.nloc, !(nofact || equil));
    Factors F(g, n, nb, waf, AF.dev, AF.ld);
    if (nofact || equil) {                                                         // pdgesvx.f:713-722
        if (w.mloc > 0 && w.nloc > 0) launch_copy2d<double>(w.mloc, w.nloc, A.dev, A.ld, AF.dev, AF.ld, rt().s_main);
        SLB_CUDA(cudaStreamSynchronize(rt().s_main));
        F.ipiv.assign((size_t)n, 0);
        if (descaf[M_] == 1) { ipiv[0] = 1; F.ipiv[0] = 1; }                       // PDGETRF's quick return (pdgetrf.f:201-203)
        else {
            getrf_device<double>(g, n, n, AF.dev, AF.ld, nb, waf.rsrc, waf.csrc, F.ipiv.data(), info, nullptr);
            fill_local_ipiv(F.ipiv, n, nb, waf.rsrc, P, myrow, ipiv + waf.loff_r, iaf - 1);
        }
        AF.download();
        if (*info != 0) { if (*info > 0) *rcond = 0.0; return; }
    } else gather_global_ipiv(g, n, nb, waf.rsrc, ipiv + waf.loff_r, iaf - 1, F.ipiv);

    // ||A|| and the condition estimate (pdgesvx.f:726-741)
    Reducer R(g, wa, A.dev, A.ld, n, n);
    std::vector<double> v, v2;
    R.run(RM_DOTABS, !notran, nullptr, v, &v2);                                    // '1': column sums, 'I': row sums
    const double anorm = vmax(v2);
    *rcond = 0.0;
    if (n == 0) *rcond = 1.0;
    else if (anorm == 0.0) *rcond = 0.0;
    else if (n == 1) *rcond = 1.0;
    else gecon_core(g, notran, n, nb, waf, AF.dev, AF.ld, anorm, rcond);
    if (*rcond < EPS_) { *info = ia + n; return; }                                 // pdgesvx.f:738-741 (sic: IA + N)

    // x = op(A)^-1 b, refined (pdgesvx.f:745-757)
    std::vector<double> xg = bg, fe, be;
    for (int k = 0; k < nrhs; ++k) F.solve(trans == 'C' ? 'T' : trans, xg.data() + (size_t)k * n, true);
    if (n > 1 && nrhs > 0) gerfs_core(g, trans == 'C' ? 'T' : trans, n, nrhs, wa, A.dev, A.ld, F, bg, xg, fe, be);
    else { fe.assign((size_t)nrhs, 0.0); be.assign((size_t)nrhs, 0.0); }
    // undo the scaling of the solution (pdgesvx.f:771-812)
    double fscale = 1.0;
    const std::vector<double> *xs = notran ? (colequ ? &cg : nullptr) : (rowequ ? &rg : nullptr);
    if (xs) {
        for (int k = 0; k < nrhs; ++k) for (int i = 0; i < n; ++i) xg[(size_t)i + (size_t)k * n] = (*xs)[(size_t)i] * xg[(size_t)i + (size_t)k * n];
        fscale = notran ? colcnd : rowcnd;
    }
    scatter_small(g, n, nrhs, x, ix, jx, descx, xg);
    store_err(nrhs, jb, descb, Q, mycol, fe, ferr, fscale);
    store_err(nrhs, jb, descb, Q, mycol, be, berr);
    work[0] = (double)lwmin; iwork[0] = liwmin;
}

}  // namespace

}  // namespace slb

using namespace slb;

extern "C" {

double pdlange_(const char *norm, const int *m, const int *n, const double *a, const int *ia, const int *ja, const int *desca, double *work)
{ (void)work; return lange_impl(up(norm) == 'O' ? '1' : (norm[0] == '1' ? '1' : up(norm)), *m, *n, a, *ia, *ja, desca); }

void pdgeequ_(const int *m, const int *n, const double *a, const int *ia, const int *ja, const int *desca, double *r, double *c,
              double *rowcnd, double *colcnd, double *amax, int *info)
{ geequ_impl(*m, *n, a, *ia, *ja, desca, r, c, rowcnd, colcnd, amax, info); }

void pdlaqge_(const int *m, const int *n, double *a, const int *ia, const int *ja, const int *desca, const double *r, const double *c,
              const double *rowcnd, const double *colcnd, const double *amax, char *equed)
{ laqge_impl(*m, *n, a, *ia, *ja, desca, r, c, *rowcnd, *colcnd, *amax, equed); }

void pdgecon_(const char *norm, const int *n, const double *a, const int *ia, const int *ja, const int *desca, const double *anorm,
              double *rcond, double *work, const int *lwork, int *iwork, const int *liwork, int *info)
{ gecon_impl(norm, *n, a, *ia, *ja, desca, *anorm, rcond, work, *lwork, iwork, *liwork, info); }

void pdgerfs_(const char *trans, const int *n, const int *nrhs, const double *a, const int *ia, const int *ja, const int *desca,
              const double *af, const int *iaf, const int *jaf, const int *descaf, const int *ipiv, const double *b, const int *ib,
              const int *jb, const int *descb, double *x, const int *ix, const int *jx, const int *descx, double *ferr, double *berr,
              double *work, const int *lwork, int *iwork, const int *liwork, int *info)
{ gerfs_impl(trans, *n, *nrhs, a, *ia, *ja, desca, af, *iaf, *jaf, descaf, ipiv, b, *ib, *jb, descb, x, *ix, *jx, descx, ferr, berr, work, *lwork, iwork, *liwork, info); }

void pdgesvx_(const char *fact, const char *trans, const int *n, const int *nrhs, double *a, const int *ia, const int *ja,
              const int *desca, double *af, const int *iaf, const int *jaf, const int *descaf, int *ipiv, char *equed, double *r,
              double *c, double *b, const int *ib, const int *jb, const int *descb, double *x, const int *ix, const int *jx,
              const int *descx, double *rcond, double *ferr, double *berr, double *work, const int *lwork, int *iwork,
              const int *liwork, int *info)
{ gesvx_impl(fact, trans, *n, *nrhs, a, *ia, *ja, desca, af, *iaf, *jaf, descaf, ipiv, equed, r, c, b, *ib, *jb, descb, x, *ix, *jx, descx, rcond, ferr, berr, work, *lwork, iwork, *liwork, info); }

}  // extern "C"
