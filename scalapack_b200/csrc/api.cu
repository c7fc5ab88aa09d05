// api.cu -- the reference-facing entry points: PDGETRF / PDGETRS / PDGESV (+ PZ*) with the reference's
// argument checks, INFO encoding and quick returns (SRC/pdgetrf.f:169-206, pdgetrs.f:185-242,
// pdgesv.f:185-237), host<->device staging for host-resident callers, and the test-driver helpers.
#include "common.h"
#include "kernels.cuh"
#include "lu.h"
#include "ncclw.h"
#include "stage.h"
#include "entry.h"

namespace slb {
void getrs_l3_entry(Grid *g, char trans, int n, int nrhs, const double *Adev, int64_t lda, int nb, int rsrc, int csrc, const int *ipg,
                    double *b, int ib, int jb, const int *descb);   // pblas.cu
}

#include <cmath>

namespace slb {

// A host-resident 1-D / small array staged through HBM in one piece (test-driver helpers); a device-resident one is used in place.
template <typename T>
struct Staged {
    T *dev = nullptr; T *host = nullptr; size_t bytes = 0; bool staged = false;
    Staged(const char *name, T *p, size_t elems)
    {
        bytes = elems * sizeof(T);
        if (elems == 0 || is_device_ptr(p)) { dev = p; return; }
        staged = true; host = p;
        dev = (T *)workspace(name, bytes);
        SLB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, rt().s_main));
        SLB_CUDA(cudaStreamSynchronize(rt().s_main));
        counter_add("h2d_bytes", (int64_t)bytes);
    }
    void writeback()
    {
        if (!staged) return;
        SLB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, rt().s_main));
        SLB_CUDA(cudaStreamSynchronize(rt().s_main));
        counter_add("d2h_bytes", (int64_t)bytes);
    }
};

// The window of a host- or device-resident local array as the device sees it.  Host-resident: a staging copy in HBM (even
// leading dimension, so the 16-byte epilogues of the update stay aligned) reached through a HostLink; device-resident: in place.
template <typename T>
struct DevWindow {
    T *dev = nullptr; int64_t ld = 1; bool staged = false;
    HostMat hm; HostLink *link = nullptr;
    DevWindow(const char *name, T *p, int64_t lld, int64_t loff_r, int64_t loff_c, int64_t rows, int64_t cols)
    {
        T *win = p + loff_r + loff_c * lld;
        if (rows <= 0 || cols <= 0 || is_device_ptr(p)) { dev = win; ld = lld; return; }
        staged = true;
        ld = (rows + 1) & ~(int64_t)1;
        dev = (T *)workspace(name, (size_t)ld * (size_t)cols * sizeof(T));
        hm.p = win; hm.ld = lld; hm.rows = rows; hm.cols = cols; hm.elem = sizeof(T); hm.pinned = host_ptr_is_pinned(p);
        link = new HostLink(hm, dev, ld);
    }
    void upload_all() { if (link) link->wait(link->upload(0, hm.rows, 0, hm.cols)); }
    void download_all() { if (link) { link->download(0, hm.rows, 0, hm.cols, nullptr); link->finish(); } }
    ~DevWindow() { delete link; }
};

template <typename T>
static void getrf_entry(const char *name, const int *m, const int *n, T *a, const int *ia, const int *ja, const int *desca,
                        int *ipiv, int *info)
{
    const int ictxt = desca[CTXT_];
    int nprow, npcol, myrow, mycol;
    blacs_gridinfo_(&ictxt, &nprow, &npcol, &myrow, &mycol);
    *info = 0;
    if (nprow == -1) *info = -(600 + CTXT_ + 1);
    else {
        chk1mat(*m, 1, *n, 2, *ia, *ja, desca, 6, info);
        if (*info == 0) {
            int iroff = (*ia - 1) % desca[MB_], icoff = (*ja - 1) % desca[NB_];
            if (iroff != 0) *info = -4;
            else if (icoff != 0) *info = -5;
            else if (desca[MB_] != desca[NB_]) *info = -(600 + NB_ + 1);
        }
        int zero = 0, one = 1, two = 2, six = 6, idum = 0;
        pchk1mat_(m, &one, n, &two, ia, ja, desca, &six, &zero, &idum, &idum, info);
    }
    if (*info != 0) { xerbla(ictxt, name, *info); return; }
    if (desca[M_] == 1) { ipiv[0] = 1; return; }
    if (*m == 0 || *n == 0) return;
    Grid *g = grid_of(ictxt);
    const int nb = desca[NB_];
    const int64_t lld = desca[LLD_];
    const Window w = window(*m, *n, *ia, *ja, desca, nprow, npcol, myrow, mycol);
    const int mn = *m < *n ? *m : *n;
    DevWindow<T> A("stage_A", a, lld, w.loff_r, w.loff_c, w.mloc, w.nloc);
    std::vector<int> ipg((size_t)mn);
    getrf_device<T>(g, *m, *n, A.dev, A.ld, nb, w.rsrc, w.csrc, ipg.data(), info, A.link);
    if (A.link && !g_last_lu.host_written) A.download_all();
    fill_local_ipiv(ipg, mn, nb, w.rsrc, nprow, myrow, ipiv + w.loff_r, *ia - 1);
    // INFO = k > 0: U(IA+k-1, JA+k-1) is exactly zero -- k is already relative to sub(A) (pdgetrf.f:129-132)
}

template <typename T>
static void getrs_checks(const char *name, int descpos_a, int descpos_b, const char *trans, const int *n, const int *nrhs,
                         const int *ia, const int *ja, const int *desca, const int *ib, const int *jb, const int *descb,
                         int *info, bool has_trans)
{
    const int ictxt = desca[CTXT_];
    int nprow, npcol, myrow, mycol;
    blacs_gridinfo_(&ictxt, &nprow, &npcol, &myrow, &mycol);
    *info = 0;
    const int DA = descpos_a, DB = descpos_b;       // positions of DESCA / DESCB in the argument list
    if (nprow == -1) { *info = -(DA * 100 + CTXT_ + 1); return; }
    const int pn = has_trans ? 2 : 1, pnrhs = has_trans ? 3 : 2;
    chk1mat(*n, pn, *n, pn, *ia, *ja, desca, DA, info);
    chk1mat(*n, pn, *nrhs, pnrhs, *ib, *jb, descb, DB, info);
    if (*info == 0) {
        int iarow = indxg2p(*ia, desca[MB_], desca[RSRC_], nprow), ibrow = indxg2p(*ib, descb[MB_], descb[RSRC_], nprow);
        int iroffa = (*ia - 1) % desca[MB_], icoffa = (*ja - 1) % desca[NB_], iroffb = (*ib - 1) % descb[MB_];
        if (has_trans) {
            char t = trans[0] & ~0x20;
            if (t != 'N' && t != 'T' && t != 'C') *info = -1;
            else if (iroffa != 0) *info = -5;
            else if (icoffa != 0) *info = -6;
            else if (desca[MB_] != desca[NB_]) *info = -(DA * 100 + NB_ + 1);
            else if (iroffb != 0 || ibrow != iarow) *info = -10;
            else if (descb[MB_] != desca[NB_]) *info = -(DB * 100 + NB_ + 1);
            else if (ictxt != descb[CTXT_]) *info = -(DB * 100 + CTXT_ + 1);
        } else {
            if (iroffa != 0) *info = -4;
            else if (icoffa != 0) *info = -5;
            else if (desca[MB_] != desca[NB_]) *info = -(DA * 100 + NB_ + 1);
            else if (ibrow != iarow || icoffa != iroffb) *info = -9;
            else if (descb[MB_] != desca[NB_]) *info = -(DB * 100 + NB_ + 1);
            else if (ictxt != descb[CTXT_]) *info = -(DB * 100 + CTXT_ + 1);
        }
    }
    int nextra = has_trans ? 1 : 0, ex[1] = { 0 }, expos[1] = { 1 };
    if (has_trans) { char t = trans[0] & ~0x20; ex[0] = (t == 'N') ? 'N' : (t == 'T' ? 'T' : 'C'); }
    int da = DA, db = DB;
    pchk2mat_(n, &pn, n, &pn, ia, ja, desca, &da, n, &pn, nrhs, &pnrhs, ib, jb, descb, &db, &nextra, ex, expos, info);
    (void)name;
}

template <typename T>
static void getrs_entry(const char *name, const char *trans, const int *n, const int *nrhs, const T *a, const int *ia,
                        const int *ja, const int *desca, const int *ipiv, T *b, const int *ib, const int *jb, const int *descb,
                        int *info)
{
    const int ictxt = desca[CTXT_];
    getrs_checks<T>(name, 7, 12, trans, n, nrhs, ia, ja, desca, ib, jb, descb, info, true);
    if (*info != 0) { xerbla(ictxt, name, *info); return; }
    if (*n == 0 || *nrhs == 0) return;
    const char t = trans[0] & ~0x20;
    int nprow, npcol, myrow, mycol;
    blacs_gridinfo_(&ictxt, &nprow, &npcol, &myrow, &mycol);
    Grid *g = grid_of(ictxt);
    const int nb = desca[NB_];
    const Window w = window(*n, *n, *ia, *ja, desca, nprow, npcol, myrow, mycol);
    const RhsWindow wb = rhs_window(*ib, descb, nprow, npcol, myrow, mycol);
    std::vector<int> ipg;
    gather_global_ipiv(g, *n, nb, w.rsrc, ipiv + w.loff_r, *ia - 1, ipg);
    DevWindow<T> A("stage_A", const_cast<T *>(a), desca[LLD_], w.loff_r, w.loff_c, w.mloc, w.nloc);
    if constexpr (sizeof(T) == sizeof(double)) {
        // many right-hand sides (the PB_CptrsmAB case): block-cyclic copy of sub(B), level-3 sweeps on the tensor cores (pblas.cu)
        const int64_t l3min = opt("solve_l3_min_nrhs", 64);
        if (l3min > 0 && *nrhs > l3min) {
            A.upload_all();
            getrs_l3_entry(g, t == 'N' ? 'N' : 'T', *n, *nrhs, reinterpret_cast<const double *>(A.dev), A.ld, nb, w.rsrc, w.csrc, ipg.data(),
                           reinterpret_cast<double *>(b), *ib, *jb, descb);
            return;
        }
    }
    DevWindow<T> B("stage_B", b, descb[LLD_], wb.loff_r, 0, w.mloc, wb.nloc_all);
    A.upload_all(); B.upload_all();
    getrs_device<T>(g, t, *n, *nrhs, A.dev, A.ld, nb, w.rsrc, w.csrc, ipg.data(), B.dev, B.ld, descb[NB_], descb[CSRC_], *jb - 1, wb.nloc_all);
    B.download_all();
}

template <typename T>
static void gesv_entry(const char *name, const int *n, const int *nrhs, T *a, const int *ia, const int *ja, const int *desca,
                       int *ipiv, T *b, const int *ib, const int *jb, const int *descb, int *info)
{
    const int ictxt = desca[CTXT_];
    getrs_checks<T>(name, 6, 11, "N", n, nrhs, ia, ja, desca, ib, jb, descb, info, false);
    if (*info != 0) { xerbla(ictxt, name, *info); return; }
    if (*n == 0) return;
    int nprow, npcol, myrow, mycol;
    blacs_gridinfo_(&ictxt, &nprow, &npcol, &myrow, &mycol);
    Grid *g = grid_of(ictxt);
    const int nb = desca[NB_];
    const Window w = window(*n, *n, *ia, *ja, desca, nprow, npcol, myrow, mycol);
    const RhsWindow wb = rhs_window(*ib, descb, nprow, npcol, myrow, mycol);
    // one staging of A for factor + solve: the factors go back to a host-resident caller during the factorisation and stay in HBM
    DevWindow<T> A("stage_A", a, desca[LLD_], w.loff_r, w.loff_c, w.mloc, w.nloc);
    std::vector<int> ipg((size_t)*n);
    if (desca[M_] == 1) {      // 1 x 1 system: PDGETRF's quick return leaves A, IPIV(1)=1 (pdgetrf.f:201-203)
        ipiv[0] = 1; ipg[0] = 1; *info = 0;
        A.upload_all();
    } else {
        getrf_device<T>(g, *n, *n, A.dev, A.ld, nb, w.rsrc, w.csrc, ipg.data(), info, A.link);
        if (A.link && !g_last_lu.host_written) A.download_all();
        fill_local_ipiv(ipg, *n, nb, w.rsrc, nprow, myrow, ipiv + w.loff_r, *ia - 1);
    }
    if (*info == 0 && *nrhs > 0) {                       // pdgesv.f:231
        DevWindow<T> B("stage_B", b, descb[LLD_], wb.loff_r, 0, w.mloc, wb.nloc_all);
        B.upload_all();
        getrs_device<T>(g, 'N', *n, *nrhs, A.dev, A.ld, nb, w.rsrc, w.csrc, ipg.data(), B.dev, B.ld, descb[NB_], descb[CSRC_], *jb - 1, wb.nloc_all);
        B.download_all();
    }
}

}  // namespace slb

using namespace slb;

extern "C" {

void pdgetrf_(const int *m, const int *n, double *a, const int *ia, const int *ja, const int *desca, int *ipiv, int *info)
{ getrf_entry<double>("PDGETRF", m, n, a, ia, ja, desca, ipiv, info); }
void pzgetrf_(const int *m, const int *n, slb200_z *a, const int *ia, const int *ja, const int *desca, int *ipiv, int *info)
{ getrf_entry<zcomplex>("PZGETRF", m, n, reinterpret_cast<zcomplex *>(a), ia, ja, desca, ipiv, info); }

void pdgetrs_(const char *trans, const int *n, const int *nrhs, const double *a, const int *ia, const int *ja, const int *desca,
              const int *ipiv, double *b, const int *ib, const int *jb, const int *descb, int *info)
{ getrs_entry<double>("PDGETRS", trans, n, nrhs, a, ia, ja, desca, ipiv, b, ib, jb, descb, info); }
void pzgetrs_(const char *trans, const int *n, const int *nrhs, const slb200_z *a, const int *ia, const int *ja, const int *desca,
              const int *ipiv, slb200_z *b, const int *ib, const int *jb, const int *descb, int *info)
{ getrs_entry<zcomplex>("PZGETRS", trans, n, nrhs, reinterpret_cast<const zcomplex *>(a), ia, ja, desca, ipiv, reinterpret_cast<zcomplex *>(b), ib, jb, descb, info); }

void pdgesv_(const int *n, const int *nrhs, double *a, const int *ia, const int *ja, const int *desca, int *ipiv, double *b,
             const int *ib, const int *jb, const int *descb, int *info)
{ gesv_entry<double>("PDGESV", n, nrhs, a, ia, ja, desca, ipiv, b, ib, jb, descb, info); }
void pzgesv_(const int *n, const int *nrhs, slb200_z *a, const int *ia, const int *ja, const int *desca, int *ipiv, slb200_z *b,
             const int *ib, const int *jb, const int *descb, int *info)
{ gesv_entry<zcomplex>("PZGESV", n, nrhs, reinterpret_cast<zcomplex *>(a), ia, ja, desca, ipiv, reinterpret_cast<zcomplex *>(b), ib, jb, descb, info); }

double slb200_last_factor_ms(void) { return g_last_lu.factor_ms; }
double slb200_last_solve_ms(void) { return g_last_lu.solve_ms; }
double slb200_last_update_ms(void) { return g_last_lu.update_ms; }
double slb200_last_update_flops(void) { return g_last_lu.update_flops; }
int64_t slb200_last_update_launches(void) { return g_last_lu.update_launches; }

// ---- test-driver helpers ------------------------------------------------------------------------------
void slb200_pdmatgen(const int *ictxt, const int *m, const int *n, const int *mb, const int *nb, double *a, const int *lda,
                     const int *iarow, const int *iacol, const int *iseed)
{
    int P, Q, r, c; blacs_gridinfo_(ictxt, &P, &Q, &r, &c);
    if (P < 0) return;
    int64_t nloc = numroc(*n, *nb, c, *iacol, Q);
    Staged<double> A("stage_A", a, (size_t)*lda * (size_t)nloc);
    launch_pdmatgen_local(*m, *n, *mb, *nb, A.dev, *lda, *iarow, *iacol, *iseed, r, c, P, Q, rt().s_main);
    SLB_CUDA(cudaStreamSynchronize(rt().s_main));
    A.writeback();
}

static void matgen64_any(const int *ictxt, int64_t m, int64_t n, int mb, int nb, double *a, int64_t lda, int iarow, int iacol,
                         uint64_t seed, int cplx)
{
    int P, Q, r, c; blacs_gridinfo_(ictxt, &P, &Q, &r, &c);
    if (P < 0) return;
    int64_t nloc = numroc((int)n, nb, c, iacol, Q);
    size_t elems = (size_t)lda * (size_t)nloc * (cplx ? 2 : 1);
    Staged<double> A("stage_A", a, elems);
    launch_matgen64_local(m, n, mb, nb, A.dev, lda, iarow, iacol, seed, r, c, P, Q, cplx, rt().s_main);
    SLB_CUDA(cudaStreamSynchronize(rt().s_main));
    A.writeback();
}
void slb200_matgen64(const int *ictxt, const int64_t *m, const int64_t *n, const int *mb, const int *nb, double *a,
                     const int64_t *lda, const int *iarow, const int *iacol, const uint64_t *seed)
{ matgen64_any(ictxt, *m, *n, *mb, *nb, a, *lda, *iarow, *iacol, *seed, 0); }
void slb200_zmatgen64(const int *ictxt, const int64_t *m, const int64_t *n, const int *mb, const int *nb, slb200_z *a,
                      const int64_t *lda, const int *iarow, const int *iacol, const uint64_t *seed)
{ matgen64_any(ictxt, *m, *n, *mb, *nb, reinterpret_cast<double *>(a), *lda, *iarow, *iacol, *seed, 1); }

// Solve residual of TESTING/traditional/LIN/pdlaschk.f:187,296 with A and B regenerated on the device:
//   max_j ||b_j - A x_j||_inf / (||x_j||_inf ||A||_inf eps N),  eps = 2^-53 (PDLAMCH 'eps').
// Real, IA=JA=1, RSRC=CSRC=0, X distributed like B of PDGESV (row blocks nb, column blocks descx[NB_]).
double slb200_pdlaschk(const int *ictxt, const int *n_, const int *nrhs_, const double *x, const int *descx, const int *desca,
                       const uint64_t *aseed, const uint64_t *bseed, const int *gen)
{
    int P, Q, myrow, mycol; blacs_gridinfo_(ictxt, &P, &Q, &myrow, &mycol);
    if (P < 0) return -1.0;
    Grid *g = grid_of(*ictxt);
    Runtime &r = rt(); cudaStream_t s = r.s_main;
    const int n = *n_, nrhs = *nrhs_, nb = desca[NB_], nbx = descx[NB_];
    const int64_t mloc = numroc(n, nb, myrow, 0, P), nloc = numroc(n, nb, mycol, 0, Q);
    const int64_t nlocx = numroc(nrhs, nbx, mycol, descx[CSRC_], Q), lldx = descx[LLD_];
    // full X on every process (host): local pieces -> global, summed over the grid
    std::vector<double> xl((size_t)lldx * (nlocx > 0 ? nlocx : 1));
    if (nlocx > 0) SLB_CUDA(cudaMemcpy(xl.data(), x, xl.size() * sizeof(double), cudaMemcpyDefault));
    std::vector<double> xg((size_t)n * nrhs, 0.0);
    const int myc_rel = (Q + mycol - descx[CSRC_]) % Q;
    for (int64_t jl = 0; jl < nlocx; ++jl) {
        int64_t jg = ((jl / nbx) * Q + myc_rel) * nbx + jl % nbx;
        for (int64_t il = 0; il < mloc; ++il) { int64_t ig = ((il / nb) * P + myrow) * nb + il % nb; xg[ig + jg * n] = xl[il + jl * lldx]; }
    }
    const int np = P * Q;
    if (np > 1) {
        std::vector<double> all((size_t)n * nrhs * np);
        grid_allgather(g, 'A', xg.data(), all.data(), xg.size() * sizeof(double));
        for (size_t e = 0; e < xg.size(); ++e) { double v = 0; for (int p = 0; p < np; ++p) v += all[(size_t)p * xg.size() + e]; xg[e] = v; }
    }
    double resid = 0.0;
    double *xrow = (double *)workspace("chk_x", (size_t)(nloc > 0 ? nloc : 1) * sizeof(double));
    double *rr = (double *)workspace("chk_r", (size_t)2 * (mloc > 0 ? mloc : 1) * sizeof(double));
    double *ra = rr + (mloc > 0 ? mloc : 1);
    std::vector<double> xmine((size_t)(nloc > 0 ? nloc : 1)), hr((size_t)2 * (mloc > 0 ? mloc : 1));
    for (int c = 0; c < nrhs; ++c) {
        for (int64_t jl = 0; jl < nloc; ++jl) { int64_t jg = ((jl / nb) * Q + mycol) * nb + jl % nb; xmine[jl] = xg[jg + (int64_t)c * n]; }
        SLB_CUDA(cudaMemcpyAsync(xrow, xmine.data(), xmine.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        SLB_CUDA(cudaMemsetAsync(rr, 0, (size_t)2 * (mloc > 0 ? mloc : 1) * sizeof(double), s));
        launch_gen_matvec(n, nb, *aseed, *gen, myrow, mycol, P, Q, xrow, rr, ra, s);
        SLB_CUDA(cudaMemcpyAsync(hr.data(), rr, hr.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        SLB_CUDA(cudaStreamSynchronize(s));
        // sum partial A x and row sums of |A| over the process row
        std::vector<double> sum = hr;
        if (Q > 1) {
            std::vector<double> all(hr.size() * Q);
            grid_allgather(g, 'R', hr.data(), all.data(), hr.size() * sizeof(double));
            for (size_t e = 0; e < hr.size(); ++e) { double v = 0; for (int p = 0; p < Q; ++p) v += all[(size_t)p * hr.size() + e]; sum[e] = v; }
        }
        // b regenerated on the host for my rows (N values per right-hand side: cheap)
        double rmax = 0, amax = 0, xmax = 0;
        const int64_t ml = mloc > 0 ? mloc : 1;
        for (int64_t il = 0; il < mloc; ++il) {
            int64_t ig = ((il / nb) * P + myrow) * nb + il % nb;
            double bv;
            if (*gen == 31) {
                unsigned long long t = 1ULL + (unsigned long long)ig + (unsigned long long)c * (unsigned long long)n, a_ = 1103515245ULL, c_ = 12345ULL, ra_ = 1, rc_ = 0;
                while (t) { if (t & 1) { rc_ = (a_ * rc_ + c_) & 0x7fffffffULL; ra_ = (a_ * ra_) & 0x7fffffffULL; } c_ = ((a_ + 1) * c_) & 0x7fffffffULL; a_ = (a_ * a_) & 0x7fffffffULL; t >>= 1; }
                unsigned long long xs = (ra_ * *bseed + rc_) & 0x7fffffffULL;
                bv = 1.0 - 2.0 * ((double)xs / 2147483648.0);
            } else {
                unsigned long long t = 1ULL + (unsigned long long)ig + (unsigned long long)c * (unsigned long long)n, a_ = 6364136223846793005ULL, c_ = 1ULL, ra_ = 1, rc_ = 0;
                while (t) { if (t & 1) { rc_ = a_ * rc_ + c_; ra_ = a_ * ra_; } c_ = (a_ + 1) * c_; a_ = a_ * a_; t >>= 1; }
                unsigned long long xs = ra_ * *bseed + rc_;
                bv = (double)(xs >> 11) * (1.0 / 9007199254740992.0) - 0.5;
            }
            double rv = fabs(bv - sum[il]);
            if (rv > rmax || rv != rv) rmax = rv;
            if (sum[ml + il] > amax) amax = sum[ml + il];
        }
        for (int64_t i = 0; i < n; ++i) { double v = fabs(xg[i + (int64_t)c * n]); if (v > xmax || v != v) xmax = v; }
        double loc[2] = { rmax, amax };
        if (P > 1) {
            std::vector<double> all((size_t)2 * P);
            grid_allgather(g, 'C', loc, all.data(), sizeof(loc));
            for (int p = 0; p < P; ++p) { if (all[2 * p] > loc[0] || all[2 * p] != all[2 * p]) loc[0] = all[2 * p]; if (all[2 * p + 1] > loc[1]) loc[1] = all[2 * p + 1]; }
        }
        double v = loc[0] / (xmax * loc[1] * ldexp(1.0, -53) * (double)n);
        if (v > resid || v != v) resid = v;
    }
    return resid;
}

// test hook (host integer code only, no GPU): the local window of sub(A) as the entry points compute it
// out = { loff_r, loff_c, mloc, nloc, rsrc, csrc }
void slb200_test_window(int m, int n, int ia, int ja, const int *desc, int nprow, int npcol, int myrow, int mycol, int64_t *out)
{
    const Window w = window(m, n, ia, ja, desc, nprow, npcol, myrow, mycol);
    out[0] = w.loff_r; out[1] = w.loff_c; out[2] = w.mloc; out[3] = w.nloc; out[4] = w.rsrc; out[5] = w.csrc;
}

// micro-benchmarks exported for bench.py (roofline denominators)
double slb200_bench_dmma_tflops(int iters) { return bench_dmma_peak_tflops(iters); }
double slb200_bench_dfma_tflops(int iters) { return bench_dfma_peak_tflops(iters); }
double slb200_bench_copy_gbs(int64_t bytes) { return bench_copy_gbs((size_t)bytes); }

}  // extern "C"
