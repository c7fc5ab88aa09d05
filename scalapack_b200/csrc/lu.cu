// lu.cu -- blocked right-looking LU with partial pivoting on a P x Q block-cyclic grid, one GPU per process.
// Device-side orchestration of what SRC/pdgetrf.f:219-302 does with PDGETF2 / PDLASWP / PDTRSM / PDGEMM and
// the BLACS broadcasts of PBLAS/SRC/PTOOLS/PB_CInV.c:327-347,472-493 -- redesigned for NVLink-connected GPUs:
//
//  * panel (PDGETF2): the process column's panel is gathered onto the GPU that owns the diagonal block and
//    factored there by ONE cooperative kernel (panel.cu); P <= 2..8 peers on NVSwitch make the gather (one
//    message per peer) far cheaper than the reference's ~3 latency-bound messages per COLUMN.
//  * pivots + L11 + L21 travel in ONE message per process ("Pbuf"): scatter down the column, broadcast
//    along the row (the reference: IGEBS2D of pivots + Cdgebs2d of L strips + the L21 panel).
//  * row interchanges (PDLASWP left and right) + U12 formation are fused: each GPU packs the rows it owns
//    that end in the top block, the process column all-gathers them, and EVERY process row then holds the
//    whole permuted block row; each solves U12 = L11^-1 (...) redundantly, so the reference's separate U12
//    column broadcast disappears.
//  * trailing update (PDGEMM): local DMMA kernel (gemm.cu) on Lloc (contiguous) x U (contiguous).
// All message sizes are known on the host up front; pivots stay on the device (no host sync per step).
#include "common.h"
#include "kernels.cuh"
#include "ncclw.h"
#include "lu.h"
#include "stage.h"

#include <algorithm>

namespace slb {

template <typename T> struct Ops;
template <> struct Ops<double> {
    static void gemm(int64_t M, int64_t N, int K, const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc, cudaStream_t s, int chunk = 0, int flags = 0)
    { launch_dgemm_minus(M, N, K, A, lda, B, ldb, C, ldc, s, chunk, flags); }
    static bool packs(int64_t M, int K) { return dgemm_takes_packed(M, K, GEMM_MAIN); }
    static void trsm(int jb, int64_t n, const double *L, int64_t ldl, double *B, int64_t ldb, cudaStream_t s) { launch_dtrsm_llnu(jb, n, L, ldl, B, ldb, s); }
    static void panel(int m, int jb, double *W, int64_t ldw, const PanelRowMap &map, int *ipiv, int *info, int off, void *work, cudaStream_t s, int gmax = 0)
    { launch_dpanel(m, jb, W, ldw, map, ipiv, info, off, work, s, gmax); }
    static constexpr double flop_mul = 1.0;
};
template <> struct Ops<zcomplex> {
    static void gemm(int64_t M, int64_t N, int K, const zcomplex *A, int64_t lda, const zcomplex *B, int64_t ldb, zcomplex *C, int64_t ldc, cudaStream_t s, int chunk = 0, int flags = 0)
    { launch_zgemm_minus(M, N, K, A, lda, B, ldb, C, ldc, s, chunk, flags); }
    static bool packs(int64_t M, int K) { return zgemm_takes_packed(M, K, GEMM_MAIN); }
    static void trsm(int jb, int64_t n, const zcomplex *L, int64_t ldl, zcomplex *B, int64_t ldb, cudaStream_t s) { launch_ztrsm_llnu(jb, n, L, ldl, B, ldb, s); }
    static void panel(int m, int jb, zcomplex *W, int64_t ldw, const PanelRowMap &map, int *ipiv, int *info, int off, void *work, cudaStream_t s, int gmax = 0)
    { launch_zpanel(m, jb, W, ldw, map, ipiv, info, off, work, s, gmax); }
    static constexpr double flop_mul = 4.0;
};

LuStats g_last_lu;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- 1 x 1 grid: look-ahead + two-half software pipeline --------------------------------------------------------
// The reference's loop is strictly serial (SRC/pdgetrf.f:254-295).  Here four streams run one step's phases against
// the previous step's trailing update:
//   sp (highest priority)  panel k+1 (+ its interchange plan) as soon as the update of step k has done the next
//                          panel's columns;
//   sq (high)              "prep" = row interchanges (PDLASWP, right part) + U12 solve (PDTRSM) of step k, in two column
//                          halves: near = [cr, b), far = [b, N);
//   sg (low)               the DMMA update (PDGEMM) of step k: next panel's columns, rest of near, far -- with CHUNKED CTAs
//                          that retire every few tiles so the kernels of sp / sq find SMs;
//   sc (high)              interchanges of the already factored columns [0, j0) (nothing reads them again).
// prep_k(near) runs under update_{k-1}(far), prep_k(far) under update_k(near): the HBM-latency-bound interchanges and the
// launch-latency-bound U12 solve leave the critical path, which becomes  update_k(near) + update_k(far)  per step.
// The boundary b only moves when near has shrunk below a quarter of the trailing columns (then one step waits for both halves).
//
// HOST-RESIDENT CALLER (link != nullptr; the drop-in case of an unmodified Fortran program): A arrives over
// PCIe in column slabs while the sweep is already running, and leaves in block rows while it is still running:
//   * upload: the sweep works on the columns [0, Np) that are PRESENT.  A slab that has arrived JOINS at the top of a step k:
//     it is first taken through steps 0 .. k-1 ("replay": interchanges, U12 solve and update of exactly these columns with
//     the kept plan and a kept copy of each panel in its step-time row order), then it is part of the trailing matrix.  Every
//     element sees the same operations in the same order as in the device-resident sweep, so the factors are bit-identical.
//     Which slabs join when follows their ACTUAL arrival (the host stays one step ahead of the device while slabs are
//     pending), except that the columns of the next panel are waited for.
//   * download: rows [0, j0 + jb) never change after step k (later interchanges only touch rows below), so block row k goes
//     back to the caller as soon as its interchanges and U12 solve are done.
template <typename T>
static int getrf_lookahead_1x1(int M, int N, T *A, int64_t lld, int nb, int *ipiv_glob_host, int *info_host, HostLink *link)
{
    Runtime &r = rt();
    cudaStream_t sg = r.s_main, sp = r.s_panel, sc = r.s_copy, sq = r.s_prep, sx = r.s_aux;
    const int mn = M < N ? M : N;
    const int nsteps = (mn + nb - 1) / nb;
    int *ipiv_dev = (int *)workspace("lu_ipiv", (size_t)(mn + nb + 16) * sizeof(int));
    int *info_dev = (int *)workspace("lu_info", 64);
    int *plan_mem = (int *)workspace("lu_plan", (size_t)3 * nb * (nsteps + 1) * sizeof(int));     // one plan per step (the replay re-reads them)
    auto plan_of = [&](int k) { int *p = plan_mem + (size_t)3 * nb * k; return SwapPlan{ p, p + nb, p + 2 * nb }; };
    void *panel_work = workspace("lu_panelwork", panel_work_bytes(nb), true);
    T *Ubuf = (T *)workspace("lu_U", (size_t)nb * N * sizeof(T));
    T *Obuf = (T *)workspace("lu_O", (size_t)nb * N * sizeof(T));
    const int gmax_opt = (int)opt("panel_gmax", 32), chunk_opt = (int)opt("gemm_chunk", 4);
    const double overlap_min_ms = (double)opt("lookahead_min_us", 4000) * 1e-3;
    const bool pipe = opt("la_pipeline", 1) != 0;
    // fewer trailing columns: the step is not split.  The near half must hold the next panel and the re-split test below works in
    // quarters of the trailing width, so the option is taken no lower than 2 NB and 4 (found by scripts/fuzz_lu_path.py: with 1-3
    // trailing columns split, the near half ran empty and the next panel was factored before its update)
    const int64_t split_min = std::max<int64_t>(opt("la_split_min", 6144), std::max<int64_t>(2 * (int64_t)nb, 4));
    const bool trace = opt("la_trace", 0) != 0;               // per-step timeline on stderr

    auto mkev = [](std::vector<cudaEvent_t> &v, size_t n) { v.resize(n); for (auto &e : v) SLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); };
    cudaEvent_t ev0, ev1, evs, rj;
    SLB_CUDA(cudaEventCreate(&ev0)); SLB_CUDA(cudaEventCreate(&ev1)); SLB_CUDA(cudaEventCreateWithFlags(&evs, cudaEventDisableTiming));
    SLB_CUDA(cudaEventCreateWithFlags(&rj, cudaEventDisableTiming));
    std::vector<cudaEvent_t> evp, evn, evl, pdp, pdn, pdf, gdn, gdf, dlev;
    mkev(evp, (size_t)nsteps + 1); mkev(evn, (size_t)nsteps + 1); mkev(evl, (size_t)nsteps + 1);
    mkev(pdp, (size_t)nsteps); mkev(pdn, (size_t)nsteps); mkev(pdf, (size_t)nsteps); mkev(gdn, (size_t)nsteps); mkev(gdf, (size_t)nsteps);
    std::vector<cudaEvent_t> gev((size_t)6 * nsteps, nullptr), eva((size_t)nsteps, nullptr);
    std::vector<char> evn_set((size_t)nsteps + 1, 0);
    std::vector<double> gflops((size_t)nsteps, 0.0);

    // ---- host-resident caller: queue the upload of every column slab now (narrow slabs first: the first panels start early) ----
    struct Slab { int64_t c0, c1; int ticket; };
    std::vector<Slab> slabs;
    size_t next_slab = 0;
    int64_t Np = N;                                            // columns [0, Np) take part in the sweep
    int ksave = 0; T *Lsave = nullptr;
    bool dl_started = false, rj_set = false;
    if (link) {
        const int64_t wmax = std::max<int64_t>(nb, (int64_t)((size_t)opt("e2e_slab_mb", 1024) << 20) / ((int64_t)M * (int64_t)sizeof(T)) / nb * nb);
        int64_t c = 0, w = nb; int i = 0;
        while (c < N) {
            const int64_t c1 = std::min<int64_t>(N, c + w);
            slabs.push_back({ c, c1, link->upload(0, M, c, c1) });
            c = c1;
            if (i++ >= 1 && w < wmax) w = std::min(wmax, 2 * w);
        }
        // the panels of the first ksave steps are kept (step-time row order) for the replay of late slabs
        const size_t panel_bytes = (size_t)M * nb * sizeof(T);
        size_t free_b = 0, total_b = 0; SLB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        size_t budget = std::min<size_t>((size_t)opt("e2e_save_mb", 16384) << 20, free_b > ((size_t)4 << 30) ? free_b - ((size_t)4 << 30) : 0);
        ksave = (int)std::min<size_t>((size_t)nsteps, budget / panel_bytes);
        if (slabs.size() <= 1) ksave = 0;
        if (ksave > 0) Lsave = (T *)workspace("lu_Lsave", (size_t)ksave * panel_bytes);
        Np = 0;
        counter_add("e2e_upload_overlapped", 1); counter_add("e2e_download_overlapped", 1);
    }
    auto pending = [&]() { return link != nullptr && next_slab < slabs.size(); };

    SLB_CUDA(cudaEventRecord(ev0, sg));
    SLB_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), sg));
    SLB_CUDA(cudaMemsetAsync(ipiv_dev, 0, (size_t)(mn + nb) * sizeof(int), sg));
    SLB_CUDA(cudaEventRecord(evs, sg));
    SLB_CUDA(cudaStreamWaitEvent(sp, evs, 0)); SLB_CUDA(cudaStreamWaitEvent(sq, evs, 0)); SLB_CUDA(cudaStreamWaitEvent(sc, evs, 0));
    RowDist rd{ nb, 1, 0, 0, 0 };
    // panel k and its interchange plan (+ the kept copy while slabs are pending)
    auto run_panel = [&](int k, int gmax) {
        const int j0 = k * nb, jb = (mn - j0) < nb ? (mn - j0) : nb;
        PanelRowMap map{ j0, nb, 1, 0 };
        T *Wp = A + j0 + (int64_t)j0 * lld;
        Ops<T>::panel(M - j0, jb, Wp, lld, map, ipiv_dev + j0, info_dev, j0, panel_work, sp, gmax);
        if (pending() && k < ksave) launch_copy2d<T>(M - j0, jb, Wp, lld, Lsave + (size_t)k * M * nb, M - j0, sp);
        launch_swap_plan(j0, jb, ipiv_dev + j0, plan_of(k), sp);
        SLB_CUDA(cudaEventRecord(evp[k], sp));
    };
    // columns [a, b) through steps 0 .. k-1 on the update stream (the interchange kernels run alone here: uncapped grids)
    auto replay = [&](int64_t a, int64_t b, int k) {
        swap_grid_override(0);
        for (int kk = 0; kk < k; ++kk) {
            const int j0 = kk * nb, jb = nb;
            const int64_t cr = j0 + jb, ldl = M - j0;
            const T *Ls = Lsave + (size_t)kk * M * nb;
            SwapPlan plan = plan_of(kk);
            T *Ur = Ubuf + a * jb, *Or = Obuf + a * jb;
            launch_swap_pack<T>(jb, j0, plan, rd, A, lld, a, b, Ur, jb, Or, jb, sg);
            launch_swap_unpack_out<T>(jb, plan, rd, A, lld, a, b, Or, jb, sg);
            Ops<T>::trsm(jb, b - a, Ls, ldl, Ur, jb, sg);
            launch_copy2d<T>(jb, b - a, Ur, jb, A + j0 + a * lld, lld, sg);
            if (M > cr) Ops<T>::gemm(M - cr, b - a, jb, Ls + jb, ldl, Ur, jb, A + cr + a * lld, lld, sg, 0, GEMM_MAIN);
        }
        swap_grid_override(-1);
    };
    // one event that completes when all of `evs` have (the aux stream carries no work)
    auto all_of = [&](std::initializer_list<cudaEvent_t> list) {
        for (cudaEvent_t e : list) if (e) SLB_CUDA(cudaStreamWaitEvent(sx, e, 0));
        cudaEvent_t o; SLB_CUDA(cudaEventCreateWithFlags(&o, cudaEventDisableTiming));
        SLB_CUDA(cudaEventRecord(o, sx));
        dlev.push_back(o);
        return o;
    };
    if (link) { link->stream_wait(slabs[0].ticket, sp); next_slab = 1; Np = slabs[0].c1; }
    run_panel(0, 0);
    int64_t b = -1;                 // absolute column of the near | far boundary of the previous step (-1: none yet)
    bool far_prev = false;          // the previous step had a far half
    cudaEvent_t last_gemm = nullptr;   // completion of the previous step's update (all of it)
    for (int k = 0; k < nsteps; ++k) {
        const int j0 = k * nb, jb = (mn - j0) < nb ? (mn - j0) : nb;
        T *Wp = A + j0 + (int64_t)j0 * lld;
        SwapPlan plan = plan_of(k);
        const int64_t cr = j0 + jb;
        // ---- slabs that have arrived (or must be waited for) join the sweep ----
        bool joined = false;
        if (pending()) {
            if (k >= 1 && evn_set[k - 1]) SLB_CUDA(cudaEventSynchronize(evn[k - 1]));       // stay one step ahead of the device, not more
            const int64_t need = std::min<int64_t>(N, cr + nb);
            const bool force = k >= ksave || k == nsteps - 1;
            const size_t first = next_slab;
            while (next_slab < slabs.size() && (force || slabs[next_slab].c0 < need || link->done(slabs[next_slab].ticket))) {
                link->stream_wait(slabs[next_slab].ticket, sg);
                ++next_slab;
            }
            if (next_slab > first) {
                replay(slabs[first].c0, slabs[next_slab - 1].c1, k);
                Np = slabs[next_slab - 1].c1;
                SLB_CUDA(cudaEventRecord(rj, sg)); rj_set = true;
                joined = true;
            }
        }
        const int64_t nright = Np - cr;
        const int64_t mrows = M - cr;
        const bool have_next = k + 1 < nsteps && mrows > 0 && nright > 0;
        const int jbn = have_next ? ((mn - (int)cr) < nb ? (mn - (int)cr) : nb) : 0;
        // ---- sc: interchanges on the already factored columns [0, j0).  Columns [j0 - nb, j0) are the L21 operand of
        // the previous step's update: wait for all of it; and stay out of the way of this step's g0 (evn[k]).
        auto left_swaps = [&]() {
            SLB_CUDA(cudaStreamWaitEvent(sc, evp[k], 0));
            if (last_gemm) SLB_CUDA(cudaStreamWaitEvent(sc, last_gemm, 0));
            if (have_next) SLB_CUDA(cudaStreamWaitEvent(sc, evn[k], 0));
            launch_swap_pack<T>(jb, j0, plan, rd, A, lld, 0, j0, Ubuf, jb, Obuf, jb, sc);
            launch_swap_unpack_out<T>(jb, plan, rd, A, lld, 0, j0, Obuf, jb, sc);
            launch_copy2d<T>(jb, j0, Ubuf, jb, A + j0, lld, sc);
            SLB_CUDA(cudaEventRecord(evl[k], sc));
        };
        // block row k (and, the first time, every block row above it) is final: back to the host-resident caller
        auto block_row_home = [&](cudaEvent_t e1, cudaEvent_t e2) {
            if (!link || pending()) return;
            link->download(dl_started ? j0 : 0, j0 + jb, 0, N, all_of({ e1, e2, evl[k], rj_set ? rj : nullptr }));
            dl_started = true;
        };
        if (nright <= 0) {
            left_swaps();
            block_row_home(evp[k], nullptr);
            continue;
        }
        // ---- near | far boundary ----
        int64_t bk = Np;
        bool resplit = false;
        if (pipe && have_next && nright >= split_min) {
            if (b < 0 || b - cr < nright / 4 || b >= Np) { bk = cr + ((nright / 2 + nb - 1) / nb) * nb; resplit = true; }
            else bk = b;
            if (bk >= Np) bk = Np;
        }
        if (bk > b || b < 0) resplit = true;          // near grows into the previous far half (or there was no split)
        auto prep = [&](int64_t c_lo, int64_t c_hi) {
            launch_swap_pack<T>(jb, j0, plan, rd, A, lld, c_lo, c_hi, Ubuf + c_lo * jb, jb, Obuf + c_lo * jb, jb, sq);
            launch_swap_unpack_out<T>(jb, plan, rd, A, lld, c_lo, c_hi, Obuf + c_lo * jb, jb, sq);
            Ops<T>::trsm(jb, c_hi - c_lo, Wp, lld, Ubuf + c_lo * jb, jb, sq);
            launch_copy2d<T>(jb, c_hi - c_lo, Ubuf + c_lo * jb, jb, A + j0 + c_lo * lld, lld, sq);
        };
        const T *Lop = Wp + jb;
        bool a_packed = false;                                           // L21 of this step already in packed form
        auto timed_gemm = [&](int slot, int64_t c_lo, int64_t c_hi, bool may_chunk) {
            const int64_t nn = c_hi - c_lo;
            if (mrows <= 0 || nn <= 0) return;
            const double est_ms = 2.0 * (double)mrows * (double)nn * jb * Ops<T>::flop_mul / 30e12 * 1e3;
            const int chunk = (may_chunk && est_ms >= overlap_min_ms * 0.5) ? chunk_opt : 0;
            SLB_CUDA(cudaEventCreate(&gev[6 * k + slot])); SLB_CUDA(cudaEventCreate(&gev[6 * k + slot + 1]));
            SLB_CUDA(cudaEventRecord(gev[6 * k + slot], sg));
            Ops<T>::gemm(mrows, nn, jb, Lop, lld, Ubuf + c_lo * jb, jb, A + cr + c_lo * lld, lld, sg, chunk, GEMM_MAIN | (a_packed ? GEMM_REUSE_A : 0));
            a_packed = Ops<T>::packs(mrows, jb);
            SLB_CUDA(cudaEventRecord(gev[6 * k + slot + 1], sg));
            gflops[k] += 2.0 * (double)mrows * (double)nn * jb * Ops<T>::flop_mul;
        };
        // ---- (a) sq: prep of the near half, the next panel's columns first (shortest path to panel k+1) ----
        SLB_CUDA(cudaStreamWaitEvent(sq, evp[k], 0));
        if (k > 0) SLB_CUDA(cudaStreamWaitEvent(sq, gdn[k - 1], 0));
        if (k > 0 && far_prev && resplit) SLB_CUDA(cudaStreamWaitEvent(sq, gdf[k - 1], 0));
        if (joined) SLB_CUDA(cudaStreamWaitEvent(sq, rj, 0));           // the new columns are in HBM and caught up
        if (jbn > 0 && cr + jbn < bk) {
            prep(cr, cr + jbn);
            SLB_CUDA(cudaEventRecord(pdp[k], sq));
            prep(cr + jbn, bk);
        } else {
            prep(cr, bk);
            SLB_CUDA(cudaEventRecord(pdp[k], sq));
        }
        SLB_CUDA(cudaEventRecord(pdn[k], sq));
        // ---- (b) sg: next panel's columns, then hand them to the panel stream ----
        SLB_CUDA(cudaStreamWaitEvent(sg, pdp[k], 0));
        if (trace) { SLB_CUDA(cudaEventCreate(&eva[k])); SLB_CUDA(cudaEventRecord(eva[k], sg)); }
        if (have_next) {
            const double rest_ms = 2.0 * (double)mrows * (double)(nright - jbn) * jb * Ops<T>::flop_mul / 30e12 * 1e3;
            const bool overlap = nright - jbn > 0 && rest_ms >= overlap_min_ms;
            timed_gemm(0, cr, cr + jbn, false);
            SLB_CUDA(cudaEventRecord(evn[k], sg)); evn_set[k] = 1;
            SLB_CUDA(cudaStreamWaitEvent(sp, evn[k], 0));
            run_panel(k + 1, overlap ? gmax_opt : 0);
        }
        // ---- (c) sc: left interchanges; (d) sq: prep of the far half -- both only after g0 (they would take its SMs) ----
        left_swaps();
        if (bk < Np) {
            if (k > 0 && far_prev) SLB_CUDA(cudaStreamWaitEvent(sq, gdf[k - 1], 0));
            if (have_next) SLB_CUDA(cudaStreamWaitEvent(sq, evn[k], 0));
            prep(bk, Np);
            SLB_CUDA(cudaEventRecord(pdf[k], sq));
        }
        block_row_home(pdn[k], bk < Np ? pdf[k] : nullptr);
        // ---- (e) sg: rest of near, far ----
        SLB_CUDA(cudaStreamWaitEvent(sg, pdn[k], 0));
        timed_gemm(2, cr + jbn, bk, have_next);
        SLB_CUDA(cudaEventRecord(gdn[k], sg));
        last_gemm = gdn[k];
        if (bk < Np) {
            SLB_CUDA(cudaStreamWaitEvent(sg, pdf[k], 0));
            timed_gemm(4, bk, Np, true);
            SLB_CUDA(cudaEventRecord(gdf[k], sg));
            last_gemm = gdf[k];
        }
        far_prev = bk < Np;
        b = bk;
    }
    SLB_CUDA(cudaStreamWaitEvent(sg, evp[nsteps - 1], 0));
    SLB_CUDA(cudaStreamWaitEvent(sg, evl[nsteps - 1], 0));
    SLB_CUDA(cudaEventRecord(evs, sq));
    SLB_CUDA(cudaStreamWaitEvent(sg, evs, 0));
    SLB_CUDA(cudaEventRecord(ev1, sg));
    SLB_CUDA(cudaMemcpyAsync(ipiv_glob_host, ipiv_dev, (size_t)mn * sizeof(int), cudaMemcpyDeviceToHost, sg));
    int info_local = 0;
    SLB_CUDA(cudaMemcpyAsync(&info_local, info_dev, sizeof(int), cudaMemcpyDeviceToHost, sg));
    g_last_lu.host_written = false;
    if (link) {                                                        // rows below the last block row (M > N): final only now
        if (M > mn) link->download(mn, M, 0, N, ev1);
        link->finish();
        g_last_lu.host_written = true;
    }
    SLB_CUDA(cudaStreamSynchronize(sg));
    SLB_CUDA(cudaStreamSynchronize(sp)); SLB_CUDA(cudaStreamSynchronize(sq)); SLB_CUDA(cudaStreamSynchronize(sc));
    float ms = 0; SLB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    g_last_lu.factor_ms = ms;
    g_last_lu.update_ms = 0; g_last_lu.update_flops = 0; g_last_lu.update_launches = 0;
    if (trace) {
        // step k: [gap = update stream idle: waits for panel k + prep of the near half] [g0 = next panel's columns] [gn = near] [gf = far]
        double t_gap = 0, t_g0 = 0, t_gn = 0, t_gf = 0;
        cudaEvent_t prev_end = ev0;
        fprintf(stderr, "la_trace: k m gap_ms g0_ms gnear_ms gfar_ms\n");
        for (int k = 0; k < nsteps; ++k) {
            if (!eva[k]) continue;
            float gap = 0, g[3] = { 0, 0, 0 };
            SLB_CUDA(cudaEventElapsedTime(&gap, prev_end, eva[k]));
            cudaEvent_t last = eva[k];
            for (int q = 0; q < 3; ++q)
                if (gev[6 * k + 2 * q]) { SLB_CUDA(cudaEventElapsedTime(&g[q], gev[6 * k + 2 * q], gev[6 * k + 2 * q + 1])); last = gev[6 * k + 2 * q + 1]; }
            prev_end = last;
            t_gap += gap; t_g0 += g[0]; t_gn += g[1]; t_gf += g[2];
            if (k % 8 == 0 || k >= nsteps - 8 || (link && k < 24)) fprintf(stderr, "la_trace: %d %d %.3f %.3f %.3f %.3f\n", k, M - k * nb, gap, g[0], g[1], g[2]);
        }
        fprintf(stderr, "la_trace: total %.1f ms = gap %.1f + g0 %.1f + gnear %.1f + gfar %.1f (+ waits inside the update stream)\n", ms, t_gap, t_g0, t_gn, t_gf);
        for (auto &e : eva) if (e) cudaEventDestroy(e);
    }
    for (int k = 0; k < nsteps; ++k) {
        for (int slot = 0; slot < 6; slot += 2)
            if (gev[6 * k + slot]) {
                float t = 0; SLB_CUDA(cudaEventElapsedTime(&t, gev[6 * k + slot], gev[6 * k + slot + 1]));
                g_last_lu.update_ms += t; g_last_lu.update_launches += 1;
                cudaEventDestroy(gev[6 * k + slot]); cudaEventDestroy(gev[6 * k + slot + 1]);
            }
        g_last_lu.update_flops += gflops[k];
    }
    for (auto *v : { &evp, &evn, &evl, &pdp, &pdn, &pdf, &gdn, &gdf, &dlev }) for (auto &e : *v) cudaEventDestroy(e);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(evs); cudaEventDestroy(rj);
    *info_host = info_local;
    return 0;
}

template <typename T>
int getrf_device(Grid *g, int M, int N, T *A, int64_t lld, int nb, int rsrc, int csrc, int *ipiv_glob_host, int *info_host, HostLink *link)
{
    g_last_lu.host_written = false;
    const bool prof = opt("profile", 0) != 0;
    Runtime &r = rt();
    const int P = g->nprow, Q = g->npcol, myrow = g->myrow, mycol = g->mycol;
    const int mn = M < N ? M : N;
    const int64_t mloc = numroc(M, nb, myrow, rsrc, P), nloc = numroc(N, nb, mycol, csrc, Q);
    // small host-resident problems are not worth the slab / block-row machinery: one copy in, one copy out
    const bool stream_io = link != nullptr && opt("e2e_overlap", 1) != 0 &&
                           (size_t)mloc * (size_t)nloc * sizeof(T) >= ((size_t)opt("e2e_overlap_min_mb", 256) << 20);
    if (link && !(stream_io && P * Q == 1 && opt("lookahead", 1) != 0 && !prof)) link->wait(link->upload(0, mloc, 0, nloc));
    if (P * Q == 1 && opt("lookahead", 1) != 0 && !prof) {
        int rc = getrf_lookahead_1x1<T>(M, N, A, lld, nb, ipiv_glob_host, info_host, stream_io ? link : nullptr);
        if (link && !stream_io) { link->download(0, mloc, 0, nloc, nullptr); link->finish(); g_last_lu.host_written = true; }
        return rc;
    }
    const bool multi = P * Q > 1;
    if (multi && !g->nccl) g->nccl = nccl_create(g);
    NcclComms *nc = g->nccl;
    // look-ahead on P x Q grids: the panel phase of step k+1 (gather, factor, scatter, row broadcast) runs on the
    // high-priority panel stream with its own column communicator while the update stream finishes step k
    const bool la = multi && opt("lookahead", 1) != 0 && opt("lookahead_multi", 1) != 0 && !prof;
    cudaStream_t sm = r.s_main, sp = la ? r.s_panel : r.s_main;
    const int gmax_opt = (int)opt("panel_gmax", 32), chunk_opt = (int)opt("gemm_chunk", 4);
    const double overlap_min_ms = (double)opt("lookahead_min_us", 4000) * 1e-3;

    // ---- workspaces ----
    int *ipiv_dev = (int *)workspace("lu_ipiv", (size_t)(mn + nb + 16) * sizeof(int));
    int *info_dev = (int *)workspace("lu_info", 64);
    int *plan_mem = (int *)workspace("lu_plan", (size_t)6 * nb * sizeof(int));
    SwapPlan plans[2] = { { plan_mem, plan_mem + nb, plan_mem + 2 * nb }, { plan_mem + 3 * nb, plan_mem + 4 * nb, plan_mem + 5 * nb } };
    void *panel_work = workspace("lu_panelwork", panel_work_bytes(nb), true);
    T *Ubuf = (T *)workspace("lu_U", (size_t)nb * (nloc > 0 ? nloc : 1) * sizeof(T));
    T *Obuf = (T *)workspace("lu_O", (size_t)nb * (nloc > 0 ? nloc : 1) * sizeof(T));
    const size_t hdr_bytes = align_up((size_t)nb * nb * sizeof(T), 256) + align_up((size_t)nb * sizeof(int), 256);
    const size_t ipiv_off = align_up((size_t)nb * nb * sizeof(T), 256);
    const size_t pbuf_bytes = align_up(hdr_bytes + (size_t)(mloc + nb) * nb * sizeof(T), 256);
    unsigned char *Pb[2] = { nullptr, nullptr }, *Psend = nullptr;
    T *Wbuf = nullptr, *Stage = nullptr, *Cmine = nullptr, *Call = nullptr;
    if (multi) {
        Pb[0] = (unsigned char *)workspace("lu_Pbuf", 2 * pbuf_bytes);
        Pb[1] = la ? Pb[0] + pbuf_bytes : Pb[0];
        if (P > 1) {
            Wbuf = (T *)workspace("lu_W", (size_t)(M + nb) * nb * sizeof(T));
            Stage = (T *)workspace("lu_stage", (size_t)(M + nb) * nb * sizeof(T));
            Psend = (unsigned char *)workspace("lu_Psend", (size_t)P * hdr_bytes + (size_t)(M + nb) * nb * sizeof(T));
            Cmine = (T *)workspace("lu_Cmine", (size_t)nb * (nloc > 0 ? nloc : 1) * sizeof(T));
            Call = (T *)workspace("lu_Call", (size_t)P * nb * (nloc > 0 ? nloc : 1) * sizeof(T));
        }
    }
    ncclComm_t_ colp = multi ? (la ? nc->colp : nc->col) : nullptr;     // column communicator of the panel phase

    cudaEvent_t ev0, ev1, evs;
    SLB_CUDA(cudaEventCreate(&ev0)); SLB_CUDA(cudaEventCreate(&ev1)); SLB_CUDA(cudaEventCreateWithFlags(&evs, cudaEventDisableTiming));
    const int nsteps = (mn + nb - 1) / nb;
    std::vector<cudaEvent_t> gev((size_t)4 * nsteps, nullptr), evp((size_t)nsteps + 1), evn((size_t)nsteps + 1);
    for (auto &e : evp) SLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : evn) SLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    std::vector<double> gflops((size_t)nsteps, 0.0);
    SLB_CUDA(cudaEventRecord(ev0, sm));
    SLB_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), sm));
    SLB_CUDA(cudaMemsetAsync(ipiv_dev, 0, (size_t)(mn + nb) * sizeof(int), sm));
    SLB_CUDA(cudaEventRecord(evs, sm));
    if (sp != sm) SLB_CUDA(cudaStreamWaitEvent(sp, evs, 0));

    // optional per-phase profile (SLB200_PROFILE=1, serial schedule only): events, no extra synchronisation
    std::vector<cudaEvent_t> pev;
    auto mark = [&]() { if (prof) { cudaEvent_t e; SLB_CUDA(cudaEventCreate(&e)); SLB_CUDA(cudaEventRecord(e, sm)); pev.push_back(e); } };
    RowDist rd{ nb, P, myrow, rsrc, 0 };
    auto rows_before = [&](int prow, int gidx) { return (int64_t)numroc(gidx, nb, prow, rsrc, P); };
    auto mloc_of = [&](int prow) { return (int64_t)numroc(M, nb, prow, rsrc, P); };

    // per-step timeline (SLB200_LA_TRACE=1): timed events on the panel stream (start | panel gathered + factored + scattered down
    // the column | row broadcast done) and on the prep stream (near start | near done | far + left done); printed by every rank
    const bool trace = opt("la_trace", 0) != 0;
    std::vector<cudaEvent_t> tev;                                 // [nsteps][8]
    if (trace) { tev.assign((size_t)8 * ((M < N ? M : N) / nb + 2), nullptr); }
    auto tmark = [&](int k, int slot, cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e; SLB_CUDA(cudaEventCreate(&e)); SLB_CUDA(cudaEventRecord(e, st)); tev[(size_t)8 * k + slot] = e;
    };
    // =============== panel phase of step k (stream sp) ===============
    auto panel_phase = [&](int k, int gmax) {
        cudaStream_t s = sp;
        tmark(k, 0, s);
        const int j0 = k * nb;
        const int jb = (mn - j0) < nb ? (mn - j0) : nb;
        const int pr = (rsrc + k) % P, pc = (csrc + k) % Q;
        const int64_t lr0 = rows_before(myrow, j0);
        const int64_t mtr = mloc - lr0;
        const int64_t lcl = numroc(j0, nb, mycol, csrc, Q);
        const int m = M - j0;
        if (!multi) {
            PanelRowMap map{ j0, nb, 1, 0 };
            Ops<T>::panel(m, jb, A + lr0 + lcl * lld, lld, map, ipiv_dev + j0, info_dev, j0, panel_work, s, gmax);
        } else {
            unsigned char *Pbuf = Pb[k & 1];
            T *pL11 = (T *)Pbuf; int *pIpiv = (int *)(Pbuf + ipiv_off); T *pLloc = (T *)(Pbuf + hdr_bytes);
            const size_t my_pbytes = hdr_bytes + (size_t)mtr * jb * sizeof(T);
            if (mycol == pc) {
                if (P == 1) {
                    PanelRowMap map{ j0, nb, 1, rsrc };
                    T *Wp = A + lr0 + lcl * lld;
                    Ops<T>::panel(m, jb, Wp, lld, map, ipiv_dev + j0, info_dev, j0, panel_work, s, gmax);
                    launch_copy2d<T>(jb, jb, Wp, lld, pL11, jb, s);
                    SLB_CUDA(cudaMemcpyAsync(pIpiv, ipiv_dev + j0, (size_t)jb * sizeof(int), cudaMemcpyDeviceToDevice, s));
                    launch_copy2d<T>(mtr, jb, Wp, lld, pLloc, mtr, s);
                } else {
                    // ---- gather the column's panel on the diagonal owner, rows in GLOBAL order (row v <-> global j0+v) ----
                    PanelRowMap map{ j0, nb, P, rsrc };
                    const int64_t mtot = m;
                    std::vector<int64_t> prows((size_t)P), plr0((size_t)P);
                    for (int prow = 0; prow < P; ++prow) { plr0[prow] = rows_before(prow, j0); prows[prow] = mloc_of(prow) - plr0[prow]; }
                    auto rel = [&](int prow) { return (prow - rsrc + P) % P; };
                    if (myrow != pr) {
                        launch_copy2d<T>(mtr, jb, A + lr0 + lcl * lld, lld, Stage, mtr, s);
                        if (mtr > 0) nccl_send(colp, Stage, (size_t)mtr * jb * sizeof(T), NT_U8, pr, s);
                        // ---- receive my factored rows + L11 + pivots ----
                        nccl_recv(colp, Pbuf, my_pbytes, NT_U8, pr, s);
                        launch_copy2d<T>(mtr, jb, pLloc, mtr, A + lr0 + lcl * lld, lld, s);
                    } else {
                        launch_rows_bc<T>(mtr, jb, A + lr0 + lcl * lld, lld, lr0, Wbuf, mtot, j0, nb, P, rel(myrow), 1, s);
                        nccl_group_start();
                        { int64_t off = 0;
                          for (int prow = 0; prow < P; ++prow) {
                              if (prow == pr) continue;
                              if (prows[prow] > 0) nccl_recv(colp, Stage + off, (size_t)prows[prow] * jb * sizeof(T), NT_U8, prow, s);
                              off += prows[prow] * jb;
                          } }
                        nccl_group_end();
                        { int64_t off = 0;
                          for (int prow = 0; prow < P; ++prow) {
                              if (prow == pr) continue;
                              launch_rows_bc<T>(prows[prow], jb, Stage + off, prows[prow], plr0[prow], Wbuf, mtot, j0, nb, P, rel(prow), 1, s);
                              off += prows[prow] * jb;
                          } }
                        Ops<T>::panel((int)mtot, jb, Wbuf, mtot, map, ipiv_dev + j0, info_dev, j0, panel_work, s, gmax);
                        // own copy + own Pbuf
                        launch_rows_bc<T>(mtr, jb, A + lr0 + lcl * lld, lld, lr0, Wbuf, mtot, j0, nb, P, rel(myrow), 0, s);
                        launch_copy2d<T>(jb, jb, Wbuf, mtot, pL11, jb, s);
                        SLB_CUDA(cudaMemcpyAsync(pIpiv, ipiv_dev + j0, (size_t)jb * sizeof(int), cudaMemcpyDeviceToDevice, s));
                        launch_rows_bc<T>(mtr, jb, pLloc, mtr, lr0, Wbuf, mtot, j0, nb, P, rel(myrow), 0, s);
                        // peers' Pbufs
                        size_t soff = 0; std::vector<size_t> soffs((size_t)P, 0);
                        for (int prow = 0; prow < P; ++prow) {
                            if (prow == pr) continue;
                            soffs[prow] = soff;
                            unsigned char *pb = Psend + soff;
                            SLB_CUDA(cudaMemcpyAsync(pb, Pbuf, hdr_bytes, cudaMemcpyDeviceToDevice, s));
                            launch_rows_bc<T>(prows[prow], jb, (T *)(pb + hdr_bytes), prows[prow], plr0[prow], Wbuf, mtot, j0, nb, P, rel(prow), 0, s);
                            soff += align_up(hdr_bytes + (size_t)prows[prow] * jb * sizeof(T), 256);
                        }
                        nccl_group_start();
                        for (int prow = 0; prow < P; ++prow) {
                            if (prow == pr) continue;
                            nccl_send(colp, Psend + soffs[prow], hdr_bytes + (size_t)prows[prow] * jb * sizeof(T), NT_U8, prow, s);
                        }
                        nccl_group_end();
                    }
                }
            }
            tmark(k, 1, s);
            if (Q > 1) nccl_bcast(nc->row, Pbuf, my_pbytes, NT_U8, pc, s);
            tmark(k, 2, s);
            if (!(mycol == pc && myrow == pr))
                SLB_CUDA(cudaMemcpyAsync(ipiv_dev + j0, pIpiv, (size_t)jb * sizeof(int), cudaMemcpyDeviceToDevice, s));
        }
        // the block's net row permutation, ready before the interchanges need it (plan buffer k&1: last read by step k-2)
        launch_swap_plan(j0, jb, ipiv_dev + j0, plans[k & 1], s);
        SLB_CUDA(cudaEventRecord(evp[k], s));
    };

    panel_phase(0, 0);
    const bool pipe = la && opt("la_pipeline", 1) != 0;
    std::vector<cudaEvent_t> dlev;
    if (pipe) {
        // ===== two-half software pipeline on P x Q grids (same idea as getrf_lookahead_1x1) =====
        // sq carries the "prep" of a column range: pack -> column all-gather + broadcast (NCCL, nc->col) -> select ->
        // unpack -> U12 solve; near half under the update of the previous step's far half, far half (+ the already
        // factored left columns) under this step's near update.  sg carries the three update launches of a step.
        cudaStream_t sg = sm, sq = r.s_prep;
        const int64_t split_min = std::max<int64_t>(opt("la_split_min", 6144), std::max<int64_t>(2 * (int64_t)nb, 4));   // see the 1 x 1 pipeline
        auto mkev = [](std::vector<cudaEvent_t> &v, size_t n) { v.resize(n); for (auto &e : v) SLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); };
        std::vector<cudaEvent_t> pdp, pdn, pdf, gdn, gdf;
        mkev(pdp, (size_t)nsteps); mkev(pdn, (size_t)nsteps); mkev(pdf, (size_t)nsteps); mkev(gdn, (size_t)nsteps); mkev(gdf, (size_t)nsteps);
        gev.assign((size_t)6 * nsteps, nullptr);
        SLB_CUDA(cudaStreamWaitEvent(sq, evs, 0));
        int64_t bl = -1; bool far_prev = false;
        for (int k = 0; k < nsteps; ++k) {
            const int j0 = k * nb;
            const int jb = (mn - j0) < nb ? (mn - j0) : nb;
            const int pr = (rsrc + k) % P;
            const int64_t lr0 = rows_before(myrow, j0);
            const int64_t mtr = mloc - lr0;
            const int64_t lcl = numroc(j0, nb, mycol, csrc, Q);
            const int64_t lcr = numroc(j0 + jb, nb, mycol, csrc, Q);
            unsigned char *Pbuf = Pb[k & 1];
            const T *L11 = (T *)Pbuf; const int64_t ld11 = jb;
            const T *Lop = (T *)(Pbuf + hdr_bytes) + (myrow == pr ? jb : 0); const int64_t ldl = mtr > 0 ? mtr : 1;
            SwapPlan plan = plans[k & 1];
            const int64_t nright = nloc - lcr;
            const int64_t rbeg = lr0 + (myrow == pr ? jb : 0);
            const int64_t mrows = mloc - rbeg;
            const bool have_next = k + 1 < nsteps;
            const int pcn = (csrc + k + 1) % Q;
            const int jbn = have_next ? ((mn - (j0 + jb)) < nb ? (mn - (j0 + jb)) : nb) : 0;
            const int64_t nfirst = (have_next && mycol == pcn) ? (jbn < nright ? jbn : nright) : 0;
            const double glob_ms = 2.0 * (double)(M - j0 - jb) / P * (double)(N - j0 - jb) / Q * jb * Ops<T>::flop_mul / 30e12 * 1e3;
            const bool overlap = have_next && glob_ms >= overlap_min_ms;
            // one column range: [c0, c1) local columns
            auto prep_range = [&](int64_t c0, int64_t c1, bool right) {
                const int64_t ncols = c1 - c0;
                if (ncols <= 0) return;
                T *Ur = Ubuf + c0 * jb, *Or = Obuf + c0 * jb;
                if (P == 1) launch_swap_pack<T>(jb, j0, plan, rd, A, lld, c0, c1, Ur, jb, Or, jb, sq);
                else {
                    T *Cm = Cmine + c0 * jb, *Ca = Call + (int64_t)P * c0 * jb;
                    launch_swap_pack<T>(jb, j0, plan, rd, A, lld, c0, c1, Cm, jb, Or, jb, sq);
                    const size_t cnt = (size_t)jb * ncols * sizeof(T);
                    nccl_allgather(nc->col, Cm, Ca, cnt, NT_U8, sq);
                    nccl_bcast(nc->col, Or, cnt, NT_U8, pr, sq);
                    launch_swap_select<T>(jb, plan, rd, Ca, jb, (int64_t)jb * ncols, ncols, Ur, jb, sq);
                }
                launch_swap_unpack_out<T>(jb, plan, rd, A, lld, c0, c1, Or, jb, sq);
                if (right) Ops<T>::trsm(jb, ncols, L11, ld11, Ur, jb, sq);
                if (myrow == pr) launch_copy2d<T>(jb, ncols, Ur, jb, A + lr0 + c0 * lld, lld, sq);
            };
            bool a_packed = false;
            auto timed_gemm = [&](int slot, int64_t c0, int64_t c1, bool may_chunk) {
                const int64_t nn = c1 - c0;
                if (mrows <= 0 || nn <= 0) return;
                SLB_CUDA(cudaEventCreate(&gev[6 * k + slot])); SLB_CUDA(cudaEventCreate(&gev[6 * k + slot + 1]));
                SLB_CUDA(cudaEventRecord(gev[6 * k + slot], sg));
                Ops<T>::gemm(mrows, nn, jb, Lop, ldl, Ubuf + c0 * jb, jb, A + rbeg + c0 * lld, lld, sg, (may_chunk && overlap) ? chunk_opt : 0,
                             GEMM_MAIN | (a_packed ? GEMM_REUSE_A : 0));
                a_packed = Ops<T>::packs(mrows, jb);
                SLB_CUDA(cudaEventRecord(gev[6 * k + slot + 1], sg));
                gflops[k] += 2.0 * (double)mrows * (double)nn * jb * Ops<T>::flop_mul;
            };
            // ---- near | far boundary (local columns); identical on all ranks of a process column ----
            int64_t bk = nloc;
            bool resplit = false;
            if (have_next && nright >= split_min) {
                if (bl < 0 || bl - lcr < nright / 4 || bl >= nloc) { bk = lcr + ((nright / 2 + nb - 1) / nb) * nb; resplit = true; }
                else bk = bl;
                if (bk >= nloc) bk = nloc;
            }
            if (bk > bl || bl < 0) resplit = true;
            // ---- (a) sq: near half, the next panel's columns first ----
            SLB_CUDA(cudaStreamWaitEvent(sq, evp[k], 0));
            if (k > 0) SLB_CUDA(cudaStreamWaitEvent(sq, gdn[k - 1], 0));
            if (k > 0 && far_prev && resplit) SLB_CUDA(cudaStreamWaitEvent(sq, gdf[k - 1], 0));
            tmark(k, 3, sq);
            if (nfirst > 0 && lcr + nfirst < bk) {
                prep_range(lcr, lcr + nfirst, true);
                SLB_CUDA(cudaEventRecord(pdp[k], sq));
                prep_range(lcr + nfirst, bk, true);
            } else {
                prep_range(lcr, bk, true);
                SLB_CUDA(cudaEventRecord(pdp[k], sq));
            }
            SLB_CUDA(cudaEventRecord(pdn[k], sq));
            tmark(k, 4, sq);
            // ---- (b) sg: next panel's columns, hand-over to the panel stream ----
            SLB_CUDA(cudaStreamWaitEvent(sg, pdp[k], 0));
            if (have_next) {
                timed_gemm(0, lcr, lcr + nfirst, false);
                SLB_CUDA(cudaEventRecord(evn[k], sg));
                SLB_CUDA(cudaStreamWaitEvent(sp, evn[k], 0));
                panel_phase(k + 1, overlap ? gmax_opt : 0);
            }
            // ---- (c) sq: far half, then the already factored left columns (nothing waits for those but the end) ----
            if (k > 0 && far_prev) SLB_CUDA(cudaStreamWaitEvent(sq, gdf[k - 1], 0));
            if (have_next) SLB_CUDA(cudaStreamWaitEvent(sq, evn[k], 0));
            prep_range(bk, nloc, true);
            SLB_CUDA(cudaEventRecord(pdf[k], sq));
            prep_range(0, lcl, false);
            tmark(k, 5, sq);
            if (stream_io && myrow == pr) {      // block row k never changes again (later interchanges touch rows below): back to the host caller
                cudaEvent_t e; SLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                SLB_CUDA(cudaEventRecord(e, sq)); dlev.push_back(e);
                link->download(lr0, lr0 + jb, 0, nloc, e);
            }
            // ---- (d) sg: rest of near, far ----
            SLB_CUDA(cudaStreamWaitEvent(sg, pdn[k], 0));
            timed_gemm(2, lcr + nfirst, bk, true);
            SLB_CUDA(cudaEventRecord(gdn[k], sg));
            SLB_CUDA(cudaStreamWaitEvent(sg, pdf[k], 0));
            timed_gemm(4, bk, nloc, true);
            SLB_CUDA(cudaEventRecord(gdf[k], sg));
            far_prev = true;               // gdf[k] is always recorded (an empty far half completes at once)
            bl = bk;
        }
        SLB_CUDA(cudaStreamWaitEvent(sm, evp[nsteps - 1], 0));
        SLB_CUDA(cudaEventRecord(evs, sq));
        SLB_CUDA(cudaStreamWaitEvent(sm, evs, 0));
        for (auto *v : { &pdp, &pdn, &pdf, &gdn, &gdf }) for (auto &e : *v) cudaEventDestroy(e);
    } else
    for (int k = 0; k < nsteps; ++k) {
        cudaStream_t s = sm;
        const int j0 = k * nb;
        const int jb = (mn - j0) < nb ? (mn - j0) : nb;
        const int pr = (rsrc + k) % P;
        const int64_t lr0 = rows_before(myrow, j0);
        const int64_t mtr = mloc - lr0;
        const int64_t lcl = numroc(j0, nb, mycol, csrc, Q);
        const int64_t lcr = numroc(j0 + jb, nb, mycol, csrc, Q);
        const T *L11, *Lop; int64_t ldl, ld11;
        if (!multi) { T *Wp = A + lr0 + lcl * lld; L11 = Wp; ld11 = lld; Lop = Wp + jb; ldl = lld; }
        else {
            unsigned char *Pbuf = Pb[k & 1];
            L11 = (T *)Pbuf; ld11 = jb;
            Lop = (T *)(Pbuf + hdr_bytes) + (myrow == pr ? jb : 0); ldl = mtr > 0 ? mtr : 1;
        }
        if (sp != sm) SLB_CUDA(cudaStreamWaitEvent(sm, evp[k], 0));

        // =============== row interchanges + U12 ===============
        mark();
        SwapPlan plan = plans[k & 1];
        const int64_t nright = nloc - lcr;
        T *Uall = Ubuf;                      // jb x nloc, column index = local column
        if (P == 1) {
            launch_swap_pack<T>(jb, j0, plan, rd, A, lld, 0, lcl, Uall, jb, Obuf, jb, s);
            launch_swap_pack<T>(jb, j0, plan, rd, A, lld, lcr, nloc, Uall + lcr * jb, jb, Obuf + lcr * jb, jb, s);
        } else {
            launch_swap_pack<T>(jb, j0, plan, rd, A, lld, 0, lcl, Cmine, jb, Obuf, jb, s);
            launch_swap_pack<T>(jb, j0, plan, rd, A, lld, lcr, nloc, Cmine + lcr * jb, jb, Obuf + lcr * jb, jb, s);
            const size_t cnt = (size_t)jb * nloc * sizeof(T);
            if (cnt > 0) {
                nccl_allgather(nc->col, Cmine, Call, cnt, NT_U8, s);
                nccl_bcast(nc->col, Obuf, cnt, NT_U8, pr, s);
                launch_swap_select<T>(jb, plan, rd, Call, jb, (int64_t)jb * nloc, nloc, Uall, jb, s);
            }
        }
        launch_swap_unpack_out<T>(jb, plan, rd, A, lld, 0, lcl, Obuf, jb, s);
        launch_swap_unpack_out<T>(jb, plan, rd, A, lld, lcr, nloc, Obuf + lcr * jb, jb, s);
        if (myrow == pr) launch_copy2d<T>(jb, lcl, Uall, jb, A + lr0, lld, s);           // left columns: final rows
        mark();
        const bool have_next = k + 1 < nsteps;
        T *U = Uall + lcr * jb;
        const int64_t rbeg = lr0 + (myrow == pr ? jb : 0);
        const int64_t mrows = mloc - rbeg;
        if (nright > 0) {
            Ops<T>::trsm(jb, nright, L11, ld11, U, jb, s);
            if (myrow == pr) launch_copy2d<T>(jb, nright, U, jb, A + lr0 + lcr * lld, lld, s);
        }
        // =============== trailing update (+ hand-over of the next panel's columns) ===============
        mark();
        bool a_packed = false;
        auto timed_gemm = [&](int slot, int64_t c_off, int64_t nn, int chunk) {
            if (mrows <= 0 || nn <= 0) return;
            SLB_CUDA(cudaEventCreate(&gev[4 * k + slot])); SLB_CUDA(cudaEventCreate(&gev[4 * k + slot + 1]));
            SLB_CUDA(cudaEventRecord(gev[4 * k + slot], s));
            Ops<T>::gemm(mrows, nn, jb, Lop, ldl, U + c_off * jb, jb, A + rbeg + (lcr + c_off) * lld, lld, s, chunk, GEMM_MAIN | (a_packed ? GEMM_REUSE_A : 0));
            a_packed = Ops<T>::packs(mrows, jb);
            SLB_CUDA(cudaEventRecord(gev[4 * k + slot + 1], s));
            gflops[k] += 2.0 * (double)mrows * (double)nn * jb * Ops<T>::flop_mul;
        };
        if (!have_next) { timed_gemm(0, 0, nright, 0); mark(); mark(); continue; }
        const int pcn = (csrc + k + 1) % Q;
        const int jbn = (mn - (j0 + jb)) < nb ? (mn - (j0 + jb)) : nb;
        if (la) {
            const int64_t nfirst = (mycol == pcn) ? (jbn < nright ? jbn : nright) : 0;
            const int64_t rest = nright - nfirst;
            const double rest_ms = 2.0 * (double)(mrows > 0 ? mrows : 0) * (double)rest * jb * Ops<T>::flop_mul / 28e12 * 1e3;
            // the overlap decision must be the same on every rank of the grid: use the global trailing size
            const double glob_ms = 2.0 * (double)(M - j0 - jb) / P * (double)(N - j0 - jb) / Q * jb * Ops<T>::flop_mul / 28e12 * 1e3;
            const bool overlap = glob_ms >= overlap_min_ms;
            (void)rest_ms;
            timed_gemm(0, 0, nfirst, 0);                                  // next panel's columns first
            SLB_CUDA(cudaEventRecord(evn[k], sm));
            SLB_CUDA(cudaStreamWaitEvent(sp, evn[k], 0));
            panel_phase(k + 1, overlap ? gmax_opt : 0);
            timed_gemm(2, nfirst, rest, overlap ? chunk_opt : 0);
        } else {
            timed_gemm(0, 0, nright, 0);
            mark();
            panel_phase(k + 1, 0);
            mark();
        }
    }
    if (sp != sm) SLB_CUDA(cudaStreamWaitEvent(sm, evp[nsteps - 1], 0));
    SLB_CUDA(cudaEventRecord(ev1, sm));
    SLB_CUDA(cudaMemcpyAsync(ipiv_glob_host, ipiv_dev, (size_t)mn * sizeof(int), cudaMemcpyDeviceToHost, sm));
    int info_local = 0;
    SLB_CUDA(cudaMemcpyAsync(&info_local, info_dev, sizeof(int), cudaMemcpyDeviceToHost, sm));
    SLB_CUDA(cudaStreamSynchronize(sm));
    if (sp != sm) SLB_CUDA(cudaStreamSynchronize(sp));
    if (link) {
        if (stream_io && pipe) { const int64_t l0 = numroc(mn, nb, myrow, rsrc, P); link->download(l0, mloc, 0, nloc, nullptr); }
        else link->download(0, mloc, 0, nloc, nullptr);
        link->finish();
        g_last_lu.host_written = true;
        for (auto &e : dlev) cudaEventDestroy(e);
        if (stream_io && pipe) counter_add("e2e_download_overlapped", 1);
    }
    float ms = 0; SLB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    g_last_lu.factor_ms = ms;
    g_last_lu.update_ms = 0; g_last_lu.update_flops = 0; g_last_lu.update_launches = 0;
    if (trace && pipe) {
        // per step: update launches (g0 | near | far) on the update stream, the gaps before them, panel-phase pieces, prep pieces
        double t_upd = 0, t_gap = 0, t_pf = 0, t_pb = 0, t_pn = 0, t_pfar = 0;
        cudaEvent_t prev_end = ev0;
        fprintf(stderr, "la_trace[%d,%d]: k | gap g0 near far | panel: factor+col bcast_row | prep: near far+left  (ms)\n", myrow, mycol);
        for (int k = 0; k < nsteps; ++k) {
            float g[3] = { 0, 0, 0 }, gap = 0, pf = 0, pbr = 0, pn = 0, pfar = 0;
            cudaEvent_t first = nullptr, last = nullptr;
            for (int q = 0; q < 3; ++q)
                if (gev[(size_t)6 * k + 2 * q]) {
                    SLB_CUDA(cudaEventElapsedTime(&g[q], gev[(size_t)6 * k + 2 * q], gev[(size_t)6 * k + 2 * q + 1]));
                    if (!first) first = gev[(size_t)6 * k + 2 * q];
                    last = gev[(size_t)6 * k + 2 * q + 1];
                }
            if (first) { SLB_CUDA(cudaEventElapsedTime(&gap, prev_end, first)); prev_end = last; }
            auto el = [&](int a, int b) { float t = 0; if (tev[(size_t)8 * k + a] && tev[(size_t)8 * k + b]) SLB_CUDA(cudaEventElapsedTime(&t, tev[(size_t)8 * k + a], tev[(size_t)8 * k + b])); return t; };
            pf = el(0, 1); pbr = el(1, 2); pn = el(3, 4); pfar = el(4, 5);
            t_upd += g[0] + g[1] + g[2]; t_gap += gap; t_pf += pf; t_pb += pbr; t_pn += pn; t_pfar += pfar;
            if (k % 16 == 0 || k >= nsteps - 4)
                fprintf(stderr, "la_trace[%d,%d]: %d | %.2f %.2f %.2f %.2f | %.2f %.2f | %.2f %.2f\n", myrow, mycol, k, gap, g[0], g[1], g[2], pf, pbr, pn, pfar);
        }
        fprintf(stderr, "la_trace[%d,%d]: total %.1f ms: update launches %.1f + gaps on the update stream %.1f; panel phases: factor+column %.1f, row broadcast %.1f; prep: near %.1f, far+left %.1f\n",
                myrow, mycol, ms, t_upd, t_gap, t_pf, t_pb, t_pn, t_pfar);
    }
    for (auto &e : tev) if (e) cudaEventDestroy(e);
    for (size_t i = 0; i + 1 < gev.size(); i += 2)
        if (gev[i]) {
            float t = 0; SLB_CUDA(cudaEventElapsedTime(&t, gev[i], gev[i + 1]));
            g_last_lu.update_ms += t; g_last_lu.update_launches += 1;
            cudaEventDestroy(gev[i]); cudaEventDestroy(gev[i + 1]);
        }
    for (int k = 0; k < nsteps; ++k) g_last_lu.update_flops += gflops[k];
    if (prof) {   // 5 marks per step: [swap][trsm][gemm][next panel]
        double tp = 0, tsw = 0, ttr = 0, tg = 0;
        for (size_t i = 0; i + 4 < pev.size(); i += 5) {
            float a, b, c, d;
            SLB_CUDA(cudaEventElapsedTime(&a, pev[i], pev[i + 1])); SLB_CUDA(cudaEventElapsedTime(&b, pev[i + 1], pev[i + 2]));
            SLB_CUDA(cudaEventElapsedTime(&c, pev[i + 2], pev[i + 3])); SLB_CUDA(cudaEventElapsedTime(&d, pev[i + 3], pev[i + 4]));
            tsw += a; ttr += b; tg += c; tp += d;
        }
        for (auto e : pev) cudaEventDestroy(e);
        counter_add("prof_panel_us", (int64_t)(tp * 1e3)); counter_add("prof_swap_us", (int64_t)(tsw * 1e3));
        counter_add("prof_trsm_us", (int64_t)(ttr * 1e3)); counter_add("prof_gemm_us", (int64_t)(tg * 1e3));
    }
    for (auto &e : evp) cudaEventDestroy(e);
    for (auto &e : evn) cudaEventDestroy(e);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(evs);
    // INFO: first zero pivot is known on the diagonal owners only -> min over the grid (SRC/pdgetrf.f:297-302)
    int inf = info_local == 0 ? mn + 1 : info_local;
    inf = grid_imin(g, 'A', inf);
    *info_host = inf == mn + 1 ? 0 : inf;
    return 0;
}

template int getrf_device<double>(Grid *, int, int, double *, int64_t, int, int, int, int *, int *, HostLink *);
template int getrf_device<zcomplex>(Grid *, int, int, zcomplex *, int64_t, int, int, int, int *, int *, HostLink *);

}  // namespace slb
