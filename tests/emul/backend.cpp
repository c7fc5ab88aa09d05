// backend.cpp -- TEST INFRASTRUCTURE (see tests/emul/cuda_runtime.h).  Stands in, by documented contract, for the pieces of
// the product that only exist on a GPU, so that the host logic of the SURVEY 8(f) files can be exercised on a CPU-only machine:
//   * runtime.cu            -> rt(), workspace(): "device" memory is host memory
//   * gemm*.cu / trsm.cu    -> launch_dgemm_minus (C -= A B), launch_dtrsm_llnu (B <- unit_lower(L)^-1 B), launch_copy2d
//   * lu.cu / solve.cu      -> getrf_device / getrs_device on a P x Q grid: gather, serial LAPACK-style algorithm, scatter
//   * ncclw.cpp             -> broadcast / all-gather / all-reduce over the TCP control plane (hostcomm.cpp), same call order
// Nothing here is linked into libscalapack_b200.so, and nothing here is timed or shipped.
#include "common.h"
#include "kernels.cuh"
#include "lu.h"
#include "ncclw.h"
#include "stage.h"

#include <cmath>
#include <map>

namespace slb {

struct ncclComm { Grid *g; char scope; int id; };     // id: 0 all, 1 row, 2 col, 3 colp (independent message sequences)

static Runtime g_rt;
Runtime &rt() { g_rt.cuda_ok = true; g_rt.device = 0; g_rt.sm_count = 148; g_rt.smem_optin = 227 * 1024; return g_rt; }
bool cuda_available() { return true; }

struct WsEntry { void *p = nullptr; size_t bytes = 0; };
static std::map<std::string, WsEntry> g_ws;
void *workspace(const char *name, size_t bytes, bool zero_on_alloc)
{
    WsEntry &e = g_ws[name];
    if (e.bytes < bytes) {
        cudaFree(e.p);                                       // like the product: a grown workspace moves, stale pointers die
        e.bytes = bytes + 64;
        cudaMalloc(&e.p, e.bytes);                           // registered as device memory and poisoned (cuda_runtime.h)
        if (zero_on_alloc) memset(e.p, 0, e.bytes);
    }
    return e.p;
}
void workspace_release_all() { for (auto &kv : g_ws) cudaFree(kv.second.p); g_ws.clear(); }

// the pointer arguments of the contract functions below are device pointers on a GPU
#define DEV(p) emul_need_dev((const void *)(p), "a device routine was given a pointer that is not device memory (" #p ")")


// ---- arithmetic helpers on double / zcomplex ---------------------------------------------------------------------------
static inline double cmul_sub(double c, double a, double b) { return c - a * b; }
static inline zcomplex cmul_sub(zcomplex c, zcomplex a, zcomplex b)
{ return make_double2(c.x - (a.x * b.x - a.y * b.y), c.y - (a.x * b.y + a.y * b.x)); }
static inline double cabs1(double a) { return fabs(a); }
static inline double cabs1(zcomplex a) { return fabs(a.x) + fabs(a.y); }
static inline bool is0(double a) { return a == 0.0; }
static inline bool is0(zcomplex a) { return a.x == 0.0 && a.y == 0.0; }
static inline double cdiv(double a, double b) { return a / b; }
static inline zcomplex cdiv(zcomplex a, zcomplex b)
{ double d = b.x * b.x + b.y * b.y; return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d); }
static inline double cconj(double a) { return a; }
static inline zcomplex cconj(zcomplex a) { return make_double2(a.x, -a.y); }

template <typename T>
static void gemm_minus(int64_t M, int64_t N, int K, const T *A, int64_t lda, const T *B, int64_t ldb, T *C, int64_t ldc)
{
    // k runs DOWNWARDS: another summation order than the oracle's, like the tensor-core kernel's (tolerances must not depend on it)
    for (int64_t j = 0; j < N; ++j)
        for (int k = K - 1; k >= 0; --k) {
            const T b = B[k + j * ldb];
            for (int64_t i = 0; i < M; ++i) C[i + j * ldc] = cmul_sub(C[i + j * ldc], A[i + (int64_t)k * lda], b);
        }
}
void launch_dgemm_minus(int64_t M, int64_t N, int K, const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc,
                        cudaStream_t, int, int)
{ if (M <= 0 || N <= 0 || K <= 0) return; DEV(A); DEV(B); DEV(C); gemm_minus<double>(M, N, K, A, lda, B, ldb, C, ldc); counter_add("kernel_launches", 1); }
void launch_zgemm_minus(int64_t M, int64_t N, int K, const zcomplex *A, int64_t lda, const zcomplex *B, int64_t ldb, zcomplex *C, int64_t ldc,
                        cudaStream_t, int, int)
{ if (M <= 0 || N <= 0 || K <= 0) return; DEV(A); DEV(B); DEV(C); gemm_minus<zcomplex>(M, N, K, A, lda, B, ldb, C, ldc); counter_add("kernel_launches", 1); }

template <typename T>
static void trsm_llnu(int jb, int64_t n, const T *L, int64_t ldl, T *B, int64_t ldb)
{
    for (int64_t j = 0; j < n; ++j)
        for (int k = 0; k < jb; ++k) {
            const T x = B[k + j * ldb];
            for (int i = k + 1; i < jb; ++i) B[i + j * ldb] = cmul_sub(B[i + j * ldb], L[i + (int64_t)k * ldl], x);
        }
}
void launch_dtrsm_llnu(int jb, int64_t n, const double *L, int64_t ldl, double *B, int64_t ldb, cudaStream_t)
{ if (jb <= 0 || n <= 0) return; DEV(L); DEV(B); trsm_llnu<double>(jb, n, L, ldl, B, ldb); counter_add("kernel_launches", 1); }
void launch_ztrsm_llnu(int jb, int64_t n, const zcomplex *L, int64_t ldl, zcomplex *B, int64_t ldb, cudaStream_t)
{ if (jb <= 0 || n <= 0) return; DEV(L); DEV(B); trsm_llnu<zcomplex>(jb, n, L, ldl, B, ldb); counter_add("kernel_launches", 1); }

template <typename T>
void launch_copy2d(int64_t rows, int64_t cols, const T *src, int64_t lds, T *dst, int64_t ldd, cudaStream_t)
{
    if (rows <= 0 || cols <= 0) return;
    DEV(src); DEV(dst);
    for (int64_t c = 0; c < cols; ++c) for (int64_t i = 0; i < rows; ++i) dst[i + c * ldd] = src[i + c * lds];
    counter_add("kernel_launches", 1);
}
template void launch_copy2d<double>(int64_t, int64_t, const double *, int64_t, double *, int64_t, cudaStream_t);
template void launch_copy2d<zcomplex>(int64_t, int64_t, const zcomplex *, int64_t, zcomplex *, int64_t, cudaStream_t);

// ---- PDGETRS kernels by contract (kernels.cuh) ------------------------------------------------------------------------------
template <typename T>
void launch_trsv_block(int kb, const T *A, int64_t lda, T *X, int64_t ldx, int nrhs, int mode, cudaStream_t)
{
    DEV(A); DEV(X);
    const bool upper = mode & TRSV_UPPER, trans = mode & TRSV_TRANS, cj = (mode & TRSV_CONJ) != 0;
    const bool unit = !upper && !(mode & TRSV_NONUNIT_L);
    auto op = [&](int i, int k) { T v = trans ? A[k + (int64_t)i * lda] : A[i + (int64_t)k * lda]; return cj ? cconj(v) : v; };   // element (i, k) of op(A)
    const bool fwd = upper == trans;
    for (int c = 0; c < nrhs; ++c) {
        T *x = X + (int64_t)c * ldx;
        for (int q = 0; q < kb; ++q) {
            const int i = fwd ? q : kb - 1 - q;
            T v = x[i];
            if (fwd) for (int k = 0; k < i; ++k) v = cmul_sub(v, op(i, k), x[k]);
            else for (int k = i + 1; k < kb; ++k) v = cmul_sub(v, op(i, k), x[k]);
            x[i] = unit ? v : cdiv(v, op(i, i));
        }
    }
    counter_add("kernel_launches", 1);
}
template <typename T>
void launch_gemv_minus(int64_t rows, int kb, const T *A, int64_t lda, const T *X, int64_t ldx, T *Y, int64_t ldy, int nrhs, cudaStream_t)
{
    if (rows <= 0 || kb <= 0 || nrhs <= 0) return;
    DEV(A); DEV(X); DEV(Y);
    for (int c = 0; c < nrhs; ++c) for (int k = 0; k < kb; ++k) for (int64_t i = 0; i < rows; ++i)
        Y[i + (int64_t)c * ldy] = cmul_sub(Y[i + (int64_t)c * ldy], A[i + (int64_t)k * lda], X[k + (int64_t)c * ldx]);
    counter_add("kernel_launches", 1);
}
template <typename T>
void launch_gemvt_minus(int kb, int64_t ncols, const T *A, int64_t lda, const T *X, int64_t ldx, T *Y, int64_t ldy, int nrhs, bool conj, cudaStream_t)
{
    if (ncols <= 0 || kb <= 0 || nrhs <= 0) return;
    DEV(A); DEV(X); DEV(Y);
    for (int c = 0; c < nrhs; ++c) for (int64_t j = 0; j < ncols; ++j) for (int i = 0; i < kb; ++i)
        Y[j + (int64_t)c * ldy] = cmul_sub(Y[j + (int64_t)c * ldy], conj ? cconj(A[i + j * lda]) : A[i + j * lda], X[i + (int64_t)c * ldx]);
    counter_add("kernel_launches", 1);
}
#define INST(T)                                                                                                  \
    template void launch_trsv_block<T>(int, const T *, int64_t, T *, int64_t, int, int, cudaStream_t);           \
    template void launch_gemv_minus<T>(int64_t, int, const T *, int64_t, const T *, int64_t, T *, int64_t, int, cudaStream_t); \
    template void launch_gemvt_minus<T>(int, int64_t, const T *, int64_t, const T *, int64_t, T *, int64_t, int, bool, cudaStream_t);
INST(double)
INST(zcomplex)
#undef INST

// ---- block-cyclic gather / scatter of an M x N matrix over the grid (local windows start at A, first block on (rsrc, csrc)) ----
template <typename T>
static std::vector<T> gather_bc(Grid *g, int M, int N, const T *A, int64_t lld, int nb, int rsrc, int csrc)
{
    const int P = g->nprow, Q = g->npcol, np = P * Q;
    const int64_t mmax = numroc(M, nb, 0, 0, P), nmax = numroc(N, nb, 0, 0, Q);      // process 0 of a dimension holds the most
    const int64_t mloc = numroc(M, nb, g->myrow, rsrc, P), nloc = numroc(N, nb, g->mycol, csrc, Q);
    std::vector<T> mine((size_t)(mmax * nmax)), all((size_t)(mmax * nmax) * np);
    memset(mine.data(), 0, mine.size() * sizeof(T));
    for (int64_t c = 0; c < nloc; ++c) for (int64_t i = 0; i < mloc; ++i) mine[(size_t)(i + c * mmax)] = A[i + c * lld];
    if (np > 1) grid_allgather(g, 'A', mine.data(), all.data(), mine.size() * sizeof(T)); else all = mine;
    std::vector<T> G((size_t)M * N);
    for (int64_t j = 0; j < N; ++j) {
        const int pc = indxg2p((int)j + 1, nb, csrc, Q); const int64_t jl = indxg2l((int)j + 1, nb, Q) - 1;
        for (int64_t i = 0; i < M; ++i) {
            const int pr = indxg2p((int)i + 1, nb, rsrc, P); const int64_t il = indxg2l((int)i + 1, nb, P) - 1;
            G[(size_t)(i + j * M)] = all[(size_t)(pr * Q + pc) * (size_t)(mmax * nmax) + (size_t)(il + jl * mmax)];
        }
    }
    return G;
}
template <typename T>
static void scatter_bc(Grid *g, int M, int N, const std::vector<T> &G, T *A, int64_t lld, int nb, int rsrc, int csrc)
{
    const int P = g->nprow, Q = g->npcol;
    for (int64_t j = 0; j < N; ++j) {
        if (indxg2p((int)j + 1, nb, csrc, Q) != g->mycol) continue;
        const int64_t jl = indxg2l((int)j + 1, nb, Q) - 1;
        for (int64_t i = 0; i < M; ++i) {
            if (indxg2p((int)i + 1, nb, rsrc, P) != g->myrow) continue;
            A[(indxg2l((int)i + 1, nb, P) - 1) + jl * lld] = G[(size_t)(i + j * M)];
        }
    }
}

// solve.cu (the PDGETRS orchestration) is compiled for real; what it calls from solve_kernels.cu / solve_fast.cu by contract:
bool getrs_fast_applies(int, int, char, int, int) { return false; }      // the 1x1 two-stream fast path is GPU-only
void getrs_fast_device(int, int, const double *, int64_t, int, double *) { fatal("emulation: the solve fast path is not modelled"); }
double solve_fast_probe(int, int, int64_t, const double *, int64_t, int, int) { return 0; }
template <typename T>
void launch_gather_rows(int64_t n, const int *perm, const T *src, int64_t lds, T *dst, int64_t ldd, int nrhs, cudaStream_t, bool scatter)
{
    if (n <= 0 || nrhs <= 0) return;
    DEV(perm); DEV(src); DEV(dst);
    for (int c = 0; c < nrhs; ++c)
        for (int64_t i = 0; i < n; ++i) {
            const int p = perm[i];
            if (p < 0) continue;
            if (scatter) dst[p + (int64_t)c * ldd] = src[i + (int64_t)c * lds]; else dst[i + (int64_t)c * ldd] = src[p + (int64_t)c * lds];
        }
    counter_add("kernel_launches", 1);
}
template void launch_gather_rows<double>(int64_t, const int *, const double *, int64_t, double *, int64_t, int, cudaStream_t, bool);
template void launch_gather_rows<zcomplex>(int64_t, const int *, const zcomplex *, int64_t, zcomplex *, int64_t, int, cudaStream_t, bool);

// ---- the LU's own kernels by contract (kernels.cuh; panel.cu, swap.cu) so that lu.cu / api.cu run their REAL orchestration --------
static inline double crecip(double x) { return 1.0 / x; }
static inline zcomplex crecip(zcomplex z)          // Smith's division, like devmath.cuh t_recip
{
    if (fabs(z.x) >= fabs(z.y)) { const double t = z.y / z.x, d = z.x + z.y * t; return make_double2(1.0 / d, -t / d); }
    const double t = z.x / z.y, d = z.x * t + z.y;
    return make_double2(t / d, -1.0 / d);
}
static inline double cmul(double a, double b) { return a * b; }
static inline zcomplex cmul(zcomplex a, zcomplex b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
size_t panel_work_bytes(int) { return 256; }
// PDGETF2 on the m x jb panel W (rows in virtual order, row v = global row map.g0 + v): pivot = max |.| (complex |Re|+|Im|), ties to
// the lower process row then the lower row (PBLAS/SRC/pdamax_.c:436-465); swap over the panel's jb columns; reciprocal scaling
template <typename T>
static void panel_contract(int m, int jb, T *W, int64_t ldw, const PanelRowMap &map, int *ipiv_out, int *info_out, int info_offset)
{
    DEV(W); DEV(ipiv_out); DEV(info_out);
    const int mn = m < jb ? m : jb;
    auto prow = [&](int v) { return (map.rsrc + (map.g0 + v) / map.nb) % map.nprow; };
    for (int j = 0; j < mn; ++j) {
        int p = j; double best = cabs1(W[j + (int64_t)j * ldw]);
        for (int v = j + 1; v < m; ++v) {
            const double a = cabs1(W[v + (int64_t)j * ldw]);
            if (a > best || (a == best && prow(v) < prow(p))) { best = a; p = v; }
        }
        ipiv_out[j] = map.g0 + p + 1;
        if (!is0(W[p + (int64_t)j * ldw])) {
            if (p != j) for (int c = 0; c < jb; ++c) { T t = W[j + (int64_t)c * ldw]; W[j + (int64_t)c * ldw] = W[p + (int64_t)c * ldw]; W[p + (int64_t)c * ldw] = t; }
            const T r = crecip(W[j + (int64_t)j * ldw]);
            for (int v = j + 1; v < m; ++v) W[v + (int64_t)j * ldw] = cmul(W[v + (int64_t)j * ldw], r);
        } else if (*info_out == 0) *info_out = info_offset + j + 1;
        for (int c = j + 1; c < jb; ++c) {
            const T u = W[j + (int64_t)c * ldw];
            for (int v = j + 1; v < m; ++v) W[v + (int64_t)c * ldw] = cmul_sub(W[v + (int64_t)c * ldw], W[v + (int64_t)j * ldw], u);
        }
    }
    counter_add("kernel_launches", 1);
}
void launch_dpanel(int m, int jb, double *W, int64_t ldw, const PanelRowMap &map, int *ipiv_out, int *info_out, int info_offset, void *, cudaStream_t, int)
{ panel_contract<double>(m, jb, W, ldw, map, ipiv_out, info_out, info_offset); }
void launch_zpanel(int m, int jb, zcomplex *W, int64_t ldw, const PanelRowMap &map, int *ipiv_out, int *info_out, int info_offset, void *, cudaStream_t, int)
{ panel_contract<zcomplex>(m, jb, W, ldw, map, ipiv_out, info_out, info_offset); }

// the net permutation of one block of interchanges, as documented in kernels.cuh (SwapPlan)
void launch_swap_plan(int j0, int jb, const int *ipiv_blk, SwapPlan plan, cudaStream_t)
{
    DEV(ipiv_blk); DEV(plan.top_src);
    auto trace = [&](int pos) { for (int s = jb - 1; s >= 0; --s) { const int r = j0 + s, p = ipiv_blk[s] - 1; if (pos == r) pos = p; else if (pos == p) pos = r; } return pos; };
    for (int t = 0; t < jb; ++t) {
        plan.top_src[t] = trace(j0 + t);
        const int p = ipiv_blk[t] - 1;
        int dst = -1, src = 0;
        if (p >= j0 + jb) {
            bool first = true;
            for (int s = 0; s < t; ++s) if (ipiv_blk[s] - 1 == p) { first = false; break; }
            if (first) { dst = p; src = trace(p) - j0; }
        }
        plan.out_dst[t] = dst; plan.out_src[t] = src;
    }
    counter_add("kernel_launches", 1);
}
static inline int row_owner(const RowDist &rd, int g) { return (rd.rsrc + g / rd.nb) % rd.nprow; }
static inline int64_t row_local(const RowDist &rd, int g) { return (int64_t)rd.nb * (g / ((int64_t)rd.nb * rd.nprow)) + g % rd.nb - rd.shift; }
template <typename T>
void launch_swap_pack(int jb, int j0, SwapPlan plan, RowDist rd, const T *A, int64_t lda, int64_t c0, int64_t c1, T *Ubuf, int64_t ldu, T *Obuf, int64_t ldo, cudaStream_t)
{
    if (c1 <= c0 || jb <= 0) return;
    DEV(plan.top_src);
    bool checked = false;                                     // a process that owns none of the rows never reads A: check at first use
    auto touch = [&]() { if (!checked) { DEV(A); checked = true; } };
    const bool own_top = row_owner(rd, j0) == rd.myrow;
    for (int64_t c = c0; c < c1; ++c)
        for (int t = 0; t < jb; ++t) {
            const int src = plan.top_src[t];
            if (row_owner(rd, src) == rd.myrow) { touch(); DEV(Ubuf); Ubuf[t + (c - c0) * ldu] = A[row_local(rd, src) + c * lda]; }
            if (own_top && Obuf && plan.out_dst[t] >= 0) { touch(); DEV(Obuf); Obuf[t + (c - c0) * ldo] = A[row_local(rd, j0 + plan.out_src[t]) + c * lda]; }
        }
    counter_add("kernel_launches", 1);
}
template <typename T>
void launch_swap_unpack_out(int jb, SwapPlan plan, RowDist rd, T *A, int64_t lda, int64_t c0, int64_t c1, const T *Obuf, int64_t ldo, cudaStream_t)
{
    if (c1 <= c0 || jb <= 0) return;
    DEV(plan.out_dst);
    for (int64_t c = c0; c < c1; ++c)
        for (int t = 0; t < jb; ++t) {
            const int d = plan.out_dst[t];
            if (d >= 0 && row_owner(rd, d) == rd.myrow) { if (c == c0) { DEV(A); DEV(Obuf); } A[row_local(rd, d) + c * lda] = Obuf[t + (c - c0) * ldo]; }
        }
    counter_add("kernel_launches", 1);
}
void swap_grid_override(int) {}
template <typename T>
void launch_swap_select(int jb, SwapPlan plan, RowDist rd, const T *Call, int64_t ldc, int64_t stride_p, int64_t ncols, T *U, int64_t ldu, cudaStream_t)
{
    if (ncols <= 0 || jb <= 0) return;
    DEV(Call); DEV(U);
    for (int64_t c = 0; c < ncols; ++c)
        for (int t = 0; t < jb; ++t) U[t + c * ldu] = Call[(int64_t)row_owner(rd, plan.top_src[t]) * stride_p + t + c * ldc];
    counter_add("kernel_launches", 1);
}
template <typename T>
void launch_rows_bc(int64_t rows, int cols, T *L, int64_t ldl, int64_t l0, T *G, int64_t ldg, int64_t gshift, int nb, int nprow, int prow_rel, int to_global, cudaStream_t)
{
    if (rows <= 0 || cols <= 0) return;
    DEV(L); DEV(G);
    for (int64_t i = 0; i < rows; ++i) {
        const int64_t l = l0 + i, gi = ((l / nb) * nprow + prow_rel) * nb + l % nb - gshift;
        for (int c = 0; c < cols; ++c) { if (to_global) G[gi + (int64_t)c * ldg] = L[i + (int64_t)c * ldl]; else L[i + (int64_t)c * ldl] = G[gi + (int64_t)c * ldg]; }
    }
    counter_add("kernel_launches", 1);
}
#define INST(T)                                                                                                        \
    template void launch_swap_pack<T>(int, int, SwapPlan, RowDist, const T *, int64_t, int64_t, int64_t, T *, int64_t, T *, int64_t, cudaStream_t); \
    template void launch_swap_unpack_out<T>(int, SwapPlan, RowDist, T *, int64_t, int64_t, int64_t, const T *, int64_t, cudaStream_t); \
    template void launch_swap_select<T>(int, SwapPlan, RowDist, const T *, int64_t, int64_t, int64_t, T *, int64_t, cudaStream_t); \
    template void launch_rows_bc<T>(int64_t, int, T *, int64_t, int64_t, T *, int64_t, int64_t, int, int, int, int, cudaStream_t);
INST(double)
INST(zcomplex)
#undef INST
bool dgemm_takes_packed(int64_t, int, int) { return false; }
bool zgemm_takes_packed(int64_t, int, int) { return false; }
// the test-matrix generators and micro-benchmarks are GPU-only tools of the test driver
void launch_pdmatgen_local(int, int, int, int, double *, int64_t, int, int, int, int, int, int, int, cudaStream_t) { fatal("emulation: the on-device generators are not modelled"); }
void launch_matgen64_local(int64_t, int64_t, int, int, double *, int64_t, int, int, uint64_t, int, int, int, int, int, cudaStream_t) { fatal("emulation: the on-device generators are not modelled"); }
void launch_gen_matvec(int64_t, int, uint64_t, int, int, int, int, int, const double *, double *, double *, cudaStream_t) { fatal("emulation: the on-device generators are not modelled"); }
double bench_dmma_peak_tflops(int) { return 0; }
double bench_dfma_peak_tflops(int) { return 0; }
double bench_copy_gbs(size_t) { return 0; }

// HostLink (stage.cu) by contract.  Downloads happen at once (everything recorded before them has completed: the emulation runs in
// program order).  Uploads can ARRIVE LATE: with SLB200_EMUL_SLAB_DELAY = d > 0 ticket t only lands in "device" memory after
// (7 t + 3) mod (d + 1) polls of done(t), or when something waits for it -- until then the staging copy holds poison, so a sweep that
// touches a slab before it has joined (lu.cu's replay logic) produces garbage instead of passing by luck.
struct HostLink::Impl {
    struct Up { int64_t r0, r1, c0, c1; int polls_left; bool landed; };
    std::vector<Up> ups;
};
bool host_ptr_is_pinned(const void *) { return false; }
HostLink::HostLink(const HostMat &h, void *dev, int64_t ldd) : im_(new Impl()), h_(h), dev_(dev), ldd_(ldd) {}
HostLink::~HostLink() {}
int HostLink::upload(int64_t r0, int64_t r1, int64_t c0, int64_t c1)
{
    static const int delay = getenv("SLB200_EMUL_SLAB_DELAY") ? atoi(getenv("SLB200_EMUL_SLAB_DELAY")) : 0;
    const int t = (int)im_->ups.size();
    im_->ups.push_back(Impl::Up{ r0, r1, c0, c1, delay > 0 ? (7 * t + 3) % (delay + 1) : 0, false });
    up_bytes_ += (r1 - r0) * (c1 - c0) * (int64_t)h_.elem;
    if (im_->ups.back().polls_left == 0) wait(t);
    return t;
}
int HostLink::download(int64_t r0, int64_t r1, int64_t c0, int64_t c1, cudaEvent_t)
{
    for (int64_t c = c0; c < c1; ++c) memcpy((char *)h_.p + (size_t)(r0 + c * h_.ld) * h_.elem, (const char *)dev_ + (size_t)(r0 + c * ldd_) * h_.elem, (size_t)(r1 - r0) * h_.elem);
    down_bytes_ += (r1 - r0) * (c1 - c0) * (int64_t)h_.elem;
    return -1;
}
void HostLink::wait(int t)
{
    if (t < 0 || t >= (int)im_->ups.size()) return;
    Impl::Up &u = im_->ups[(size_t)t];
    if (u.landed) return;
    for (int64_t c = u.c0; c < u.c1; ++c) memcpy((char *)dev_ + (size_t)(u.r0 + c * ldd_) * h_.elem, (const char *)h_.p + (size_t)(u.r0 + c * h_.ld) * h_.elem, (size_t)(u.r1 - u.r0) * h_.elem);
    u.landed = true;
}
bool HostLink::done(int t)
{
    if (t < 0 || t >= (int)im_->ups.size()) return true;
    Impl::Up &u = im_->ups[(size_t)t];
    if (!u.landed && --u.polls_left <= 0) { wait(t); if (getenv("SLB200_EMUL_TRACE")) fprintf(stderr, "emulated HostLink: slab %d (columns %lld..%lld) arrived late\n", t, (long long)u.c0, (long long)u.c1); }
    return u.landed;
}
void HostLink::stream_wait(int t, cudaStream_t) { wait(t); }
void HostLink::finish() { for (int t = 0; t < (int)im_->ups.size(); ++t) wait(t); }

// ---- NCCL wrappers over the TCP control plane ------------------------------------------------------------------------------
static size_t tsize(NcclType t) { return t == NT_F64 ? 8 : (t == NT_I32 ? 4 : 1); }
NcclComms *nccl_create(Grid *g)
{
    NcclComms *c = new NcclComms();
    c->all = new ncclComm{ g, 'A', 0 }; c->row = new ncclComm{ g, 'R', 1 }; c->col = new ncclComm{ g, 'C', 2 }; c->colp = new ncclComm{ g, 'C', 3 };
    return c;
}
void nccl_destroy(NcclComms *c) { if (!c) return; delete c->all; delete c->row; delete c->col; delete c->colp; delete c; }
void nccl_group_start() {}
void nccl_group_end() {}
void nccl_bcast(ncclComm_t_ comm, void *buf, size_t count, NcclType t, int root, cudaStream_t)
{
    DEV(buf);
    const int np = grid_scope_size(comm->g, comm->scope);
    const size_t len = count * tsize(t);
    if (np <= 1 || len == 0) return;
    std::vector<char> all(len * (size_t)np);
    grid_allgather(comm->g, comm->scope, buf, all.data(), len);
    memcpy(buf, all.data() + (size_t)root * len, len);
}
// point to point: a two-member exchange over the control plane, keyed by (communicator, sender, receiver, message number).
// Blocking like everything here: the program order of lu.cu (receives posted in peer order, senders independent) cannot deadlock.
static std::map<uint64_t, uint64_t> g_p2p_seq;
static void p2p(ncclComm_t_ comm, int src, int dst, bool sending, const void *sendbuf, void *recvbuf, size_t len)
{
    Grid *g = comm->g;
    const uint64_t coord = comm->scope == 'R' ? (uint64_t)g->myrow : (comm->scope == 'C' ? (uint64_t)g->mycol : 0);
    const uint64_t chan = ((uint64_t)((g->uid + 1) & 0xff) << 20) | ((uint64_t)comm->id << 18) | ((coord & 0x3f) << 12) | ((uint64_t)(src & 0x3f) << 6) | (uint64_t)(dst & 0x3f);
    const uint64_t seq = g_p2p_seq[chan]++;
    const uint64_t key = ((uint64_t)1 << 62) | (chan << 24) | (seq & 0xffffff);      // bit 62: never a key of grid_allgather
    std::vector<char> in(len, 0), out(2 * len);
    if (sending) memcpy(in.data(), sendbuf, len);
    hc_allgather(key, 2, sending ? 0 : 1, in.data(), out.data(), len);
    if (!sending) memcpy(recvbuf, out.data(), len);
}
void nccl_send(ncclComm_t_ comm, const void *buf, size_t count, NcclType t, int peer, cudaStream_t)
{ DEV(buf); if (count) p2p(comm, grid_scope_index(comm->g, comm->scope), peer, true, buf, nullptr, count * tsize(t)); }
void nccl_recv(ncclComm_t_ comm, void *buf, size_t count, NcclType t, int peer, cudaStream_t)
{ DEV(buf); if (count) p2p(comm, peer, grid_scope_index(comm->g, comm->scope), false, nullptr, buf, count * tsize(t)); }
void nccl_allgather(ncclComm_t_ comm, const void *send, void *recv, size_t sendcount, NcclType t, cudaStream_t)
{
    DEV(send); DEV(recv);
    const size_t len = sendcount * tsize(t);
    if (grid_scope_size(comm->g, comm->scope) <= 1) { if (recv != send) memmove(recv, send, len); return; }
    std::vector<char> tmp((const char *)send, (const char *)send + len);
    grid_allgather(comm->g, comm->scope, tmp.data(), recv, len);
}
template <typename T, typename F>
static void allreduce(ncclComm_t_ comm, const void *send, void *recv, size_t count, F f)
{
    DEV(send); DEV(recv);
    const int np = grid_scope_size(comm->g, comm->scope);
    std::vector<T> mine((const T *)send, (const T *)send + count);
    if (np > 1) {
        std::vector<T> all(count * (size_t)np);
        grid_allgather(comm->g, comm->scope, mine.data(), all.data(), count * sizeof(T));
        for (size_t e = 0; e < count; ++e) { T v = all[e]; for (int p = 1; p < np; ++p) v = f(v, all[(size_t)p * count + e]); mine[e] = v; }
    }
    memcpy(recv, mine.data(), count * sizeof(T));
}
void nccl_allreduce_min_i32(ncclComm_t_ c, const void *s, void *r, size_t n, cudaStream_t) { allreduce<int>(c, s, r, n, [](int a, int b) { return a < b ? a : b; }); }
void nccl_allreduce_sum_f64(ncclComm_t_ c, const void *s, void *r, size_t n, cudaStream_t) { allreduce<double>(c, s, r, n, [](double a, double b) { return a + b; }); }
void nccl_allreduce_max_f64(ncclComm_t_ c, const void *s, void *r, size_t n, cudaStream_t) { allreduce<double>(c, s, r, n, [](double a, double b) { return a > b ? a : b; }); }
void nccl_alltoallv(ncclComm_t_ comm, int np, int me, const void *send, const size_t *scount, const size_t *sdispl, void *recv,
                    const size_t *rcount, const size_t *rdispl, cudaStream_t)
{
    DEV(send); DEV(recv);
    // every rank publishes its counts / displacements and its whole send buffer (padded to the largest); each picks its parts
    size_t mytot = 0; for (int p = 0; p < np; ++p) if (sdispl[p] + scount[p] > mytot) mytot = sdispl[p] + scount[p];
    std::vector<size_t> meta((size_t)2 * np + 1), allmeta(((size_t)2 * np + 1) * np);
    for (int p = 0; p < np; ++p) { meta[(size_t)p] = scount[p]; meta[(size_t)np + p] = sdispl[p]; }
    meta[(size_t)2 * np] = mytot;
    if (np > 1) grid_allgather(comm->g, comm->scope, meta.data(), allmeta.data(), meta.size() * sizeof(size_t)); else allmeta = meta;
    size_t maxtot = 1; for (int p = 0; p < np; ++p) if (allmeta[(size_t)p * (2 * np + 1) + 2 * np] > maxtot) maxtot = allmeta[(size_t)p * (2 * np + 1) + 2 * np];
    std::vector<char> mine(maxtot, 0), all(maxtot * (size_t)np);
    if (mytot) memcpy(mine.data(), send, mytot);
    if (np > 1) grid_allgather(comm->g, comm->scope, mine.data(), all.data(), maxtot); else all = mine;
    for (int p = 0; p < np; ++p) {
        const size_t cnt = allmeta[(size_t)p * (2 * np + 1) + me], dsp = allmeta[(size_t)p * (2 * np + 1) + np + me];
        if (cnt != rcount[p]) fatal("emulated alltoallv: rank %d sends %zu bytes to %d which expects %zu", p, cnt, me, rcount[p]);
        if (cnt) memcpy((char *)recv + rdispl[p], all.data() + (size_t)p * maxtot + dsp, cnt);
    }
}
const char *nccl_version_string() { return "emulated"; }

}  // namespace slb

extern "C" {
int slb200_has_cuda(void) { return 0; }       // the emulation library never claims a GPU
int slb200_device(void) { return -1; }
int slb200_is_emulation(void) { return 1; }
// hooks of the GPU-only test drivers (testhooks.cu) that scalapack_b200/api.py resolves at load time: present, never callable here
#define NOT_EMULATED(name) double name(void) { slb::fatal(#name " is not part of the host-logic emulation"); }
NOT_EMULATED(slb200_test_gemm) NOT_EMULATED(slb200_test_panel)
}
