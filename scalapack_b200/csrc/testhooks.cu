// testhooks.cu -- C-ABI hooks that run ONE kernel of the LU path on caller-provided arrays so tests/ can
// check each kernel against the CPU oracle in isolation (pattern of the reference's PBLAS testers:
// per-kernel serial recompute, PBLAS/TESTING/pdblas3tst.f).  Arrays may be host (staged) or device.
#include "common.h"
#include "kernels.cuh"
#include "lu.h"

namespace slb {

static bool dev_ptr(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}
struct Tmp {
    void *d = nullptr, *h = nullptr; size_t bytes = 0; bool staged = false;
    Tmp(const void *p, size_t b) : bytes(b)
    {
        if (b == 0 || dev_ptr(p)) { d = const_cast<void *>(p); return; }
        staged = true; h = const_cast<void *>(p);
        SLB_CUDA(cudaMalloc(&d, b)); SLB_CUDA(cudaMemcpy(d, h, b, cudaMemcpyHostToDevice));
    }
    void back() { if (staged) SLB_CUDA(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost)); }
    ~Tmp() { if (staged && d) cudaFree(d); }
};

}  // namespace slb
using namespace slb;

extern "C" {

// C[MxN] -= A[MxK] B[KxN]; is_complex selects the zgemm kernel.  reps > 1 returns the mean ms per launch.
double slb200_test_gemm(int64_t M, int64_t N, int K, const void *A, int64_t lda, const void *B, int64_t ldb, void *C, int64_t ldc,
                        int is_complex, int reps)
{
    Runtime &r = rt();
    size_t es = is_complex ? 16 : 8;
    Tmp a(A, (size_t)lda * K * es), b(B, (size_t)ldb * N * es), c(C, (size_t)ldc * N * es);
    cudaEvent_t e0, e1; SLB_CUDA(cudaEventCreate(&e0)); SLB_CUDA(cudaEventCreate(&e1));
    SLB_CUDA(cudaEventRecord(e0, r.s_main));
    for (int i = 0; i < (reps > 0 ? reps : 1); ++i) {
        if (is_complex) launch_zgemm_minus(M, N, K, (const zcomplex *)a.d, lda, (const zcomplex *)b.d, ldb, (zcomplex *)c.d, ldc, r.s_main, (int)opt("gemm_test_chunk", 0), GEMM_MAIN);
        else launch_dgemm_minus(M, N, K, (const double *)a.d, lda, (const double *)b.d, ldb, (double *)c.d, ldc, r.s_main, (int)opt("gemm_test_chunk", 0), GEMM_MAIN);
    }
    SLB_CUDA(cudaEventRecord(e1, r.s_main));
    SLB_CUDA(cudaStreamSynchronize(r.s_main));
    float ms; SLB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    c.back();
    return ms / (reps > 0 ? reps : 1);
}

void slb200_test_trsm(int jb, int64_t n, const void *L, int64_t ldl, void *B, int64_t ldb, int is_complex)
{
    Runtime &r = rt();
    size_t es = is_complex ? 16 : 8;
    Tmp l(L, (size_t)ldl * jb * es), b(B, (size_t)ldb * n * es);
    if (is_complex) launch_ztrsm_llnu(jb, n, (const zcomplex *)l.d, ldl, (zcomplex *)b.d, ldb, r.s_main);
    else launch_dtrsm_llnu(jb, n, (const double *)l.d, ldl, (double *)b.d, ldb, r.s_main);
    SLB_CUDA(cudaStreamSynchronize(r.s_main));
    b.back();
}

// Panel factorisation of a single-segment m x jb panel (global row = row index).  ipiv: jb ints (1-based rows).
double slb200_test_panel(int m, int jb, void *W, int64_t ldw, int *ipiv, int *info, int is_complex)
{
    Runtime &r = rt();
    size_t es = is_complex ? 16 : 8;
    Tmp w(W, (size_t)ldw * jb * es);
    int *dpiv = (int *)workspace("t_piv", (size_t)(jb + 1) * sizeof(int));
    void *work = workspace("lu_panelwork", panel_work_bytes(jb > 512 ? jb : 512), true);
    SLB_CUDA(cudaMemsetAsync(dpiv, 0, (size_t)(jb + 1) * sizeof(int), r.s_main));
    PanelRowMap map{}; map.g0 = 0; map.nb = m > 0 ? m : 1; map.nprow = 1; map.rsrc = 0;
    cudaEvent_t e0, e1; SLB_CUDA(cudaEventCreate(&e0)); SLB_CUDA(cudaEventCreate(&e1));
    SLB_CUDA(cudaEventRecord(e0, r.s_main));
    if (is_complex) launch_zpanel(m, jb, (zcomplex *)w.d, ldw, map, dpiv, dpiv + jb, 0, work, r.s_main);
    else launch_dpanel(m, jb, (double *)w.d, ldw, map, dpiv, dpiv + jb, 0, work, r.s_main);
    SLB_CUDA(cudaEventRecord(e1, r.s_main));
    SLB_CUDA(cudaStreamSynchronize(r.s_main));
    float ms; SLB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    std::vector<int> h((size_t)jb + 1);
    SLB_CUDA(cudaMemcpy(h.data(), dpiv, (size_t)(jb + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    for (int j = 0; j < jb; ++j) ipiv[j] = h[j];
    *info = h[jb];
    w.back();
    return ms;
}

// copies the panel kernel's debug timestamps (SLB200_PANEL_DEBUG=1): out[2][4096][8] u64
void slb200_test_panel_dbg(unsigned long long *out)
{
    void *d = workspace("panel_dbg", 2 * 4096 * 8 * 8, true);
    SLB_CUDA(cudaMemcpy(out, d, 2 * 4096 * 8 * 8, cudaMemcpyDeviceToHost));
}

// Row interchanges of one block on an m x n real matrix (single process row): rows j0+t <-> ipiv[t]-1.
void slb200_test_laswp(int m, int64_t n, double *A, int64_t lda, int j0, int jb, const int *ipiv_blk)
{
    Runtime &r = rt(); cudaStream_t s = r.s_main;
    Tmp a(A, (size_t)lda * n * 8);
    int *dp = (int *)workspace("t_piv", (size_t)(jb + 1) * sizeof(int));
    SLB_CUDA(cudaMemcpyAsync(dp, ipiv_blk, (size_t)jb * sizeof(int), cudaMemcpyHostToDevice, s));
    int *pm = (int *)workspace("lu_plan", (size_t)3 * jb * sizeof(int));
    SwapPlan plan{ pm, pm + jb, pm + 2 * jb };
    double *U = (double *)workspace("lu_U", (size_t)jb * n * 8), *O = (double *)workspace("lu_O", (size_t)jb * n * 8);
    RowDist rd{ m > 0 ? m : 1, 1, 0, 0, 0 };
    launch_swap_plan(j0, jb, dp, plan, s);
    launch_swap_pack<double>(jb, j0, plan, rd, (double *)a.d, lda, 0, n, U, jb, O, jb, s);
    launch_swap_unpack_out<double>(jb, plan, rd, (double *)a.d, lda, 0, n, O, jb, s);
    launch_copy2d<double>(jb, n, U, jb, (double *)a.d + j0, lda, s);
    SLB_CUDA(cudaStreamSynchronize(s));
    a.back();
}

// standalone timing of the PDGETRS fast-path kernels (solve_fast.cu); A: device pointer to N x N factors
double slb200_test_solve_probe(int which, int nb, int64_t nr, const double *A, int64_t lld, int N, int reps)
{ return solve_fast_probe(which, nb, nr, A, lld, N, reps); }

}  // extern "C"
