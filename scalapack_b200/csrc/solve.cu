// solve.cu -- PDGETRS 'N' (SRC/pdgetrs.f:255-266): x = U^-1 L^-1 P b on the P x Q grid.
//
// The right-hand sides are tiny next to A (N x NRHS vs N x N), so they are replicated: every GPU holds the
// whole N x NRHS block X in global row order.  The interchanges of PDLAPIV/PDLAPV2 (SRC/pdlapv2.f:195-242:
// N sequential PDSWAPs) collapse into ONE row gather with the net permutation.  Each triangular solve then
// walks the diagonal blocks: the process row owning block k sums its partial products (row all-reduce), the
// diagonal owner solves the nb x nb block (warp-shuffle substitution) and broadcasts x_k, and the process
// column owning block column k folds x_k into its partial products with an HBM-bound GEMV over its rows.
// L and U are each read exactly once.
#include "common.h"
#include "kernels.cuh"
#include "ncclw.h"
#include "lu.h"
#include "devmath.cuh"

namespace slb {

namespace {

// Xg[ig + jg*N] = B[il + jl*lldb] for my local entries (others untouched)
template <typename T>
__global__ void rhs_scatter_kernel(int64_t mlocB, int64_t nlocB, int nb, int nbb, int P, int Q, int myrow_rel, int mycol_rel,
                                   const T *__restrict__ B, int64_t lldb, T *__restrict__ Xg, int64_t N, int to_global)
{
    int64_t il = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t jl = blockIdx.y;
    if (il >= mlocB || jl >= nlocB) return;
    int64_t ig = ((il / nb) * P + myrow_rel) * nb + il % nb;
    int64_t jg = ((jl / nbb) * Q + mycol_rel) * nbb + jl % nbb;
    if (to_global) Xg[ig + jg * N] = B[il + jl * lldb];
    else const_cast<T *>(B)[il + jl * lldb] = Xg[ig + jg * N];
}

template <typename T>
__global__ void add_block_kernel(int kb, int nrhs, const T *__restrict__ xg, int64_t ldx, const T *__restrict__ red, int64_t ldr,
                                 T *__restrict__ out, int64_t ldo)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (i >= kb || c >= nrhs) return;
    T a = xg[i + (int64_t)c * ldx], b = red[i + (int64_t)c * ldr];
    out[i + (int64_t)c * ldo] = t_add(a, b);
}

}  // namespace

template <typename T>
int getrs_device(Grid *g, int N, int nrhs, const T *A, int64_t lld, int nb, int rsrc, int csrc, const int *ipiv_glob_host,
                 T *B, int64_t lldb, int nbb, int csrcb)
{
    Runtime &r = rt();
    cudaStream_t s = r.s_main;
    const int P = g->nprow, Q = g->npcol, myrow = g->myrow, mycol = g->mycol;
    const bool multi = P * Q > 1;
    if (multi && !g->nccl) g->nccl = nccl_create(g);
    NcclComms *nc = g->nccl;
    const int64_t mloc = numroc(N, nb, myrow, rsrc, P);
    const int64_t nlocB = numroc(nrhs, nbb, mycol, csrcb, Q);
    const int nelem = sizeof(T) / sizeof(double);

    // ---- net permutation of the N forward interchanges (host, O(N)) ----
    std::vector<int> perm((size_t)N);
    for (int i = 0; i < N; ++i) perm[i] = i;
    for (int i = 0; i < N; ++i) { int p = ipiv_glob_host[i] - 1; if (p != i) { int t = perm[i]; perm[i] = perm[p]; perm[p] = t; } }
    int *perm_dev = (int *)workspace("rs_perm", (size_t)N * sizeof(int));
    SLB_CUDA(cudaMemcpyAsync(perm_dev, perm.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, s));

    T *X0 = (T *)workspace("rs_X0", (size_t)N * nrhs * sizeof(T));
    T *Xg = (T *)workspace("rs_Xg", (size_t)N * nrhs * sizeof(T));
    T *acc = (T *)workspace("rs_acc", (size_t)(mloc > 0 ? mloc : 1) * nrhs * sizeof(T));
    T *red = (T *)workspace("rs_red", (size_t)2 * nb * nrhs * sizeof(T));
    T *tmp = red + (size_t)nb * nrhs;

    cudaEvent_t ev0, ev1; SLB_CUDA(cudaEventCreate(&ev0)); SLB_CUDA(cudaEventCreate(&ev1));
    SLB_CUDA(cudaEventRecord(ev0, s));

    // ---- replicate B in global order ----
    SLB_CUDA(cudaMemsetAsync(X0, 0, (size_t)N * nrhs * sizeof(T), s));
    if (mloc > 0 && nlocB > 0) {
        dim3 grid((unsigned)((mloc + 255) / 256), (unsigned)nlocB);
        rhs_scatter_kernel<T><<<grid, 256, 0, s>>>(mloc, nlocB, nb, nbb, P, Q, (P + myrow - rsrc) % P, (Q + mycol - csrcb) % Q, B, lldb, X0, N, 1);
        SLB_CUDA(cudaGetLastError()); counter_add("kernel_launches", 1);
    }
    if (multi) nccl_allreduce_sum_f64(nc->all, X0, X0, (size_t)N * nrhs * nelem, s);
    // ---- P b ----
    launch_gather_rows<T>(N, perm_dev, X0, N, Xg, N, nrhs, s);

    const int nblk = (N + nb - 1) / nb;
    for (int pass = 0; pass < 2; ++pass) {          // 0: L (unit lower, forward)   1: U (upper, backward)
        SLB_CUDA(cudaMemsetAsync(acc, 0, (size_t)(mloc > 0 ? mloc : 1) * nrhs * sizeof(T), s));
        for (int q = 0; q < nblk; ++q) {
            const int k = pass == 0 ? q : nblk - 1 - q;
            const int j0 = k * nb, jb = (N - j0) < nb ? (N - j0) : nb;
            const int pr = (rsrc + k) % P, pc = (csrc + k) % Q;
            const int64_t lr0 = numroc(j0, nb, myrow, rsrc, P);           // my local rows above block k
            const int64_t lc0 = numroc(j0, nb, mycol, csrc, Q);
            T *xk = Xg + j0;                                              // ld = N
            if (myrow == pr) {
                const T *part = acc + lr0;                                // ld = mloc
                if (Q > 1) {
                    launch_copy2d<T>(jb, nrhs, part, mloc, red, jb, s);
                    nccl_allreduce_sum_f64(nc->row, red, red, (size_t)jb * nrhs * nelem, s);
                    part = red;
                }
                if (mycol == pc) {
                    dim3 grid((unsigned)((jb + 127) / 128), (unsigned)nrhs);
                    add_block_kernel<T><<<grid, 128, 0, s>>>(jb, nrhs, xk, N, part, Q > 1 ? jb : mloc, tmp, jb);
                    SLB_CUDA(cudaGetLastError()); counter_add("kernel_launches", 1);
                    launch_trsv_block<T>(jb, A + lr0 + lc0 * lld, lld, tmp, jb, nrhs, pass, s);
                }
            }
            if (multi) {
                int root = pr * Q + pc;
                nccl_bcast(nc->all, tmp, (size_t)jb * nrhs * sizeof(T), NT_U8, root, s);
            }
            launch_copy2d<T>(jb, nrhs, tmp, jb, xk, N, s);
            if (mycol == pc) {
                if (pass == 0) {
                    const int64_t rbeg = numroc(j0 + jb, nb, myrow, rsrc, P);     // my rows below block k
                    launch_gemv_minus<T>(mloc - rbeg, jb, A + rbeg + lc0 * lld, lld, tmp, jb, acc + rbeg, mloc, nrhs, s);
                } else {
                    launch_gemv_minus<T>(lr0, jb, A + lc0 * lld, lld, tmp, jb, acc, mloc, nrhs, s);   // my rows above
                }
            }
        }
    }
    // ---- back into the caller's block-cyclic B ----
    if (mloc > 0 && nlocB > 0) {
        dim3 grid((unsigned)((mloc + 255) / 256), (unsigned)nlocB);
        rhs_scatter_kernel<T><<<grid, 256, 0, s>>>(mloc, nlocB, nb, nbb, P, Q, (P + myrow - rsrc) % P, (Q + mycol - csrcb) % Q, B, lldb, Xg, N, 0);
        SLB_CUDA(cudaGetLastError()); counter_add("kernel_launches", 1);
    }
    SLB_CUDA(cudaEventRecord(ev1, s));
    SLB_CUDA(cudaStreamSynchronize(s));
    float ms = 0; SLB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    g_last_lu.solve_ms = ms;
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    return 0;
}

template int getrs_device<double>(Grid *, int, int, const double *, int64_t, int, int, int, const int *, double *, int64_t, int, int);
template int getrs_device<zcomplex>(Grid *, int, int, const zcomplex *, int64_t, int, int, int, const int *, zcomplex *, int64_t, int, int);

}  // namespace slb
