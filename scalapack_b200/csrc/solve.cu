// solve.cu -- PDGETRS (SRC/pdgetrs.f:255-284): x = U^-1 L^-1 P b ('N') or x = P^T L^-T U^-T b ('T', 'C': ^H) on the P x Q grid.
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
#include "launch.h"

namespace slb {

namespace {

// Xg[ig + (jg - jb0)*N] = B[il + jl*lldb] for my local entries of sub(B) (global columns [jb0, jb0 + nrhs) of B); others untouched
template <typename T>
__global__ void rhs_scatter_kernel(int64_t mlocB, int64_t nlocB, int nb, int nbb, int P, int Q, int myrow_rel, int mycol_rel,
                                   const T *__restrict__ B, int64_t lldb, T *__restrict__ Xg, int64_t N, int jb0, int nrhs, int to_global)
{
    int64_t il = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (il >= mlocB) return;
    int64_t ig = ((il / nb) * P + myrow_rel) * nb + il % nb;
    for (int64_t jl = blockIdx.y; jl < nlocB; jl += gridDim.y) {
        int64_t jg = ((jl / nbb) * Q + mycol_rel) * nbb + jl % nbb - jb0;
        if (jg < 0 || jg >= nrhs) continue;
        if (to_global) Xg[ig + jg * N] = B[il + jl * lldb];
        else const_cast<T *>(B)[il + jl * lldb] = Xg[ig + jg * N];
    }
}

template <typename T>
__global__ void add_block_kernel(int kb, int nrhs, const T *__restrict__ xg, int64_t ldx, const T *__restrict__ red, int64_t ldr,
                                 T *__restrict__ out, int64_t ldo)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kb) return;
    for (int c = blockIdx.y; c < nrhs; c += gridDim.y) {
        T a = xg[i + (int64_t)c * ldx], b = red[i + (int64_t)c * ldr];
        out[i + (int64_t)c * ldo] = t_add(a, b);
    }
}

}  // namespace

template <typename T>
int getrs_device(Grid *g, char trans, int N, int nrhs, const T *A, int64_t lld, int nb, int rsrc, int csrc, const int *ipiv_glob_host,
                 T *B, int64_t lldb, int nbb, int csrcb, int jb0, int64_t nlocB_all, const T *Xrep_in, T *Xrep_out)
{
    Runtime &r = rt();
    cudaStream_t s = r.s_main;
    const int P = g->nprow, Q = g->npcol, myrow = g->myrow, mycol = g->mycol;
    const bool multi = P * Q > 1;
    if (multi && !g->nccl) g->nccl = nccl_create(g);
    NcclComms *nc = g->nccl;
    const int64_t mloc = numroc(N, nb, myrow, rsrc, P), nloc = numroc(N, nb, mycol, csrc, Q);
    const int nelem = sizeof(T) / sizeof(double);
    const bool tr = trans != 'N', cj = trans == 'C';
    const unsigned gy = (unsigned)(nrhs < 65535 ? nrhs : 65535);

    // ---- net permutation of the N forward interchanges (host, O(N)): row i of P b is row perm[i] of b ----
    std::vector<int> perm((size_t)N);
    for (int i = 0; i < N; ++i) perm[i] = i;
    for (int i = 0; i < N; ++i) { int p = ipiv_glob_host[i] - 1; if (p != i) { int t = perm[i]; perm[i] = perm[p]; perm[p] = t; } }
    int *perm_dev = (int *)workspace("rs_perm", (size_t)N * sizeof(int));
    SLB_CUDA(cudaMemcpyAsync(perm_dev, perm.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, s));

    const int64_t nacc = tr ? nloc : mloc;               // partial products: by local row ('N') or by local column ('T','C')
    T *X0 = (T *)workspace("rs_X0", (size_t)N * nrhs * sizeof(T));
    T *Xg = (T *)workspace("rs_Xg", (size_t)N * nrhs * sizeof(T));
    T *acc = (T *)workspace("rs_acc", (size_t)(nacc > 0 ? nacc : 1) * nrhs * sizeof(T));
    T *red = (T *)workspace("rs_red", (size_t)2 * nb * nrhs * sizeof(T));
    T *tmp = red + (size_t)nb * nrhs;
    const int64_t lda_acc = nacc > 0 ? nacc : 1;

    cudaEvent_t ev0, ev1; SLB_CUDA(cudaEventCreate(&ev0)); SLB_CUDA(cudaEventCreate(&ev1));
    SLB_CUDA(cudaEventRecord(ev0, s));

    // ---- replicate sub(B) in global order ----
    const int myr_rel = (P + myrow - rsrc) % P, myc_relb = (Q + mycol - csrcb) % Q;
    unsigned gyb = (unsigned)(nlocB_all < 65535 ? (nlocB_all > 0 ? nlocB_all : 1) : 65535);
    if (Xrep_in) {
        // the caller already holds the right-hand sides replicated in global row order (PDGECON / PDGERFS work vectors)
        SLB_CUDA(cudaMemcpyAsync(X0, Xrep_in, (size_t)N * nrhs * sizeof(T), cudaMemcpyDeviceToDevice, s));
    } else {
    SLB_CUDA(cudaMemsetAsync(X0, 0, (size_t)N * nrhs * sizeof(T), s));
    if (mloc > 0 && nlocB_all > 0) {
        dim3 grid((unsigned)((mloc + 255) / 256), gyb);
        SLB_LAUNCH((rhs_scatter_kernel<T>), grid, 256, s, mloc, nlocB_all, nb, nbb, P, Q, myr_rel, myc_relb, (const T *)B, lldb, X0, (int64_t)N, jb0, nrhs, 1);
    }
    if (multi) nccl_allreduce_sum_f64(nc->all, X0, X0, (size_t)N * nrhs * nelem, s);
    }
    // ---- 'N': x = U^-1 L^-1 (P b) (pdgetrs.f:255-266);  'T','C': x = P^T (L^-T (U^-T b)) (pdgetrs.f:268-284) ----
    T *Xw = Xg;
    if (!tr) launch_gather_rows<T>(N, perm_dev, X0, N, Xg, N, nrhs, s);
    else Xw = X0;

    const int nblk = (N + nb - 1) / nb;
    bool fast = false;
    if constexpr (sizeof(T) == sizeof(double)) {
        if (getrs_fast_applies(P, Q, trans, nb, nrhs)) {
            getrs_fast_device(N, nrhs, reinterpret_cast<const double *>(A), lld, nb, reinterpret_cast<double *>(Xg));
            fast = true;
        }
    }
    for (int pass = 0; pass < 2 && !fast; ++pass) {
        // 'N': pass 0 = L forward, pass 1 = U backward.  'T'/'C': pass 0 = U^T forward, pass 1 = L^T backward.
        const bool use_u = tr ? pass == 0 : pass == 1;
        const bool fwd = pass == 0;
        SLB_CUDA(cudaMemsetAsync(acc, 0, (size_t)lda_acc * nrhs * sizeof(T), s));
        for (int q = 0; q < nblk; ++q) {
            const int k = fwd ? q : nblk - 1 - q;
            const int j0 = k * nb, jb = (N - j0) < nb ? (N - j0) : nb;
            const int pr = (rsrc + k) % P, pc = (csrc + k) % Q;
            const int64_t lr0 = numroc(j0, nb, myrow, rsrc, P);           // my local rows above block k
            const int64_t lc0 = numroc(j0, nb, mycol, csrc, Q);
            T *xk = Xw + j0;                                              // ld = N
            // the partial products of block k live on the process row ('N') / column ('T') that owns it
            const bool holds = tr ? mycol == pc : myrow == pr;
            if (holds) {
                const T *part = acc + (tr ? lc0 : lr0);
                int64_t ldp = lda_acc;
                const int nred = tr ? P : Q;
                if (nred > 1) {
                    launch_copy2d<T>(jb, nrhs, part, lda_acc, red, jb, s);
                    nccl_allreduce_sum_f64(tr ? nc->col : nc->row, red, red, (size_t)jb * nrhs * nelem, s);
                    part = red; ldp = jb;
                }
                if (myrow == pr && mycol == pc) {
                    dim3 grid((unsigned)((jb + 127) / 128), gy);
                    SLB_LAUNCH((add_block_kernel<T>), grid, 128, s, jb, nrhs, (const T *)xk, (int64_t)N, part, ldp, tmp, (int64_t)jb);
                    launch_trsv_block<T>(jb, A + lr0 + lc0 * lld, lld, tmp, jb, nrhs,
                                         (use_u ? TRSV_UPPER : 0) | (tr ? TRSV_TRANS : 0) | (cj ? TRSV_CONJ : 0), s);
                }
            }
            if (multi) {
                int root = pr * Q + pc;
                nccl_bcast(nc->all, tmp, (size_t)jb * nrhs * sizeof(T), NT_U8, root, s);
            }
            launch_copy2d<T>(jb, nrhs, tmp, jb, xk, N, s);
            if (!tr) {
                if (mycol == pc) {
                    if (fwd) {
                        const int64_t rbeg = numroc(j0 + jb, nb, myrow, rsrc, P);     // my rows below block k
                        launch_gemv_minus<T>(mloc - rbeg, jb, A + rbeg + lc0 * lld, lld, tmp, jb, acc + rbeg, lda_acc, nrhs, s);
                    } else {
                        launch_gemv_minus<T>(lr0, jb, A + lc0 * lld, lld, tmp, jb, acc, lda_acc, nrhs, s);   // my rows above
                    }
                }
            } else if (myrow == pr) {
                if (fwd) {                                                 // U^T: columns right of block k
                    const int64_t cbeg = numroc(j0 + jb, nb, mycol, csrc, Q);
                    launch_gemvt_minus<T>(jb, nloc - cbeg, A + lr0 + cbeg * lld, lld, tmp, jb, acc + cbeg, lda_acc, nrhs, cj, s);
                } else {                                                   // L^T: columns left of block k
                    launch_gemvt_minus<T>(jb, lc0, A + lr0, lld, tmp, jb, acc, lda_acc, nrhs, cj, s);
                }
            }
        }
    }
    if (tr) { launch_gather_rows<T>(N, perm_dev, X0, N, Xg, N, nrhs, s, true); }   // x = P^T z: row perm[i] of x is row i of z
    // ---- back into the caller's block-cyclic sub(B) (or to the caller's replicated copy) ----
    if (Xrep_out) SLB_CUDA(cudaMemcpyAsync(Xrep_out, Xg, (size_t)N * nrhs * sizeof(T), cudaMemcpyDeviceToDevice, s));
    else if (mloc > 0 && nlocB_all > 0) {
        dim3 grid((unsigned)((mloc + 255) / 256), gyb);
        SLB_LAUNCH((rhs_scatter_kernel<T>), grid, 256, s, mloc, nlocB_all, nb, nbb, P, Q, myr_rel, myc_relb, (const T *)B, lldb, Xg, (int64_t)N, jb0, nrhs, 0);
    }
    SLB_CUDA(cudaEventRecord(ev1, s));
    SLB_CUDA(cudaStreamSynchronize(s));
    float ms = 0; SLB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    g_last_lu.solve_ms = ms;
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    return 0;
}

template int getrs_device<double>(Grid *, char, int, int, const double *, int64_t, int, int, int, const int *, double *, int64_t, int, int, int, int64_t, const double *, double *);
template int getrs_device<zcomplex>(Grid *, char, int, int, const zcomplex *, int64_t, int, int, int, const int *, zcomplex *, int64_t, int, int, int, int64_t, const zcomplex *, zcomplex *);

}  // namespace slb
