// swap.cu -- row interchanges of one LU block step (replaces PDLASWP = jb x PDSWAP, SRC/pdlaswp.f:163-182,
// PBLAS/SRC/pdswap_.c:448-534) as batched gather / scatter kernels.
//
// The reference applies the jb interchanges (j0+t) <-> ipiv[t] one at a time, each a strided row swap and,
// across process rows, a send/recv pair.  Here the NET permutation of the block is computed once on the
// device (swap_plan_kernel: every destination row traces its source backwards through the jb swaps, all
// destinations in parallel), then all columns move in one pass:
//   pack    : gather the rows that end in the top block into U (jb x ncols, contiguous), and the original
//             top rows that end outside into O;
//   (multi-GPU: U / O are exchanged inside the process column by NCCL, lu.cu)
//   unpack  : scatter O into the outside rows.
// Threads run along t (rows of the block): U/O accesses are coalesced, A accesses are 8-byte gathers.
#include "kernels.cuh"
#include "devmath.cuh"
#include "common.h"

namespace slb {

namespace {

__global__ void swap_plan_kernel(int j0, int jb, const int *__restrict__ ipiv_blk, int *__restrict__ top_src,
                                 int *__restrict__ out_dst, int *__restrict__ out_src)
{
    extern __shared__ int piv[];      // 0-based global pivot rows
    for (int t = threadIdx.x; t < jb; t += blockDim.x) piv[t] = ipiv_blk[t] - 1;
    __syncthreads();
    for (int t = threadIdx.x; t < jb; t += blockDim.x) {
        // source of top row j0+t
        int pos = j0 + t;
        for (int s = jb - 1; s >= 0; --s) {
            int r = j0 + s, p = piv[s];
            if (pos == r) pos = p; else if (pos == p) pos = r;
        }
        top_src[t] = pos;
        // outside destination represented by the first t that names it
        int p = piv[t];
        int dst = -1, src = 0;
        if (p >= j0 + jb) {
            bool first = true;
            for (int s = 0; s < t; ++s) if (piv[s] == p) { first = false; break; }
            if (first) {
                dst = p;
                int q = p;
                for (int s = jb - 1; s >= 0; --s) {
                    int r = j0 + s, pp = piv[s];
                    if (q == r) q = pp; else if (q == pp) q = r;
                }
                src = q - j0;     // always inside the top block (swap plan)
            }
        }
        out_dst[t] = dst;
        out_src[t] = src;
    }
}

__device__ __forceinline__ int row_owner(const RowDist &rd, int g) { return (rd.rsrc + g / rd.nb) % rd.nprow; }
__device__ __forceinline__ int64_t row_local(const RowDist &rd, int g)
{ return (int64_t)rd.nb * (g / ((int64_t)rd.nb * rd.nprow)) + g % rd.nb - rd.shift; }

constexpr int SWAP_COLS = 8;      // columns per block

template <typename T>
__global__ void __launch_bounds__(256)
swap_pack_kernel(int jb, int j0, const int *__restrict__ top_src, const int *__restrict__ out_dst,
                 const int *__restrict__ out_src, RowDist rd, const T *__restrict__ A, int64_t lda, int64_t c0, int64_t c1,
                 T *__restrict__ Ubuf, int64_t ldu, T *__restrict__ Obuf, int64_t ldo)
{
    bool own_top = row_owner(rd, j0) == rd.myrow;
    for (int64_t cb = c0 + (int64_t)blockIdx.x * SWAP_COLS; cb < c1; cb += (int64_t)gridDim.x * SWAP_COLS) {
    int nc = (int)min((int64_t)SWAP_COLS, c1 - cb);
    for (int t = threadIdx.x; t < jb; t += blockDim.x) {
        int src = top_src[t];
        if (row_owner(rd, src) == rd.myrow) {
            const T *ap = A + row_local(rd, src) + cb * lda;
            T v[SWAP_COLS];
#pragma unroll
            for (int c = 0; c < SWAP_COLS; ++c) if (c < nc) v[c] = ap[(int64_t)c * lda];
#pragma unroll
            for (int c = 0; c < SWAP_COLS; ++c) if (c < nc) Ubuf[t + (cb - c0 + c) * ldu] = v[c];
        }
        if (own_top && Obuf != nullptr) {
            int d = out_dst[t];
            if (d >= 0) {
                const T *ap = A + row_local(rd, j0 + out_src[t]) + cb * lda;
                T v[SWAP_COLS];
#pragma unroll
                for (int c = 0; c < SWAP_COLS; ++c) if (c < nc) v[c] = ap[(int64_t)c * lda];
#pragma unroll
                for (int c = 0; c < SWAP_COLS; ++c) if (c < nc) Obuf[t + (cb - c0 + c) * ldo] = v[c];
            }
        }
    }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
swap_unpack_out_kernel(int jb, const int *__restrict__ out_dst, RowDist rd, T *__restrict__ A, int64_t lda, int64_t c0,
                       int64_t c1, const T *__restrict__ Obuf, int64_t ldo)
{
    for (int64_t cb = c0 + (int64_t)blockIdx.x * SWAP_COLS; cb < c1; cb += (int64_t)gridDim.x * SWAP_COLS) {
    int nc = (int)min((int64_t)SWAP_COLS, c1 - cb);
    for (int t = threadIdx.x; t < jb; t += blockDim.x) {
        int d = out_dst[t];
        if (d < 0 || row_owner(rd, d) != rd.myrow) continue;
        T *ap = A + row_local(rd, d) + cb * lda;
        T v[SWAP_COLS];
#pragma unroll
        for (int c = 0; c < SWAP_COLS; ++c) if (c < nc) v[c] = Obuf[t + (cb - c0 + c) * ldo];
#pragma unroll
        for (int c = 0; c < SWAP_COLS; ++c) if (c < nc) ap[(int64_t)c * lda] = v[c];
    }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
swap_select_kernel(int jb, const int *__restrict__ top_src, RowDist rd, const T *__restrict__ Call, int64_t ldc,
                   int64_t stride_p, int64_t ncols, T *__restrict__ U, int64_t ldu)
{
    int64_t cb = (int64_t)blockIdx.x * SWAP_COLS;
    int nc = (int)min((int64_t)SWAP_COLS, ncols - cb);
    for (int t = threadIdx.x; t < jb; t += blockDim.x) {
        int p = row_owner(rd, top_src[t]);
        const T *sp = Call + (int64_t)p * stride_p + t + cb * ldc;
#pragma unroll
        for (int c = 0; c < SWAP_COLS; ++c) if (c < nc) U[t + (cb + c) * ldu] = sp[(int64_t)c * ldc];
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
copy2d_kernel(int64_t rows, int64_t cols, const T *__restrict__ src, int64_t lds, T *__restrict__ dst, int64_t ldd)
{
    int64_t c = blockIdx.y;
    for (; c < cols; c += gridDim.y)
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x)
            dst[i + c * ldd] = src[i + c * lds];
}

template <typename T>
__global__ void __launch_bounds__(256)
rows_bc_kernel(int64_t rows, int cols, T *__restrict__ L, int64_t ldl, int64_t l0, T *__restrict__ G, int64_t ldg, int64_t gshift,
               int nb, int nprow, int prow_rel, int to_global)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    int64_t l = l0 + i;
    int64_t g = ((l / nb) * nprow + prow_rel) * nb + l % nb - gshift;
    for (int c = blockIdx.y; c < cols; c += gridDim.y) {
        if (to_global) G[g + (int64_t)c * ldg] = L[i + (int64_t)c * ldl];
        else L[i + (int64_t)c * ldl] = G[g + (int64_t)c * ldg];
    }
}

}  // namespace

template <typename T>
void launch_rows_bc(int64_t rows, int cols, T *L, int64_t ldl, int64_t l0, T *G, int64_t ldg, int64_t gshift, int nb, int nprow,
                    int prow_rel, int to_global, cudaStream_t s)
{
    if (rows <= 0 || cols <= 0) return;
    dim3 grid((unsigned)((rows + 255) / 256), (unsigned)(cols < 64 ? cols : 64));
    rows_bc_kernel<T><<<grid, 256, 0, s>>>(rows, cols, L, ldl, l0, G, ldg, gshift, nb, nprow, prow_rel, to_global);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}

// The gather / scatter kernels are memory-latency bound: a few hundred CTAs keep the HBM queues full, and a capped
// grid leaves the other SMs to the trailing update running underneath (lu.cu pipelines the two).  SLB200_SWAP_GRID=0: uncapped.
static thread_local int g_grid_override = -1;      // >= 0: replaces the swap_grid option (0 = uncapped)
void swap_grid_override(int cap) { g_grid_override = cap; }
static unsigned capped_grid(int64_t groups)
{
    const int64_t cap = g_grid_override >= 0 ? g_grid_override : opt("swap_grid", 48);
    return (unsigned)((cap > 0 && groups > cap) ? cap : groups);
}

void launch_swap_plan(int j0, int jb, const int *ipiv_blk, SwapPlan plan, cudaStream_t s)
{
    if (jb <= 0) return;
    int threads = jb < 1024 ? ((jb + 31) / 32) * 32 : 1024;
    swap_plan_kernel<<<1, threads, (size_t)jb * sizeof(int), s>>>(j0, jb, ipiv_blk, plan.top_src, plan.out_dst, plan.out_src);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}

template <typename T>
void launch_swap_pack(int jb, int j0, SwapPlan plan, RowDist rd, const T *A, int64_t lda, int64_t c0, int64_t c1,
                      T *Ubuf, int64_t ldu, T *Obuf, int64_t ldo, cudaStream_t s)
{
    if (c1 <= c0 || jb <= 0) return;
    unsigned grid = capped_grid((c1 - c0 + SWAP_COLS - 1) / SWAP_COLS);
    swap_pack_kernel<T><<<grid, 256, 0, s>>>(jb, j0, plan.top_src, plan.out_dst, plan.out_src, rd, A, lda, c0, c1, Ubuf, ldu, Obuf, ldo);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}
template <typename T>
void launch_swap_unpack_out(int jb, SwapPlan plan, RowDist rd, T *A, int64_t lda, int64_t c0, int64_t c1, const T *Obuf,
                            int64_t ldo, cudaStream_t s)
{
    if (c1 <= c0 || jb <= 0) return;
    unsigned grid = capped_grid((c1 - c0 + SWAP_COLS - 1) / SWAP_COLS);
    swap_unpack_out_kernel<T><<<grid, 256, 0, s>>>(jb, plan.out_dst, rd, A, lda, c0, c1, Obuf, ldo);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}
template <typename T>
void launch_swap_select(int jb, SwapPlan plan, RowDist rd, const T *Call, int64_t ldc, int64_t stride_p, int64_t ncols, T *U,
                        int64_t ldu, cudaStream_t s)
{
    if (ncols <= 0 || jb <= 0) return;
    unsigned grid = (unsigned)((ncols + SWAP_COLS - 1) / SWAP_COLS);
    swap_select_kernel<T><<<grid, 256, 0, s>>>(jb, plan.top_src, rd, Call, ldc, stride_p, ncols, U, ldu);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}
template <typename T>
void launch_copy2d(int64_t rows, int64_t cols, const T *src, int64_t lds, T *dst, int64_t ldd, cudaStream_t s)
{
    if (rows <= 0 || cols <= 0) return;
    // capped grid (both loops of the kernel are grid-stride): a copy needs few SMs to reach its bandwidth and must not
    // flush the trailing update running underneath off the GPU
    unsigned gx = (unsigned)min((int64_t)64, (rows + 255) / 256);
    unsigned gy = (unsigned)min((int64_t)max(1, 1024 / (int)gx), cols);
    copy2d_kernel<T><<<dim3(gx, gy), 256, 0, s>>>(rows, cols, src, lds, dst, ldd);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}

#define INST(T)                                                                                                        \
    template void launch_swap_pack<T>(int, int, SwapPlan, RowDist, const T *, int64_t, int64_t, int64_t, T *, int64_t, \
                                      T *, int64_t, cudaStream_t);                                                    \
    template void launch_swap_unpack_out<T>(int, SwapPlan, RowDist, T *, int64_t, int64_t, int64_t, const T *, int64_t, \
                                            cudaStream_t);                                                            \
    template void launch_swap_select<T>(int, SwapPlan, RowDist, const T *, int64_t, int64_t, int64_t, T *, int64_t,    \
                                        cudaStream_t);                                                                \
    template void launch_copy2d<T>(int64_t, int64_t, const T *, int64_t, T *, int64_t, cudaStream_t);                  \
    template void launch_rows_bc<T>(int64_t, int, T *, int64_t, int64_t, T *, int64_t, int64_t, int, int, int, int, cudaStream_t);
INST(double)
INST(zcomplex)
template void launch_copy2d<int>(int64_t, int64_t, const int *, int64_t, int *, int64_t, cudaStream_t);
#undef INST

}  // namespace slb
