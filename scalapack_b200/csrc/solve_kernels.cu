// solve_kernels.cu -- kernels of the triangular solves of PDGETRS (SRC/pdgetrs.f:255-266; the reference
// runs PDLAPIV + two PDTRSMs, PBLAS/SRC/PTOOLS/PB_CptrsmB.c:441-590 for few right-hand sides).
// HBM-bound: L and U are each streamed once (8 N^2 bytes in total for one right-hand side).
#include "kernels.cuh"
#include "devmath.cuh"
#include "common.h"

#include <algorithm>

namespace slb {

namespace {

__device__ __forceinline__ double shfl_t(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ zcomplex shfl_t(zcomplex v, int src)
{ return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)); }
__device__ __forceinline__ double shfl_xor_t(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ zcomplex shfl_xor_t(zcomplex v, int m)
{ return make_double2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m)); }
__device__ __forceinline__ double t_div(double a, double b) { return a / b; }
__device__ __forceinline__ zcomplex t_div(zcomplex a, zcomplex b) { return t_mul(a, t_recip(b)); }

// element (i, k) of op(A): A, A^T or A^H
template <typename T, bool TRANS, bool CONJ>
__device__ __forceinline__ T op_elem(const T *__restrict__ A, int64_t lda, int i, int k)
{
    T v = TRANS ? A[k + (int64_t)i * lda] : A[i + (int64_t)k * lda];
    return CONJ ? t_conj(v) : v;
}

// One CTA per right-hand side.  X[0:kb] <- tri(op(A))^-1 X[0:kb]; FORWARD: op(A) is lower triangular (L, or U^T / U^H),
// else upper (U, or L^T / L^H); UNIT: unit diagonal (the L factor).  32-row sub-blocks: a warp-shuffle substitution on the
// diagonal sub-block, then all threads update the rows still to be solved.
template <typename T, bool FORWARD, bool UNIT, bool TRANS, bool CONJ>
__global__ void __launch_bounds__(256)
trsv_block_kernel(int kb, const T *__restrict__ A, int64_t lda, T *__restrict__ X, int64_t ldx, int nrhs)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    T *xs = reinterpret_cast<T *>(smraw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nsb = (kb + 31) / 32;
    for (int rhs = blockIdx.x; rhs < nrhs; rhs += gridDim.x) {
    T *xcol = X + (int64_t)rhs * ldx;
    for (int i = tid; i < kb; i += blockDim.x) xs[i] = xcol[i];
    __syncthreads();
    for (int q = 0; q < nsb; ++q) {
        const int sb = FORWARD ? q * 32 : (nsb - 1 - q) * 32;
        const int bs = min(32, kb - sb);
        if (warp == 0) {
            const int i = sb + lane;
            const bool valid = lane < bs;
            T lrow[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) lrow[k] = (valid && k < bs) ? op_elem<T, TRANS, CONJ>(A, lda, i, sb + k) : t_zero(T());
            T xi = valid ? xs[i] : t_zero(T());
            if (FORWARD) {
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    if (!UNIT && valid && lane == k) xi = t_div(xi, lrow[k]);
                    T xk = shfl_t(xi, k);
                    if (valid && lane > k && k < bs) xi = t_fnma(lrow[k], xk, xi);
                }
            } else {
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) {
                    const int k = 31 - kk;
                    if (!UNIT && valid && lane == k) xi = t_div(xi, lrow[k]);
                    T xk = shfl_t(xi, k);
                    if (valid && lane < k && k < bs) xi = t_fnma(lrow[k], xk, xi);
                }
            }
            if (valid) xs[i] = xi;
        }
        __syncthreads();
        // rows outside the sub-block still to be solved
        const int r0 = FORWARD ? sb + bs : 0;
        const int r1 = FORWARD ? kb : sb;
        for (int i = r0 + tid; i < r1; i += blockDim.x) {
            T acc = xs[i];
#pragma unroll 8
            for (int k = 0; k < bs; ++k) acc = t_fnma(op_elem<T, TRANS, CONJ>(A, lda, i, sb + k), xs[sb + k], acc);
            xs[i] = acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < kb; i += blockDim.x) xcol[i] = xs[i];
    __syncthreads();
    }
}

// Y[c] -= sum_i op(A[i, c]) * X[i] for the kb rows of one block row and ncols columns (the transposed solves: the partial
// products are indexed by COLUMN).  One warp per column: the kb elements of a column are contiguous, so the loads coalesce.
template <typename T, bool CONJ>
__global__ void __launch_bounds__(256)
gemvt_minus_kernel(int kb, int64_t ncols, const T *__restrict__ A, int64_t lda, const T *__restrict__ X, int64_t ldx,
                   T *__restrict__ Y, int64_t ldy, int nrhs)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    T *xs = reinterpret_cast<T *>(smraw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int rhs = 0; rhs < nrhs; ++rhs) {
        for (int k = threadIdx.x; k < kb; k += blockDim.x) xs[k] = X[k + (int64_t)rhs * ldx];
        __syncthreads();
        for (int64_t c = (int64_t)blockIdx.x * wpb + warp; c < ncols; c += (int64_t)gridDim.x * wpb) {
            const T *ap = A + c * lda;
            T acc = t_zero(T());
            for (int i = lane; i < kb; i += 32) { T a = ap[i]; acc = t_fnma(CONJ ? t_conj(a) : a, xs[i], acc); }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc = t_add(acc, shfl_xor_t(acc, off));
            if (lane == 0) { T *yp = Y + c + (int64_t)rhs * ldy; *yp = t_add(*yp, acc); }
        }
        __syncthreads();
    }
}

// Y[rows] -= A[rows x kb] * X[kb], blockIdx.y = right-hand side.  Thread per row, K split over blockIdx.z
// slices accumulated with atomics when ksplit > 1.
template <typename T>
__global__ void __launch_bounds__(256)
gemv_minus_kernel(int64_t rows, int kb, const T *__restrict__ A, int64_t lda, const T *__restrict__ X, int64_t ldx,
                  T *__restrict__ Y, int64_t ldy, int nrhs)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    T *xs = reinterpret_cast<T *>(smraw);
    for (int rhs = blockIdx.y; rhs < nrhs; rhs += gridDim.y) {
    const T *xcol = X + (int64_t)rhs * ldx;
    __syncthreads();
    for (int k = threadIdx.x; k < kb; k += blockDim.x) xs[k] = xcol[k];
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) continue;
    const T *ap = A + i;
    T acc0 = t_zero(T()), acc1 = t_zero(T()), acc2 = t_zero(T()), acc3 = t_zero(T());
    int k = 0;
    for (; k + 4 <= kb; k += 4) {
        T a0 = ap[(int64_t)(k + 0) * lda], a1 = ap[(int64_t)(k + 1) * lda], a2 = ap[(int64_t)(k + 2) * lda], a3 = ap[(int64_t)(k + 3) * lda];
        acc0 = t_fnma(a0, xs[k + 0], acc0); acc1 = t_fnma(a1, xs[k + 1], acc1);
        acc2 = t_fnma(a2, xs[k + 2], acc2); acc3 = t_fnma(a3, xs[k + 3], acc3);
    }
    for (; k < kb; ++k) acc0 = t_fnma(ap[(int64_t)k * lda], xs[k], acc0);
    T *yp = Y + i + (int64_t)rhs * ldy;
    *yp = t_add(*yp, t_add(t_add(acc0, acc1), t_add(acc2, acc3)));
    }
}

// dst[i + c*ldd] = src[perm[i] + c*lds]   (row gather, all right-hand sides); SCATTER: dst[perm[i] + c*ldd] = src[i + c*lds]
template <typename T, bool SCATTER>
__global__ void __launch_bounds__(256)
gather_rows_kernel(int64_t n, const int *__restrict__ perm, const T *__restrict__ src, int64_t lds, T *__restrict__ dst,
                   int64_t ldd, int nrhs)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = perm[i];
    if (p < 0) return;
    for (int c = blockIdx.y; c < nrhs; c += gridDim.y) {
        if (SCATTER) dst[p + (int64_t)c * ldd] = src[i + (int64_t)c * lds];
        else dst[i + (int64_t)c * ldd] = src[p + (int64_t)c * lds];
    }
}

}  // namespace

static unsigned rhs_grid(int nrhs) { return (unsigned)(nrhs < 65535 ? nrhs : 65535); }

// mode: TRSV_UPPER = the stored triangle is U (non-unit), else L (unit); TRSV_TRANS / TRSV_CONJ: solve with A^T / A^H
template <typename T>
void launch_trsv_block(int kb, const T *Akk, int64_t lda, T *X, int64_t ldx, int nrhs, int mode, cudaStream_t s)
{
    if (kb <= 0 || nrhs <= 0) return;
    size_t sm = (size_t)kb * sizeof(T);
    if (sm > 48 * 1024) fatal("block size %d too large for the diagonal solve", kb);
    const unsigned g = rhs_grid(nrhs);
    const bool upper = mode & TRSV_UPPER, trans = mode & TRSV_TRANS, conj = (mode & TRSV_CONJ) && sizeof(T) == 16;
    if (!upper && (mode & TRSV_NONUNIT_L)) {                     // Cholesky factor: lower triangle with a general diagonal
        if (!trans) trsv_block_kernel<T, true, false, false, false><<<g, 256, sm, s>>>(kb, Akk, lda, X, ldx, nrhs);
        else if (!conj) trsv_block_kernel<T, false, false, true, false><<<g, 256, sm, s>>>(kb, Akk, lda, X, ldx, nrhs);
        else trsv_block_kernel<T, false, false, true, true><<<g, 256, sm, s>>>(kb, Akk, lda, X, ldx, nrhs);
    } else if (!trans) {
        if (upper) trsv_block_kernel<T, false, false, false, false><<<g, 256, sm, s>>>(kb, Akk, lda, X, ldx, nrhs);
        else trsv_block_kernel<T, true, true, false, false><<<g, 256, sm, s>>>(kb, Akk, lda, X, ldx, nrhs);
    } else if (!conj) {
        if (upper) trsv_block_kernel<T, true, false, true, false><<<g, 256, sm, s>>>(kb, Akk, lda, X, ldx, nrhs);
        else trsv_block_kernel<T, false, true, true, false><<<g, 256, sm, s>>>(kb, Akk, lda, X, ldx, nrhs);
    } else {
        if (upper) trsv_block_kernel<T, true, false, true, true><<<g, 256, sm, s>>>(kb, Akk, lda, X, ldx, nrhs);
        else trsv_block_kernel<T, false, true, true, true><<<g, 256, sm, s>>>(kb, Akk, lda, X, ldx, nrhs);
    }
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}
template <typename T>
void launch_gemv_minus(int64_t rows, int kb, const T *A, int64_t lda, const T *X, int64_t ldx, T *Y, int64_t ldy, int nrhs,
                       cudaStream_t s)
{
    if (rows <= 0 || kb <= 0 || nrhs <= 0) return;
    dim3 grid((unsigned)((rows + 255) / 256), rhs_grid(nrhs));
    gemv_minus_kernel<T><<<grid, 256, (size_t)kb * sizeof(T), s>>>(rows, kb, A, lda, X, ldx, Y, ldy, nrhs);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}
template <typename T>
void launch_gemvt_minus(int kb, int64_t ncols, const T *A, int64_t lda, const T *X, int64_t ldx, T *Y, int64_t ldy, int nrhs,
                        bool conj, cudaStream_t s)
{
    if (ncols <= 0 || kb <= 0 || nrhs <= 0) return;
    unsigned grid = (unsigned)std::min<int64_t>((ncols + 7) / 8, 148 * 8);
    if (conj && sizeof(T) == 16) gemvt_minus_kernel<T, true><<<grid, 256, (size_t)kb * sizeof(T), s>>>(kb, ncols, A, lda, X, ldx, Y, ldy, nrhs);
    else gemvt_minus_kernel<T, false><<<grid, 256, (size_t)kb * sizeof(T), s>>>(kb, ncols, A, lda, X, ldx, Y, ldy, nrhs);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}
template <typename T>
void launch_gather_rows(int64_t n, const int *perm, const T *src, int64_t lds, T *dst, int64_t ldd, int nrhs, cudaStream_t s, bool scatter)
{
    if (n <= 0 || nrhs <= 0) return;
    dim3 grid((unsigned)((n + 255) / 256), rhs_grid(nrhs));
    if (scatter) gather_rows_kernel<T, true><<<grid, 256, 0, s>>>(n, perm, src, lds, dst, ldd, nrhs);
    else gather_rows_kernel<T, false><<<grid, 256, 0, s>>>(n, perm, src, lds, dst, ldd, nrhs);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}

#define INST(T)                                                                                                  \
    template void launch_trsv_block<T>(int, const T *, int64_t, T *, int64_t, int, int, cudaStream_t);           \
    template void launch_gemv_minus<T>(int64_t, int, const T *, int64_t, const T *, int64_t, T *, int64_t, int, cudaStream_t); \
    template void launch_gemvt_minus<T>(int, int64_t, const T *, int64_t, const T *, int64_t, T *, int64_t, int, bool, cudaStream_t); \
    template void launch_gather_rows<T>(int64_t, const int *, const T *, int64_t, T *, int64_t, int, cudaStream_t, bool);
INST(double)
INST(zcomplex)
#undef INST

}  // namespace slb
