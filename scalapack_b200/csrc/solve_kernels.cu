// solve_kernels.cu -- kernels of the triangular solves of PDGETRS (SRC/pdgetrs.f:255-266; the reference
// runs PDLAPIV + two PDTRSMs, PBLAS/SRC/PTOOLS/PB_CptrsmB.c:441-590 for few right-hand sides).
// HBM-bound: L and U are each streamed once (8 N^2 bytes in total for one right-hand side).
#include "kernels.cuh"
#include "devmath.cuh"
#include "common.h"

namespace slb {

namespace {

__device__ __forceinline__ double shfl_t(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ zcomplex shfl_t(zcomplex v, int src)
{ return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)); }
__device__ __forceinline__ double t_div(double a, double b) { return a / b; }
__device__ __forceinline__ zcomplex t_div(zcomplex a, zcomplex b) { return t_mul(a, t_recip(b)); }

// One CTA per right-hand side.  X[0:kb] <- tri(A)^-1 X[0:kb]; 32-row sub-blocks: a warp-shuffle substitution
// on the diagonal sub-block, then all threads update the remaining rows.
template <typename T, bool UPPER>
__global__ void __launch_bounds__(256)
trsv_block_kernel(int kb, const T *__restrict__ A, int64_t lda, T *__restrict__ X, int64_t ldx)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    T *xs = reinterpret_cast<T *>(smraw);
    T *xcol = X + (int64_t)blockIdx.x * ldx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kb; i += blockDim.x) xs[i] = xcol[i];
    __syncthreads();
    const int nsb = (kb + 31) / 32;
    for (int q = 0; q < nsb; ++q) {
        const int sb = UPPER ? (nsb - 1 - q) * 32 : q * 32;
        const int bs = min(32, kb - sb);
        if (warp == 0) {
            const int i = sb + lane;
            const bool valid = lane < bs;
            T lrow[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) lrow[k] = (valid && k < bs) ? A[i + (int64_t)(sb + k) * lda] : t_zero(T());
            T xi = valid ? xs[i] : t_zero(T());
            if (!UPPER) {
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    T xk = shfl_t(xi, k);
                    if (valid && lane > k && k < bs) xi = t_fnma(lrow[k], xk, xi);
                }
            } else {
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) {
                    const int k = 31 - kk;
                    if (valid && lane == k) xi = t_div(xi, lrow[k]);
                    T xk = shfl_t(xi, k);
                    if (valid && lane < k && k < bs) xi = t_fnma(lrow[k], xk, xi);
                }
            }
            if (valid) xs[i] = xi;
        }
        __syncthreads();
        // rows outside the sub-block still to be solved
        const int r0 = UPPER ? 0 : sb + bs;
        const int r1 = UPPER ? sb : kb;
        for (int i = r0 + tid; i < r1; i += blockDim.x) {
            T acc = xs[i];
#pragma unroll 8
            for (int k = 0; k < bs; ++k) acc = t_fnma(A[i + (int64_t)(sb + k) * lda], xs[sb + k], acc);
            xs[i] = acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < kb; i += blockDim.x) xcol[i] = xs[i];
}

// Y[rows] -= A[rows x kb] * X[kb], blockIdx.y = right-hand side.  Thread per row, K split over blockIdx.z
// slices accumulated with atomics when ksplit > 1.
template <typename T>
__global__ void __launch_bounds__(256)
gemv_minus_kernel(int64_t rows, int kb, const T *__restrict__ A, int64_t lda, const T *__restrict__ X, int64_t ldx,
                  T *__restrict__ Y, int64_t ldy)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    T *xs = reinterpret_cast<T *>(smraw);
    const T *xcol = X + (int64_t)blockIdx.y * ldx;
    for (int k = threadIdx.x; k < kb; k += blockDim.x) xs[k] = xcol[k];
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const T *ap = A + i;
    T acc0 = t_zero(T()), acc1 = t_zero(T()), acc2 = t_zero(T()), acc3 = t_zero(T());
    int k = 0;
    for (; k + 4 <= kb; k += 4) {
        T a0 = ap[(int64_t)(k + 0) * lda], a1 = ap[(int64_t)(k + 1) * lda], a2 = ap[(int64_t)(k + 2) * lda], a3 = ap[(int64_t)(k + 3) * lda];
        acc0 = t_fnma(a0, xs[k + 0], acc0); acc1 = t_fnma(a1, xs[k + 1], acc1);
        acc2 = t_fnma(a2, xs[k + 2], acc2); acc3 = t_fnma(a3, xs[k + 3], acc3);
    }
    for (; k < kb; ++k) acc0 = t_fnma(ap[(int64_t)k * lda], xs[k], acc0);
    T *yp = Y + i + (int64_t)blockIdx.y * ldy;
    *yp = t_add(*yp, t_add(t_add(acc0, acc1), t_add(acc2, acc3)));
}

// dst[i + c*ldd] = src[perm[i] + c*lds]   (row gather, all right-hand sides)
template <typename T>
__global__ void __launch_bounds__(256)
gather_rows_kernel(int64_t n, const int *__restrict__ perm, const T *__restrict__ src, int64_t lds, T *__restrict__ dst,
                   int64_t ldd)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = perm[i];
    if (p >= 0) dst[i + (int64_t)blockIdx.y * ldd] = src[p + (int64_t)blockIdx.y * lds];
}

}  // namespace

template <typename T>
void launch_trsv_block(int kb, const T *Akk, int64_t lda, T *X, int64_t ldx, int nrhs, int upper, cudaStream_t s)
{
    if (kb <= 0 || nrhs <= 0) return;
    size_t sm = (size_t)kb * sizeof(T);
    if (upper) trsv_block_kernel<T, true><<<nrhs, 256, sm, s>>>(kb, Akk, lda, X, ldx);
    else trsv_block_kernel<T, false><<<nrhs, 256, sm, s>>>(kb, Akk, lda, X, ldx);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}
template <typename T>
void launch_gemv_minus(int64_t rows, int kb, const T *A, int64_t lda, const T *X, int64_t ldx, T *Y, int64_t ldy, int nrhs,
                       cudaStream_t s)
{
    if (rows <= 0 || kb <= 0 || nrhs <= 0) return;
    dim3 grid((unsigned)((rows + 255) / 256), (unsigned)nrhs);
    gemv_minus_kernel<T><<<grid, 256, (size_t)kb * sizeof(T), s>>>(rows, kb, A, lda, X, ldx, Y, ldy);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}
template <typename T>
void launch_gather_rows(int64_t n, const int *perm, const T *src, int64_t lds, T *dst, int64_t ldd, int nrhs, cudaStream_t s)
{
    if (n <= 0 || nrhs <= 0) return;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)nrhs);
    gather_rows_kernel<T><<<grid, 256, 0, s>>>(n, perm, src, lds, dst, ldd);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}

#define INST(T)                                                                                                  \
    template void launch_trsv_block<T>(int, const T *, int64_t, T *, int64_t, int, int, cudaStream_t);           \
    template void launch_gemv_minus<T>(int64_t, int, const T *, int64_t, const T *, int64_t, T *, int64_t, int, cudaStream_t); \
    template void launch_gather_rows<T>(int64_t, const int *, const T *, int64_t, T *, int64_t, int, cudaStream_t);
INST(double)
INST(zcomplex)
#undef INST

}  // namespace slb
