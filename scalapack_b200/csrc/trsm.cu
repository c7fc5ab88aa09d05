// trsm.cu -- U12 <- unit_lower(L11)^-1 * U12   (PDTRSM 'L','L','N','U' of SRC/pdgetrf.f:280).
//
// The reference strip-mines this by 32 rows (PBLAS/SRC/PTOOLS/PB_CptrsmAB.c:359-416: a dtrsm_ on the
// strip + a dgemm_ on the rows below).  Same structure here with DB-row strips: a substitution kernel on
// the DB x n strip (one thread per column of U12, the strip held in registers, the DB x DB diagonal block
// of L11 broadcast from shared memory) followed by the DMMA update kernel of gemm.cu on the rows below.
#include "kernels.cuh"
#include "devmath.cuh"
#include "common.h"

namespace slb {

namespace {

template <typename T, int DB>
__global__ void __launch_bounds__(128)
trsm_diag_kernel(int kb, int64_t n, const T *__restrict__ L, int64_t ldl, T *__restrict__ B, int64_t ldb)
{
    // Ls[k][i] = L[i][k] for i > k (strictly lower part), column k contiguous
    __shared__ T Ls[DB * DB];
    for (int e = threadIdx.x; e < DB * DB; e += blockDim.x) {
        int i = e % DB, k = e / DB;
        Ls[e] = (i < kb && k < kb && i > k) ? L[i + (int64_t)k * ldl] : t_zero(T());
    }
    __syncthreads();
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    T *col = B + c * ldb;
    T x[DB];
#pragma unroll
    for (int i = 0; i < DB; ++i) x[i] = (i < kb) ? col[i] : t_zero(T());
#pragma unroll
    for (int k = 0; k < DB - 1; ++k) {
        T xk = x[k];
#pragma unroll
        for (int i = k + 1; i < DB; ++i) x[i] = t_fnma(Ls[k * DB + i], xk, x[i]);
    }
#pragma unroll
    for (int i = 0; i < DB; ++i)
        if (i < kb) col[i] = x[i];
}

template <typename T> struct TrsmCfg;
template <> struct TrsmCfg<double> { static constexpr int DB = 64; };
template <> struct TrsmCfg<zcomplex> { static constexpr int DB = 32; };

inline void gemm_minus(int64_t M, int64_t N, int K, const double *A, int64_t lda, const double *B, int64_t ldb, double *C,
                       int64_t ldc, cudaStream_t s) { launch_dgemm_minus(M, N, K, A, lda, B, ldb, C, ldc, s); }
inline void gemm_minus(int64_t M, int64_t N, int K, const zcomplex *A, int64_t lda, const zcomplex *B, int64_t ldb,
                       zcomplex *C, int64_t ldc, cudaStream_t s) { launch_zgemm_minus(M, N, K, A, lda, B, ldb, C, ldc, s); }

template <typename T>
void trsm_llnu(int jb, int64_t n, const T *L, int64_t ldl, T *B, int64_t ldb, cudaStream_t s)
{
    constexpr int DB = TrsmCfg<T>::DB;
    if (jb <= 0 || n <= 0) return;
    for (int k0 = 0; k0 < jb; k0 += DB) {
        int kb = jb - k0 < DB ? jb - k0 : DB;
        unsigned grid = (unsigned)((n + 127) / 128);
        trsm_diag_kernel<T, DB><<<grid, 128, 0, s>>>(kb, n, L + k0 + (int64_t)k0 * ldl, ldl, B + k0, ldb);
        SLB_CUDA(cudaGetLastError());
        counter_add("kernel_launches", 1);
        int rem = jb - k0 - kb;
        if (rem > 0)   // rows below the strip: B[k0+kb:, :] -= L[k0+kb:, k0:k0+kb] * B[k0:k0+kb, :]
            gemm_minus(rem, n, kb, L + (k0 + kb) + (int64_t)k0 * ldl, ldl, B + k0, ldb, B + k0 + kb, ldb, s);
    }
}

}  // namespace

void launch_dtrsm_llnu(int jb, int64_t n, const double *L, int64_t ldl, double *B, int64_t ldb, cudaStream_t s)
{ trsm_llnu<double>(jb, n, L, ldl, B, ldb, s); }
void launch_ztrsm_llnu(int jb, int64_t n, const zcomplex *L, int64_t ldl, zcomplex *B, int64_t ldb, cudaStream_t s)
{ trsm_llnu<zcomplex>(jb, n, L, ldl, B, ldb, s); }

}  // namespace slb
