// solve_fast.cu -- PDGETRS 'N' on a 1 x 1 grid, few right-hand sides: the two triangular sweeps of SRC/pdgetrs.f:255-266 at HBM speed.
//
// The sweeps are HBM-bound (L and U are each read once: 8 N^2 bytes) but strictly sequential block to block: x_k is needed
// before the column block k can be folded into the rows below.  solve.cu walks the blocks with one stream and a single-CTA
// diagonal solve (~40 us per block: latency-bound); here the critical chain of a block step is cut to what really is serial
// and everything else runs beside it on a second stream:
//
//   critical stream:  diag(k)  ->  top(k)  ->  diag(k+1)  ->  top(k+1)  -> ...
//   bulk stream:                  bulk(k) ......  bulk(k+1) ......           (bulk(k) starts after diag(k); diag(k+2) waits for it)
//
//   diag(k)  x_k = tri(A_kk)^-1 (b_k + acc_k + sum_s P_s): one CTA per right-hand side, one WARP per 32-row sub-block.  Warp q
//            folds the already solved sub-blocks into its rows as their x appears in shared memory (flags, no block barrier), then
//            solves its 32 x 32 triangle by warp-shuffle substitution: 16 overlapped sub-steps instead of 16 serial ones.
//   top(k)   the contribution of x_k to the rows of the NEXT block only (nb x nb, split over K into partial sums P_s that
//            diag(k+1) adds in a fixed order: deterministic), a few CTAs, ~3 us.
//   bulk(k)  acc -= A[rows beyond the next block, block k] x_k: the HBM stream (16-byte loads, 8 columns in flight per thread).
// Only bulk kernels write `acc`, only top(k) writes P: no atomics, results are reproducible bit for bit.
#include "common.h"
#include "kernels.cuh"
#include "lu.h"

#include <algorithm>

namespace slb {

namespace {

constexpr int SUB = 32;            // rows per sub-block of the diagonal solve
constexpr int MAXSUB = 16;         // nb <= 512
constexpr int KSPLIT = 32;         // K slices of top()
constexpr int MAXRHS = 8;
constexpr int BULK_MAXS = 8;         // K slices of bulk()

// Loads and polls as volatile asm: the compiler keeps volatile asm statements in program order, so a load written BEFORE a poll
// is issued before the poll (left to itself it sinks the loads to their first use, i.e. behind the wait, and every fold then
// pays a full L2 round trip on the critical path -- measured: 0.66 us per fold instead of 0.1 us).
__device__ __forceinline__ double ldg_early(const double *p)
{ double v; asm volatile("ld.global.nc.f64 %0, [%1];\n" : "=d"(v) : "l"(p)); return v; }
__device__ __forceinline__ int ld_flag(const int *p)
{ int v; asm volatile("ld.volatile.shared.s32 %0, [%1];\n" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory"); return v; }

// Inverses of the 32 x 32 diagonal triangles of every nb x nb diagonal block (one warp each; lane = row, the 32 columns of the
// inverse in registers, row k broadcast by shuffles): Dinv[((which*nblk + k)*MAXSUB + s)*1024 + c*32 + i] = inv(T)(i, c), which = 0:
// unit lower triangle of L, 1: upper triangle of U.  ~20 us for N = 65536, re-done per solve (the factors are the caller's).
__global__ void __launch_bounds__(128)
inv32_kernel(int N, int nb, int nblk, const double *__restrict__ A, int64_t lda, double *__restrict__ Dinv)
{
    const int lane = threadIdx.x & 31;
    const int task = blockIdx.x * 4 + (threadIdx.x >> 5);                // (which, k, s)
    if (task >= 2 * nblk * MAXSUB) return;
    const int which = task / (nblk * MAXSUB), k = (task / MAXSUB) % nblk, s = task % MAXSUB;
    const int kb = min(nb, N - k * nb), sb = s * SUB;
    if (sb >= kb) return;
    const int bs = min(SUB, kb - sb);
    const double *T = A + ((int64_t)k * nb + sb) * (lda + 1);
    double t[SUB], inv[SUB];
#pragma unroll
    for (int c = 0; c < SUB; ++c) {
        t[c] = (lane < bs && c < bs) ? T[lane + (int64_t)c * lda] : (lane == c ? 1.0 : 0.0);
        inv[c] = lane == c ? 1.0 : 0.0;
    }
    if (which == 0) {                                                    // unit lower: forward
#pragma unroll
        for (int kk = 0; kk < SUB; ++kk) {
#pragma unroll
            for (int c = 0; c < SUB; ++c) {
                const double r = __shfl_sync(0xffffffffu, inv[c], kk);
                if (lane > kk) inv[c] = fma(-t[kk], r, inv[c]);
            }
        }
    } else {                                                             // upper, non-unit: backward
#pragma unroll
        for (int q = 0; q < SUB; ++q) {
            const int kk = SUB - 1 - q;
            const double d = __shfl_sync(0xffffffffu, t[kk], kk);
#pragma unroll
            for (int c = 0; c < SUB; ++c) {
                if (lane == kk) inv[c] = inv[c] / d;
                const double r = __shfl_sync(0xffffffffu, inv[c], kk);
                if (lane < kk) inv[c] = fma(-t[kk], r, inv[c]);
            }
        }
    }
    double *out = Dinv + (size_t)task * (SUB * SUB);
#pragma unroll
    for (int c = 0; c < SUB; ++c) out[c * SUB + lane] = inv[c];
}

// x[0:kb] = tri(A)^-1 (b + acc + sum_s P[s]);  FORWARD: unit lower triangle (L), else non-unit upper triangle (U).
// A: kb x kb diagonal block (ld = lda).  Xk, Acc: this block's rows of X and acc (ld = ldx).  P: [nparts][nrhs][kb].
// Dk: the inverted 32 x 32 diagonal triangles of this block.  NW warps; warp w owns the sub-blocks w, w + NW, ... of the sweep
// order (its second sub-block comes up long after its first).  Per sub-block the serial chain is: flag -> 32 FMAs in four
// independent chains (fold of the sub-block just solved) -> 32 shuffles + FMAs with the inverted triangle -> flag.
constexpr int NW = 8;
constexpr size_t DIAG_SMEM = (size_t)NW * 3 * SUB * SUB * sizeof(double);      // per warp: two staged 32 x 32 blocks + the inverted triangle

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc, bool pred)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = pred ? 8 : 0;                                     // 0: the 8 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// The 32 x 32 blocks a warp works on (the off-diagonal blocks it folds, one ahead, and its inverted triangle) are STAGED IN
// SHARED MEMORY by cp.async: held in registers they take > 190 registers per thread, and the compiler then funnels every
// shared-memory load of x through the same four registers -- one load latency per FMA (measured: 0.66 us per fold).
template <bool FORWARD>
__global__ void __launch_bounds__(SUB * NW, 1)
diag_solve_kernel(int kb, const double *__restrict__ A, int64_t lda, double *__restrict__ Xk, const double *__restrict__ Acc,
                  int64_t ldx, const double *__restrict__ P, int nparts, int nrhs, const double *__restrict__ Dk,
                  const double *__restrict__ pf, int pf_cols, unsigned sleep_ns, long long *__restrict__ dbg = nullptr)
{
    extern __shared__ __align__(16) double stage[];                     // [NW][3][SUB (k)][SUB (lane)]
    __shared__ double xs[SUB * MAXSUB];
    __shared__ int flag[MAXSUB];
    const long long t_begin = dbg ? clock64() : 0;
    auto stamp = [&](int q, int k) { if (dbg && (threadIdx.x & 31) == 0) dbg[q * 8 + k] = clock64() - t_begin; };
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nsub = (kb + SUB - 1) / SUB;
    const int rhs = blockIdx.x;
    double *mybuf = stage + (size_t)w * 3 * SUB * SUB;
    if (tid < MAXSUB) flag[tid] = 0;
    __syncthreads();
    for (int q = w; q < nsub; q += NW) {
        const int sq = FORWARD ? q : nsub - 1 - q;                       // sub-block q of the sweep
        const int sb = sq * SUB;
        const int i = sb + lane;
        const bool valid = i < kb;
        stamp(q, 0);
        // block p of my rows -> staging buffer p & 1 (a rolled loop with a running pointer: 32 unrolled 64-bit addresses
        // would cost the registers the FMA pipelines below need)
        auto load_blk = [&](int p) {
            const int pb = (FORWARD ? p : nsub - 1 - p) * SUB;
            const int pbs = min(SUB, kb - pb);
            double *dst = mybuf + (size_t)(p & 1) * SUB * SUB + lane;
            const double *src = A + (valid ? i : 0) + (int64_t)pb * lda;
#pragma unroll 2
            for (int k = 0; k < SUB; ++k) { cp_async8(dst, src, valid && k < pbs); dst += SUB; if (k + 1 < pbs) src += lda; }
            cp_async_commit();
        };
        {   // my row of the inverted diagonal triangle -> third staging buffer (needed last)
            double *dst = mybuf + (size_t)2 * SUB * SUB + lane;
            const double *src = Dk + (size_t)sq * (SUB * SUB) + lane;
#pragma unroll 2
            for (int c = 0; c < SUB; ++c) { cp_async8(dst, src, true); dst += SUB; src += SUB; }
            cp_async_commit();
        }
        if (q > 0) load_blk(0);
        // right-hand side of my rows: b + acc + the K slices of top() in a FIXED order; the loads are issued 16 at a time
        // (a loop of dependent load + add pairs would cost one L2 round trip per slice)
        double v = 0.0;
        if (valid) {
            v = Xk[i + (int64_t)rhs * ldx] + Acc[i + (int64_t)rhs * ldx];
#pragma unroll
            for (int s0 = 0; s0 < KSPLIT; s0 += 16) {
                double pv[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) pv[u] = (s0 + u < nparts) ? ldg_early(P + ((int64_t)(s0 + u) * nrhs + rhs) * kb + i) : 0.0;
#pragma unroll
                for (int u = 0; u < 16; ++u) v += pv[u];
            }
        }
        stamp(q, 1);
        // fold the sub-blocks solved before mine, in sweep order, as their x appears
        for (int p = 0; p < q; ++p) {
            if (p + 1 < q) { load_blk(p + 1); cp_async_wait<1>(); } else cp_async_wait<0>();
            if (lane == 0) while (ld_flag(&flag[p]) == 0) { if (sleep_ns > 0) __nanosleep(sleep_ns); }
            if (p == q - 1) stamp(q, 4);
            __syncwarp();
            if (p == q - 1) stamp(q, 5);
            const int pb = (FORWARD ? p : nsub - 1 - p) * SUB;
            const double *a = mybuf + (size_t)(p & 1) * SUB * SUB + lane;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int k0 = 0; k0 < SUB; k0 += 8) {                        // 16 shared-memory loads in flight, then 8 FMAs
                double av[8], xv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { av[u] = a[(k0 + u) * SUB]; xv[u] = xs[pb + k0 + u]; }
                s0 = fma(av[0], xv[0], s0); s1 = fma(av[1], xv[1], s1); s2 = fma(av[2], xv[2], s2); s3 = fma(av[3], xv[3], s3);
                s0 = fma(av[4], xv[4], s0); s1 = fma(av[5], xv[5], s1); s2 = fma(av[6], xv[6], s2); s3 = fma(av[7], xv[7], s3);
            }
            v -= (s0 + s1) + (s2 + s3);
            __syncwarp();                                                // buffer p & 1 is free for block p + 2
        }
        if (q == 0) cp_async_wait<0>();
        stamp(q, 2);
        // my triangle: x = inv(T) v, the 32 values of v broadcast by shuffles, four independent chains
        const double *dinv = mybuf + (size_t)2 * SUB * SUB + lane;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int c0 = 0; c0 < SUB; c0 += 8) {
            double dv[8], vv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { dv[u] = dinv[(c0 + u) * SUB]; vv[u] = __shfl_sync(0xffffffffu, v, c0 + u); }
            s0 = fma(dv[0], vv[0], s0); s1 = fma(dv[1], vv[1], s1); s2 = fma(dv[2], vv[2], s2); s3 = fma(dv[3], vv[3], s3);
            s0 = fma(dv[4], vv[4], s0); s1 = fma(dv[5], vv[5], s1); s2 = fma(dv[6], vv[6], s2); s3 = fma(dv[7], vv[7], s3);
        }
        v = (s0 + s1) + (s2 + s3);
        stamp(q, 6);
        // publish: shared memory first (the fence then only has shared-memory stores to order), the global copy afterwards
        if (valid) xs[i] = v;
        __syncwarp();
        stamp(q, 7);
        if (lane == 0) {
            __threadfence_block();
            asm volatile("st.volatile.shared.s32 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(&flag[q])), "r"(1) : "memory");
        }
        stamp(q, 3);
        if (valid) Xk[i + (int64_t)rhs * ldx] = v;
    }
    // the next diagonal block -> L2 for the next launch of this kernel (a few us away): one 128-byte line per thread and
    // column, by the warps as they run out of work
    if (pf != nullptr && blockIdx.x == 0) {
        const int lines = (pf_cols * 8 + 127) >> 7;
        for (int c = w; c < pf_cols; c += NW)
            for (int l = lane; l < lines; l += 32)
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"((const char *)(pf + (int64_t)c * lda) + l * 128));
    }
}

// rows [0, nr) of Y (-)= A[nr x kb] X[kb] for nrhs right-hand sides.  blockIdx.x: 256-row chunk, blockIdx.y: K slice.
// PARTIAL: the slice's product (negated) is WRITTEN to P[slice][rhs][row] (top); else Y[row + rhs*ldy] -= product (bulk, one slice).
template <bool VEC, bool PARTIAL, int NR, int UNR>
__global__ void __launch_bounds__(128)
gemv_rows_kernel(int64_t nr, int kb, const double *__restrict__ A, int64_t lda, const double *__restrict__ X, int64_t ldx,
                 double *__restrict__ Y, int64_t ldy, int nrhs, int kslice)
{
    __shared__ double xs[NR * SUB * MAXSUB];
    const int k0 = PARTIAL ? blockIdx.y * kslice : 0;
    const int k1 = PARTIAL ? min(kb, k0 + kslice) : kb;
    for (int e = threadIdx.x; e < NR * (k1 - k0); e += blockDim.x) {
        const int r = e / (k1 - k0), k = e % (k1 - k0);
        xs[r * (SUB * MAXSUB) + k] = r < nrhs ? X[k0 + k + (int64_t)r * ldx] : 0.0;
    }
    __syncthreads();
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (row >= nr) return;
    const bool two = row + 1 < nr;
    double acc0[NR], acc1[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) { acc0[r] = 0.0; acc1[r] = 0.0; }
    const double *ap = A + row + (int64_t)k0 * lda;
    const int kn = k1 - k0;
    int k = 0;
    for (; k + UNR <= kn; k += UNR) {
        double2 a[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const double *p = ap + (int64_t)(k + u) * lda;
            if (VEC) a[u] = __ldcs(reinterpret_cast<const double2 *>(p));
            else { a[u].x = __ldcs(p); a[u].y = two ? __ldcs(p + 1) : 0.0; }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int r = 0; r < NR; ++r) { const double x = xs[r * (SUB * MAXSUB) + k + u]; acc0[r] = fma(a[u].x, x, acc0[r]); acc1[r] = fma(a[u].y, x, acc1[r]); }
    }
    for (; k < kn; ++k) {
        const double *p = ap + (int64_t)k * lda;
        const double ax = p[0], ay = two ? p[1] : 0.0;
#pragma unroll
        for (int r = 0; r < NR; ++r) { const double x = xs[r * (SUB * MAXSUB) + k]; acc0[r] = fma(ax, x, acc0[r]); acc1[r] = fma(ay, x, acc1[r]); }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        if (r >= nrhs) continue;
        if (PARTIAL) {
            double *pp = Y + ((int64_t)blockIdx.y * nrhs + r) * nr + row;
            pp[0] = -acc0[r]; if (two) pp[1] = -acc1[r];
        } else {
            double *yp = Y + row + (int64_t)r * ldy;
            yp[0] -= acc0[r]; if (two) yp[1] -= acc1[r];
        }
    }
}

// bulk(): Y[row] -= A[nr x kb] X[kb] with the K range cut into S slices (blockIdx.y) so that few remaining rows still put
// enough loads in flight to stream at HBM speed.  Slices write their partial products to `scratch`; the LAST slice CTA of a
// row chunk to arrive (counter) adds them up in slice order 0 .. S-1 and updates Y: the result does not depend on arrival order.
template <bool VEC, int NR, int UNR>
__global__ void __launch_bounds__(128)
gemv_bulk_kernel(int64_t nr, int kb, const double *__restrict__ A, int64_t lda, const double *__restrict__ X, int64_t ldx,
                 double *__restrict__ Y, int64_t ldy, int nrhs, int kslice, int S, double *__restrict__ scratch,
                 unsigned *__restrict__ counters)
{
    __shared__ double xs[NR * SUB * MAXSUB];
    __shared__ int is_last;
    const int k0 = blockIdx.y * kslice, k1 = min(kb, k0 + kslice), kn = k1 - k0;
    for (int e = threadIdx.x; e < NR * kn; e += blockDim.x) {
        const int r = e / kn, k = e % kn;
        xs[r * (SUB * MAXSUB) + k] = r < nrhs ? X[k0 + k + (int64_t)r * ldx] : 0.0;
    }
    __syncthreads();
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const bool active = row < nr, two = row + 1 < nr;
    double acc0[NR], acc1[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) { acc0[r] = 0.0; acc1[r] = 0.0; }
    if (active) {
        const double *ap = A + row + (int64_t)k0 * lda;
        int k = 0;
        for (; k + UNR <= kn; k += UNR) {
            double2 a[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const double *p = ap + (int64_t)(k + u) * lda;
                if (VEC) a[u] = __ldcs(reinterpret_cast<const double2 *>(p));
                else { a[u].x = __ldcs(p); a[u].y = two ? __ldcs(p + 1) : 0.0; }
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u)
#pragma unroll
                for (int r = 0; r < NR; ++r) { const double x = xs[r * (SUB * MAXSUB) + k + u]; acc0[r] = fma(a[u].x, x, acc0[r]); acc1[r] = fma(a[u].y, x, acc1[r]); }
        }
        for (; k < kn; ++k) {
            const double *p = ap + (int64_t)k * lda;
            const double ax = p[0], ay = two ? p[1] : 0.0;
#pragma unroll
            for (int r = 0; r < NR; ++r) { const double x = xs[r * (SUB * MAXSUB) + k]; acc0[r] = fma(ax, x, acc0[r]); acc1[r] = fma(ay, x, acc1[r]); }
        }
    }
    if (S == 1) {
        if (active) {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                if (r >= nrhs) continue;
                double *yp = Y + row + (int64_t)r * ldy;
                yp[0] -= acc0[r]; if (two) yp[1] -= acc1[r];
            }
        }
        return;
    }
    const int64_t nrp = (nr + 1) & ~(int64_t)1;
    if (active) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double *sp = scratch + ((int64_t)blockIdx.y * NR + r) * nrp + row;
            __stcg(sp, acc0[r]); __stcg(sp + 1, acc1[r]);
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(&counters[blockIdx.x], 1u) == (unsigned)(S - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (active) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            if (r >= nrhs) continue;
            double s0 = 0.0, s1 = 0.0;
            for (int sl = 0; sl < S; ++sl) {
                const double *sp = scratch + ((int64_t)sl * NR + r) * nrp + row;
                s0 += __ldcg(sp); s1 += __ldcg(sp + 1);
            }
            double *yp = Y + row + (int64_t)r * ldy;
            yp[0] -= s0; if (two) yp[1] -= s1;
        }
    }
    if (threadIdx.x == 0) counters[blockIdx.x] = 0;
}

template <int NR>
void launch_gemv_rows_nr(int64_t nr, int kb, const double *A, int64_t lda, const double *X, int64_t ldx, double *Y, int64_t ldy, int nrhs,
                         bool partial, cudaStream_t s, double *scratch, unsigned *counters)
{
    constexpr int UNR = NR <= 2 ? 16 : 8;                              // 16-byte loads in flight per thread
    const bool vec = (((uintptr_t)A) & 15) == 0 && lda % 2 == 0 && nr % 2 == 0;
    // enough CTAs to fill the GPU also when few rows are left: 256, 128 or 64 rows per CTA
    const int sms = rt().sm_count;
    const int threads = 128;
    const unsigned gx = (unsigned)((nr + 2 * threads - 1) / (2 * threads));
    if (partial) {
        const int kslice = (kb + KSPLIT - 1) / KSPLIT;
        dim3 grid(gx, (unsigned)((kb + kslice - 1) / kslice));
        if (vec) gemv_rows_kernel<true, true, NR, UNR><<<grid, threads, 0, s>>>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, kslice);
        else gemv_rows_kernel<false, true, NR, UNR><<<grid, threads, 0, s>>>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, kslice);
    } else {
        // K slices: enough CTAs (4 per SM) also when few rows are left; slices are multiples of the load batch
        const unsigned chunks = (unsigned)((nr + 255) / 256);
        int S = 1;
        while (S < BULK_MAXS && (int64_t)chunks * S < (int64_t)sms * 4 && kb / (2 * S) >= UNR) S *= 2;
        const int kslice = ((kb + S - 1) / S + UNR - 1) / UNR * UNR;
        S = (kb + kslice - 1) / kslice;
        dim3 grid(chunks, (unsigned)S);
        if (vec) gemv_bulk_kernel<true, NR, UNR><<<grid, 128, 0, s>>>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, kslice, S, scratch, counters);
        else gemv_bulk_kernel<false, NR, UNR><<<grid, 128, 0, s>>>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, kslice, S, scratch, counters);
    }
}
void launch_gemv_rows(int64_t nr, int kb, const double *A, int64_t lda, const double *X, int64_t ldx, double *Y, int64_t ldy, int nrhs,
                      bool partial, cudaStream_t s, double *scratch = nullptr, unsigned *counters = nullptr)
{
    if (nr <= 0 || kb <= 0) return;
    if (nrhs == 1) launch_gemv_rows_nr<1>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, partial, s, scratch, counters);
    else if (nrhs == 2) launch_gemv_rows_nr<2>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, partial, s, scratch, counters);
    else if (nrhs <= 4) launch_gemv_rows_nr<4>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, partial, s, scratch, counters);
    else launch_gemv_rows_nr<8>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, partial, s, scratch, counters);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}

}  // namespace

bool getrs_fast_applies(int P, int Q, char trans, int nb, int nrhs)
{
    return P * Q == 1 && trans == 'N' && nb <= SUB * MAXSUB && nrhs >= 1 && nrhs <= MAXRHS && opt("solve_fast", 1) != 0;
}

// The two sweeps as a fork/join of the critical stream sa and the bulk stream sb (sb is the origin: on return it has waited for sa).
static void enqueue_sweeps(int N, int nrhs, const double *A, int64_t lld, int nb, double *Xg, double *acc, double *part, double *Dinv,
                           double *scratch, unsigned *counters, cudaStream_t sa, cudaStream_t sb, std::vector<cudaEvent_t> &evd, std::vector<cudaEvent_t> &evb,
                           cudaEvent_t fork, cudaEvent_t join)
{
    const int nblk = (N + nb - 1) / nb;
    const unsigned sleep_ns = (unsigned)opt("solve_poll_ns", 40);
    SLB_CUDA(cudaEventRecord(fork, sb));
    SLB_CUDA(cudaStreamWaitEvent(sa, fork, 0));
    inv32_kernel<<<(unsigned)((2 * nblk * MAXSUB + 3) / 4), 128, 0, sa>>>(N, nb, nblk, A, lld, Dinv);
    SLB_CUDA(cudaGetLastError()); counter_add("kernel_launches", 1);
    for (int pass = 0; pass < 2; ++pass) {
        const bool fwd = pass == 0;
        SLB_CUDA(cudaMemsetAsync(acc, 0, (size_t)N * nrhs * sizeof(double), sa));
        SLB_CUDA(cudaEventRecord(fork, sa));
        SLB_CUDA(cudaStreamWaitEvent(sb, fork, 0));
        int nparts_prev = 0;                                                  // K slices top(q-1) wrote
        for (int q = 0; q < nblk; ++q) {
            const int k = fwd ? q : nblk - 1 - q;
            const int64_t j0 = (int64_t)k * nb; const int kb = (int)std::min<int64_t>(nb, N - j0);
            const int kn = fwd ? k + 1 : k - 1;                                   // the block solved next
            const bool have_next = q + 1 < nblk;
            const int64_t jn = (int64_t)kn * nb; const int kbn = have_next ? (int)std::min<int64_t>(nb, N - jn) : 0;
            // ---- diag(k) ----
            if (q >= 2) SLB_CUDA(cudaStreamWaitEvent(sa, evb[q - 2], 0));         // bulk(q-2) has folded x into my rows
            const double *Pk = part + (size_t)(q & 1) * KSPLIT * MAXRHS * nb;     // written by top(q-1)
            const double *pf = have_next ? A + jn + jn * lld : nullptr;
            const double *Dk = Dinv + ((size_t)(fwd ? 0 : 1) * nblk + k) * MAXSUB * SUB * SUB;
            if (fwd) diag_solve_kernel<true><<<nrhs, SUB * NW, DIAG_SMEM, sa>>>(kb, A + j0 + j0 * lld, lld, Xg + j0, acc + j0, N, Pk, nparts_prev, nrhs, Dk, pf, kbn, sleep_ns);
            else diag_solve_kernel<false><<<nrhs, SUB * NW, DIAG_SMEM, sa>>>(kb, A + j0 + j0 * lld, lld, Xg + j0, acc + j0, N, Pk, nparts_prev, nrhs, Dk, pf, kbn, sleep_ns);
            SLB_CUDA(cudaGetLastError()); counter_add("kernel_launches", 1);
            SLB_CUDA(cudaEventRecord(evd[q], sa));
            if (!have_next) break;
            // ---- top(k): x_k into the rows of the next block, K-split partial sums ----
            double *Pn = part + (size_t)((q + 1) & 1) * KSPLIT * MAXRHS * nb;
            launch_gemv_rows(kbn, kb, A + jn + j0 * lld, lld, Xg + j0, N, Pn, 0, nrhs, true, sa);
            { const int kslice = (kb + KSPLIT - 1) / KSPLIT; nparts_prev = (kb + kslice - 1) / kslice; }
            // ---- bulk(k): x_k into the rows beyond the next block ----
            SLB_CUDA(cudaStreamWaitEvent(sb, evd[q], 0));
            if (fwd) {
                const int64_t r0 = jn + kbn;
                launch_gemv_rows(N - r0, kb, A + r0 + j0 * lld, lld, Xg + j0, N, acc + r0, N, nrhs, false, sb, scratch, counters);
            } else {
                launch_gemv_rows(jn, kb, A + j0 * lld, lld, Xg + j0, N, acc, N, nrhs, false, sb, scratch, counters);
            }
            SLB_CUDA(cudaEventRecord(evb[q], sb));
        }
        // the next pass clears acc: every bulk kernel of this one must be done
        SLB_CUDA(cudaEventRecord(join, sb));
        SLB_CUDA(cudaStreamWaitEvent(sa, join, 0));
    }
    SLB_CUDA(cudaEventRecord(join, sa));
    SLB_CUDA(cudaStreamWaitEvent(sb, join, 0));
}

// Xg: N x nrhs (ld = N), holds P b on entry and x on return.  The ~4 launches per block step are recorded ONCE into a CUDA
// graph (stream capture of the two-stream fork/join) and replayed: issued one by one from the host they cost more host time
// (~30 us per block step) than the device needs to execute them.  The graph is kept for the next solve with the same factors
// (same pointers and sizes: PDGETRS is typically called again and again on one factorisation).
static void diag_attr()
{
    static bool done = false;
    if (done) return;
    SLB_CUDA(cudaFuncSetAttribute(diag_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
    SLB_CUDA(cudaFuncSetAttribute(diag_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
    // the kernels that alternate with diag() on the same SMs ask for the same shared-memory carve-out: no L1 / shared
    // reconfiguration between consecutive launches of the critical chain
    if (opt("solve_carveout", 0)) {      // measured: no gain (profiles/r02_solve.md)
        SLB_CUDA(cudaFuncSetAttribute(gemv_rows_kernel<true, true, 1, 16>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        SLB_CUDA(cudaFuncSetAttribute(gemv_bulk_kernel<true, 1, 16>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        SLB_CUDA(cudaFuncSetAttribute(inv32_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    done = true;
}

void getrs_fast_device(int N, int nrhs, const double *A, int64_t lld, int nb, double *Xg)
{
    Runtime &r = rt();
    diag_attr();
    cudaStream_t sa = r.s_panel, sb = r.s_main;      // critical chain on the high-priority stream, the HBM stream on the low one
    const int nblk = (N + nb - 1) / nb;
    double *acc = (double *)workspace("rs_fast_acc", (size_t)N * nrhs * sizeof(double));
    double *part = (double *)workspace("rs_fast_part", (size_t)2 * KSPLIT * MAXRHS * nb * sizeof(double));
    double *Dinv = (double *)workspace("rs_fast_dinv", (size_t)2 * nblk * MAXSUB * SUB * SUB * sizeof(double));
    double *scratch = (double *)workspace("rs_fast_scratch", (size_t)BULK_MAXS * MAXRHS * (N + 2) * sizeof(double));
    unsigned *counters = (unsigned *)workspace("rs_fast_counters", (size_t)(N / 256 + 2) * sizeof(unsigned), true);
    static std::vector<cudaEvent_t> evd, evb;
    static cudaEvent_t fork = nullptr, join = nullptr;
    if (!fork) { SLB_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming)); SLB_CUDA(cudaEventCreateWithFlags(&join, cudaEventDisableTiming)); }
    while ((int)evd.size() < nblk + 2) {
        cudaEvent_t e; SLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); evd.push_back(e);
        SLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); evb.push_back(e);
    }
    if (opt("solve_graph", 1) == 0) { enqueue_sweeps(N, nrhs, A, lld, nb, Xg, acc, part, Dinv, scratch, counters, sa, sb, evd, evb, fork, join); return; }
    struct Key { int64_t N, nrhs, nb, lld; const void *A, *Xg, *acc, *part, *Dinv, *scratch, *counters; };      // no padding: compared with memcmp
    static Key key{}; static cudaGraphExec_t exec = nullptr;
    const Key now{ N, nrhs, nb, lld, A, Xg, acc, part, Dinv, scratch, counters };
    if (exec == nullptr || memcmp(&key, &now, sizeof(Key)) != 0) {
        if (exec) { SLB_CUDA(cudaGraphExecDestroy(exec)); exec = nullptr; }
        cudaGraph_t graph = nullptr;
        SLB_CUDA(cudaStreamBeginCapture(sb, cudaStreamCaptureModeRelaxed));
        enqueue_sweeps(N, nrhs, A, lld, nb, Xg, acc, part, Dinv, scratch, counters, sa, sb, evd, evb, fork, join);
        SLB_CUDA(cudaStreamEndCapture(sb, &graph));
        SLB_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        SLB_CUDA(cudaGraphDestroy(graph));
        key = now;
        counter_add("solve_graph_builds", 1);
    }
    SLB_CUDA(cudaGraphLaunch(exec, sb));
}


__global__ void lat_probe_kernel(long long *out, double seed)
{
    double v = seed, w = seed + 1.0; const double a = 1.0000001, b = 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) v = fma(v, a, b);
    }
    long long t1 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) w = w + b;
    }
    long long t2 = clock64();
    double s = v;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) s = __shfl_sync(0xffffffffu, s, (u + 1) & 31);
    }
    long long t3 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = (long long)(v + w + s); }
}

// test / tuning hook: `reps` back-to-back launches of one kernel of the sweep on the factors' first diagonal block (which: 0 = diag
// forward, 1 = diag backward, 2 = top, 3 = bulk over nr rows); returns the mean time per launch in microseconds
double solve_fast_probe(int which, int nb, int64_t nr, const double *A, int64_t lld, int N, int reps)
{
    Runtime &r = rt(); cudaStream_t s = r.s_main;
    diag_attr();
    const int nblk = (N + nb - 1) / nb;
    double *acc = (double *)workspace("rs_fast_acc", (size_t)N * sizeof(double));
    double *part = (double *)workspace("rs_fast_part", (size_t)2 * KSPLIT * MAXRHS * nb * sizeof(double));
    double *Dinv = (double *)workspace("rs_fast_dinv", (size_t)2 * nblk * MAXSUB * SUB * SUB * sizeof(double));
    double *scratch = (double *)workspace("rs_fast_scratch", (size_t)BULK_MAXS * MAXRHS * (N + 2) * sizeof(double));
    unsigned *counters = (unsigned *)workspace("rs_fast_counters", (size_t)(N / 256 + 2) * sizeof(unsigned), true);
    double *X = (double *)workspace("rs_Xg", (size_t)N * sizeof(double));
    SLB_CUDA(cudaMemsetAsync(X, 0, (size_t)N * sizeof(double), s)); SLB_CUDA(cudaMemsetAsync(acc, 0, (size_t)N * sizeof(double), s));
    SLB_CUDA(cudaMemsetAsync(part, 0, (size_t)2 * KSPLIT * MAXRHS * nb * sizeof(double), s));
    inv32_kernel<<<(unsigned)((2 * nblk * MAXSUB + 3) / 4), 128, 0, s>>>(N, nb, nblk, A, lld, Dinv);
    cudaEvent_t e0, e1; SLB_CUDA(cudaEventCreate(&e0)); SLB_CUDA(cudaEventCreate(&e1));
    if (which == 4) {       // per-warp timeline of one diag launch (clock64 ticks since kernel start)
        long long *dbg = (long long *)workspace("rs_fast_dbg", MAXSUB * 8 * sizeof(long long), true);
        for (int it = 0; it < 3; ++it)
            diag_solve_kernel<true><<<1, SUB * NW, DIAG_SMEM, s>>>(nb, A, lld, X, acc, N, part, KSPLIT, 1, Dinv, A + nb + (int64_t)nb * lld, nb, (unsigned)opt("solve_poll_ns", 40), dbg);
        long long h[MAXSUB * 8];
        SLB_CUDA(cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, s)); SLB_CUDA(cudaStreamSynchronize(s));
        lat_probe_kernel<<<1, 32, 0, s>>>(dbg, 1.5);
        long long l[4]; SLB_CUDA(cudaMemcpyAsync(l, dbg, sizeof(l), cudaMemcpyDeviceToHost, s)); SLB_CUDA(cudaStreamSynchronize(s));
        fprintf(stderr, "diag latency probe (1 warp, dependent chains of 1024): DFMA %.1f DADD %.1f SHFL.64 %.1f ticks per op\n", l[0] / 1024.0, l[1] / 1024.0, l[2] / 1024.0);
        for (int q = 0; q < MAXSUB; ++q) fprintf(stderr, "diag timeline q=%2d: start %6lld rhs %6lld | flag seen %6lld synced %6lld folds %6lld | tri done %6lld synced %6lld published %6lld\n", q, h[q * 8], h[q * 8 + 1], h[q * 8 + 4], h[q * 8 + 5], h[q * 8 + 2], h[q * 8 + 6], h[q * 8 + 7], h[q * 8 + 3]);
        return 0.0;
    }
    for (int it = 0; it < reps + 3; ++it) {
        if (it == 3) SLB_CUDA(cudaEventRecord(e0, s));
        if (which == 0) diag_solve_kernel<true><<<1, SUB * NW, DIAG_SMEM, s>>>(nb, A, lld, X, acc, N, part, KSPLIT, 1, Dinv, A + nb + (int64_t)nb * lld, nb, (unsigned)opt("solve_poll_ns", 40));
        else if (which == 1) diag_solve_kernel<false><<<1, SUB * NW, DIAG_SMEM, s>>>(nb, A, lld, X, acc, N, part, KSPLIT, 1, Dinv + (size_t)nblk * MAXSUB * SUB * SUB, A + nb + (int64_t)nb * lld, nb, (unsigned)opt("solve_poll_ns", 40));
        else if (which == 2) launch_gemv_rows(nb, nb, A + nb, lld, X, N, part, 0, 1, true, s);
        else launch_gemv_rows(nr, nb, A + (N - nr), lld, X, N, acc + (N - nr), N, 1, false, s, scratch, counters);
    }
    SLB_CUDA(cudaEventRecord(e1, s));
    SLB_CUDA(cudaStreamSynchronize(s));
    float ms = 0; SLB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return (double)ms * 1e3 / reps;
}

}  // namespace slb
