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

__device__ __forceinline__ int ld_flag(volatile int *p) { return *p; }

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
template <bool FORWARD>
__global__ void __launch_bounds__(SUB * NW, 1)
diag_solve_kernel(int kb, const double *__restrict__ A, int64_t lda, double *__restrict__ Xk, const double *__restrict__ Acc,
                  int64_t ldx, const double *__restrict__ P, int nparts, int nrhs, const double *__restrict__ Dk,
                  const double *__restrict__ pf, int pf_cols)
{
    __shared__ double xs[SUB * MAXSUB];
    __shared__ int flag[MAXSUB];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nsub = (kb + SUB - 1) / SUB;
    const int rhs = blockIdx.x;
    // the next diagonal block -> L2 while this one is being solved (pf_cols columns of pf_cols doubles, 128-byte lines)
    if (pf != nullptr) {
        const int lines = (pf_cols * 8 + 127) / 128;
        for (int e = blockIdx.x * blockDim.x + tid; e < pf_cols * lines; e += gridDim.x * blockDim.x)
            asm volatile("prefetch.global.L2 [%0];\n" ::"l"((const char *)(pf + (int64_t)(e / lines) * lda) + (e % lines) * 128));
    }
    if (tid < MAXSUB) flag[tid] = 0;
    __syncthreads();
    for (int q = w; q < nsub; q += NW) {
        const int sq = FORWARD ? q : nsub - 1 - q;                       // sub-block q of the sweep
        const int sb = sq * SUB;
        const int i = sb + lane;
        const bool valid = i < kb;
        double dinv[SUB];                                                // my row of the inverted diagonal triangle
#pragma unroll
        for (int c = 0; c < SUB; ++c) dinv[c] = Dk[(size_t)sq * (SUB * SUB) + c * SUB + lane];
        double v = 0.0;
        if (valid) {
            v = Xk[i + (int64_t)rhs * ldx] + Acc[i + (int64_t)rhs * ldx];
            for (int s = 0; s < nparts; ++s) v += P[((int64_t)s * nrhs + rhs) * kb + i];
        }
        // fold the sub-blocks solved before mine, in sweep order, as their x appears; the loads run one sub-block ahead
        auto load_blk = [&](double (&a)[SUB], int p) {
            const int pb = (FORWARD ? p : nsub - 1 - p) * SUB;
            const int pbs = min(SUB, kb - pb);
#pragma unroll
            for (int k = 0; k < SUB; ++k) a[k] = (valid && k < pbs) ? A[i + (int64_t)(pb + k) * lda] : 0.0;
        };
        auto fold = [&](const double (&a)[SUB], int p) {
            const int pb = (FORWARD ? p : nsub - 1 - p) * SUB;
            while (ld_flag(&flag[p]) == 0) { }
            __syncwarp();
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int k = 0; k < SUB; k += 4) {
                s0 = fma(a[k], xs[pb + k], s0); s1 = fma(a[k + 1], xs[pb + k + 1], s1);
                s2 = fma(a[k + 2], xs[pb + k + 2], s2); s3 = fma(a[k + 3], xs[pb + k + 3], s3);
            }
            v -= (s0 + s1) + (s2 + s3);
        };
        double a0[SUB], a1[SUB];
        if (q > 0) load_blk(a0, 0);
        for (int p = 0; p < q; p += 2) {
            if (p + 1 < q) load_blk(a1, p + 1);
            fold(a0, p);
            if (p + 1 < q) {
                if (p + 2 < q) load_blk(a0, p + 2);
                fold(a1, p + 1);
            }
        }
        // my triangle: x = inv(T) v, the 32 values of v broadcast by shuffles, four independent chains
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int c = 0; c < SUB; c += 4) {
            s0 = fma(dinv[c], __shfl_sync(0xffffffffu, v, c), s0); s1 = fma(dinv[c + 1], __shfl_sync(0xffffffffu, v, c + 1), s1);
            s2 = fma(dinv[c + 2], __shfl_sync(0xffffffffu, v, c + 2), s2); s3 = fma(dinv[c + 3], __shfl_sync(0xffffffffu, v, c + 3), s3);
        }
        v = (s0 + s1) + (s2 + s3);
        if (valid) { xs[i] = v; Xk[i + (int64_t)rhs * ldx] = v; }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) *((volatile int *)&flag[q]) = 1;
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

template <int NR>
void launch_gemv_rows_nr(int64_t nr, int kb, const double *A, int64_t lda, const double *X, int64_t ldx, double *Y, int64_t ldy, int nrhs,
                         bool partial, cudaStream_t s)
{
    constexpr int UNR = NR <= 2 ? 16 : 8;                              // 16-byte loads in flight per thread
    const bool vec = (((uintptr_t)A) & 15) == 0 && lda % 2 == 0 && nr % 2 == 0;
    // enough CTAs to fill the GPU also when few rows are left: 256, 128 or 64 rows per CTA
    const int sms = rt().sm_count;
    const int threads = partial ? 128 : (nr >= (int64_t)sms * 512 ? 128 : (nr >= (int64_t)sms * 256 ? 64 : 32));
    const unsigned gx = (unsigned)((nr + 2 * threads - 1) / (2 * threads));
    if (partial) {
        const int kslice = (kb + KSPLIT - 1) / KSPLIT;
        dim3 grid(gx, (unsigned)((kb + kslice - 1) / kslice));
        if (vec) gemv_rows_kernel<true, true, NR, UNR><<<grid, threads, 0, s>>>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, kslice);
        else gemv_rows_kernel<false, true, NR, UNR><<<grid, threads, 0, s>>>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, kslice);
    } else {
        if (vec) gemv_rows_kernel<true, false, NR, UNR><<<gx, threads, 0, s>>>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, kb);
        else gemv_rows_kernel<false, false, NR, UNR><<<gx, threads, 0, s>>>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, kb);
    }
}
void launch_gemv_rows(int64_t nr, int kb, const double *A, int64_t lda, const double *X, int64_t ldx, double *Y, int64_t ldy, int nrhs,
                      bool partial, cudaStream_t s)
{
    if (nr <= 0 || kb <= 0) return;
    if (nrhs == 1) launch_gemv_rows_nr<1>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, partial, s);
    else if (nrhs == 2) launch_gemv_rows_nr<2>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, partial, s);
    else if (nrhs <= 4) launch_gemv_rows_nr<4>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, partial, s);
    else launch_gemv_rows_nr<8>(nr, kb, A, lda, X, ldx, Y, ldy, nrhs, partial, s);
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
                           cudaStream_t sa, cudaStream_t sb, std::vector<cudaEvent_t> &evd, std::vector<cudaEvent_t> &evb,
                           cudaEvent_t fork, cudaEvent_t join)
{
    const int nblk = (N + nb - 1) / nb;
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
            if (fwd) diag_solve_kernel<true><<<nrhs, SUB * NW, 0, sa>>>(kb, A + j0 + j0 * lld, lld, Xg + j0, acc + j0, N, Pk, nparts_prev, nrhs, Dk, pf, kbn);
            else diag_solve_kernel<false><<<nrhs, SUB * NW, 0, sa>>>(kb, A + j0 + j0 * lld, lld, Xg + j0, acc + j0, N, Pk, nparts_prev, nrhs, Dk, pf, kbn);
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
                launch_gemv_rows(N - r0, kb, A + r0 + j0 * lld, lld, Xg + j0, N, acc + r0, N, nrhs, false, sb);
            } else {
                launch_gemv_rows(jn, kb, A + j0 * lld, lld, Xg + j0, N, acc, N, nrhs, false, sb);
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
void getrs_fast_device(int N, int nrhs, const double *A, int64_t lld, int nb, double *Xg)
{
    Runtime &r = rt();
    cudaStream_t sa = r.s_panel, sb = r.s_main;      // critical chain on the high-priority stream, the HBM stream on the low one
    const int nblk = (N + nb - 1) / nb;
    double *acc = (double *)workspace("rs_fast_acc", (size_t)N * nrhs * sizeof(double));
    double *part = (double *)workspace("rs_fast_part", (size_t)2 * KSPLIT * MAXRHS * nb * sizeof(double));
    double *Dinv = (double *)workspace("rs_fast_dinv", (size_t)2 * nblk * MAXSUB * SUB * SUB * sizeof(double));
    static std::vector<cudaEvent_t> evd, evb;
    static cudaEvent_t fork = nullptr, join = nullptr;
    if (!fork) { SLB_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming)); SLB_CUDA(cudaEventCreateWithFlags(&join, cudaEventDisableTiming)); }
    while ((int)evd.size() < nblk + 2) {
        cudaEvent_t e; SLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); evd.push_back(e);
        SLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); evb.push_back(e);
    }
    if (opt("solve_graph", 1) == 0) { enqueue_sweeps(N, nrhs, A, lld, nb, Xg, acc, part, Dinv, sa, sb, evd, evb, fork, join); return; }
    struct Key { int64_t N, nrhs, nb, lld; const void *A, *Xg, *acc, *part, *Dinv; };      // no padding: compared with memcmp
    static Key key{}; static cudaGraphExec_t exec = nullptr;
    const Key now{ N, nrhs, nb, lld, A, Xg, acc, part, Dinv };
    if (exec == nullptr || memcmp(&key, &now, sizeof(Key)) != 0) {
        if (exec) { SLB_CUDA(cudaGraphExecDestroy(exec)); exec = nullptr; }
        cudaGraph_t graph = nullptr;
        SLB_CUDA(cudaStreamBeginCapture(sb, cudaStreamCaptureModeRelaxed));
        enqueue_sweeps(N, nrhs, A, lld, nb, Xg, acc, part, Dinv, sa, sb, evd, evb, fork, join);
        SLB_CUDA(cudaStreamEndCapture(sb, &graph));
        SLB_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        SLB_CUDA(cudaGraphDestroy(graph));
        key = now;
        counter_add("solve_graph_builds", 1);
    }
    SLB_CUDA(cudaGraphLaunch(exec, sb));
}

}  // namespace slb
