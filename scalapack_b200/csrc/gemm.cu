// gemm.cu -- C[MxN] -= A[MxK] * B[KxN] in FP64 on the tensor cores, cp.async flavour (operands read in place).
//
// The trailing update of the LU step (the reference's one local dgemm_ per step, PBLAS/SRC/PTOOLS/PB_CpgemmAB.c:345
// reached from SRC/pdgetrf.f:288; >= 99 % of the flops) runs in gemm_packed.cu once its operands are worth packing
// (M >= SLB200_GEMM_PACKED_MIN); this file is the kernel for everything smaller -- the DMMA updates inside the panel
// recursion (panel.cu) and the U12 solve (trsm.cu), and short trailing updates -- plus the complex kernel.
//
// sm_100a has no FP64 kind of tcgen05.mma (kinds: f16, tf32, f8f6f4, i8, mxf8f6f4, mxf4, mxf4nvf4), so the FP64 tensor
// path is the warp-level DMMA  mma.sync.aligned.m16n8k4.row.col.f64  with register accumulators
// (profiles/r01_dmma_probe.md).  Tiling: CTA 128(m) x 128(n) x 16(k), 8 warps as 2(m) x 4(n), warp tile 64 x 32.
// The MMA is issued "transposed" (MMA-M runs along the problem's n, MMA-N along m) so that each thread's
// accumulator pair (c0,c1) is two consecutive rows of one column of column-major C: the epilogue is
// 16-byte read-modify-writes, 64 B contiguous per column per warp request.
// Operands are staged by a 4-stage cp.async (LDGSTS) ring into padded shared memory:
//   As[k][m] stride 132 doubles, Bs[n][k] stride 20 doubles -> every fragment LDS.64 is bank-conflict free.
// Persistent CTAs (one per SM) walk the tiles in groups of 16 m-tiles swept along n so a group's A rows stay
// L2-resident while B streams; the ring runs across tile boundaries.  History of the variants that led here
// (per-tile kernel 70 % DMMA-pipe busy -> persistent 16-warp 73 % -> this one 83 % -> packed 96 %): profiles/r01_gemm_v*_ncu.md.
#include "kernels.cuh"
#include "common.h"

namespace slb {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, NTHREADS = 256;
constexpr int SA = BM + 4;            // As row stride (doubles)
constexpr int SB = BK + 4;            // Bs row stride (doubles)
constexpr int AS_STAGE = BK * SA;     // doubles per stage
constexpr int BS_STAGE = BN * SB;
constexpr int GROUP_M = 16;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, int src_bytes)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// D(16x8) += A(16x4) * B(4x8), FP64 tensor core
__device__ __forceinline__ void dmma_16x8x4(double (&d)[4], double a0, double a1, double b0)
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a0), "d"(a1), "d"(b0));
}

// Stage loader.  VEC=2: 16-byte chunks (needs 16 B aligned base pointers and even lda/ldb); VEC=1: 8-byte.
template <int VEC>
__device__ __forceinline__ void load_stage(double *As, double *Bs, const double *__restrict__ A, int64_t lda,
                                           const double *__restrict__ B, int64_t ldb, int64_t m0, int64_t n0, int k0,
                                           int64_t M, int64_t N, int K, int tid)
{
    if (VEC == 2) {
#pragma unroll
        for (int i = 0; i < (BK * BM / 2) / NTHREADS; ++i) {       // A: 16 k-rows x 64 chunks
            int c = tid + i * NTHREADS;
            int k = c >> 6, mc = (c & 63) * 2;
            int64_t m = m0 + mc;
            int kk = k0 + k;
            int bytes = 0;
            if (kk < K && m < M) bytes = (M - m >= 2) ? 16 : 8;
            const double *src = bytes ? (A + m + (int64_t)kk * lda) : A;
            cp_async16(As + k * SA + mc, src, bytes);
        }
#pragma unroll
        for (int i = 0; i < (BN * BK / 2) / NTHREADS; ++i) {       // B: 128 n-cols x 8 chunks
            int c = tid + i * NTHREADS;
            int n = c >> 3, kc = (c & 7) * 2;
            int64_t nn = n0 + n;
            int kk = k0 + kc;
            int bytes = 0;
            if (nn < N && kk < K) bytes = (K - kk >= 2) ? 16 : 8;
            const double *src = bytes ? (B + kk + nn * ldb) : B;
            cp_async16(Bs + n * SB + kc, src, bytes);
        }
    } else {
#pragma unroll
        for (int i = 0; i < (BK * BM) / NTHREADS; ++i) {
            int c = tid + i * NTHREADS;
            int k = c >> 7, mc = c & 127;
            int64_t m = m0 + mc;
            int kk = k0 + k;
            int bytes = (kk < K && m < M) ? 8 : 0;
            const double *src = bytes ? (A + m + (int64_t)kk * lda) : A;
            cp_async8(As + k * SA + mc, src, bytes);
        }
#pragma unroll
        for (int i = 0; i < (BN * BK) / NTHREADS; ++i) {
            int c = tid + i * NTHREADS;
            int n = c >> 4, kc = c & 15;
            int64_t nn = n0 + n;
            int kk = k0 + kc;
            int bytes = (nn < N && kk < K) ? 8 : 0;
            const double *src = bytes ? (B + kk + nn * ldb) : B;
            cp_async8(Bs + n * SB + kc, src, bytes);
        }
    }
}

__device__ __forceinline__ void tile_coords(int64_t t, int tiles_m, int tiles_n, int64_t &m0, int64_t &n0)
{
    int group_sz = GROUP_M * tiles_n;
    int grp = (int)(t / group_sz);
    int first_m = grp * GROUP_M;
    int gm = min(GROUP_M, tiles_m - first_m);
    int r = (int)(t % group_sz);
    m0 = (int64_t)(first_m + r % gm) * BM;
    n0 = (int64_t)(r / gm) * BN;
}

// The per-stage block barrier sits in the MIDDLE of the stage: at the stage boundary the fragments of the next stage
// are already being prefetched (its data became visible at the mid-stage barrier), so the DMMA pipe never drains
// waiting for LDS after a barrier.  chunk > 0: a CTA processes `chunk` consecutive tiles and retires.
template <int VEC>
__global__ void __launch_bounds__(NTHREADS, 1)
dgemm_minus_p8b(int64_t M, int64_t N, int K, const double *__restrict__ A, int64_t lda, const double *__restrict__ B,
               int64_t ldb, double *__restrict__ C, int64_t ldc, int tiles_m, int tiles_n, int chunk)
{
    extern __shared__ __align__(16) double smem[];
    double *As = smem;
    double *Bs = smem + STAGES * AS_STAGE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int wm0 = (warp & 1) * 64, wn0 = (warp >> 1) * 32;      // 2 x 4 warps, 64 x 32 each

    const int64_t ntiles = (int64_t)tiles_m * tiles_n;
    const int KT = (K + BK - 1) / BK;
    const int64_t t_first = chunk > 0 ? (int64_t)blockIdx.x * chunk : blockIdx.x;
    const int64_t t_stride = chunk > 0 ? 1 : gridDim.x;
    const int64_t my_tiles = chunk > 0 ? max((int64_t)0, min((int64_t)chunk, ntiles - t_first)) : (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int64_t total = my_tiles * KT;

    int64_t l_lt = 0; int l_kt = 0; int64_t l_m0 = 0, l_n0 = 0;
    if (my_tiles > 0) tile_coords(t_first, tiles_m, tiles_n, l_m0, l_n0);
    auto issue_load = [&](int64_t li) {
        if (li < total) {
            int st = (int)(li % STAGES);
            load_stage<VEC>(As + st * AS_STAGE, Bs + st * BS_STAGE, A, lda, B, ldb, l_m0, l_n0, l_kt * BK, M, N, K, tid);
            if (++l_kt == KT) {
                l_kt = 0; ++l_lt;
                if (l_lt < my_tiles) tile_coords(t_first + l_lt * t_stride, tiles_m, tiles_n, l_m0, l_n0);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue_load(s);

    double acc[2][8][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[i][j][v] = 0.0;

    double fa[2][2][2], fb[2][8];
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (total > 0) {
#pragma unroll
        for (int nf = 0; nf < 2; ++nf) {
            fa[0][nf][0] = Bs[(wn0 + nf * 16 + g) * SB + tig];
            fa[0][nf][1] = Bs[(wn0 + nf * 16 + g + 8) * SB + tig];
        }
#pragma unroll
        for (int mf = 0; mf < 8; ++mf) fb[0][mf] = As[tig * SA + wm0 + mf * 8 + g];
    }
    int kt = 0; int64_t lt = 0; int64_t m0 = 0, n0 = 0;
    for (int64_t ci = 0; ci < total; ++ci) {
        if (kt == 0) {
            tile_coords(t_first + lt * t_stride, tiles_m, tiles_n, m0, n0);
#pragma unroll
            for (int nf = 0; nf < 2; ++nf)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int mq = 0; mq < 4; ++mq) {
                        int64_t n = n0 + wn0 + nf * 16 + g + h * 8;
                        int64_t m = m0 + wm0 + mq * 16 + 2 * tig;
                        if (n < N && m < M) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(C + m + n * ldc));
                    }
        }
        const double *as = As + (ci % STAGES) * AS_STAGE;
        const double *bs = Bs + (ci % STAGES) * BS_STAGE;
        const double *asn = As + ((ci + 1) % STAGES) * AS_STAGE;
        const double *bsn = Bs + ((ci + 1) % STAGES) * BS_STAGE;
#pragma unroll
        for (int s4 = 0; s4 < BK / 4; ++s4) {
            if (s4 == 1) {
                cp_async_wait<STAGES - 3>();       // stage ci+1 has landed (this thread's copies) ...
                __syncthreads();                   // ... for every thread; and everyone is done with stage ci-1
                issue_load(ci + STAGES - 1);       // refill the buffer of stage ci-1
            }
            const int cur = s4 & 1, nxt = cur ^ 1;
            if (s4 + 1 < BK / 4) {
#pragma unroll
                for (int nf = 0; nf < 2; ++nf) {
                    fa[nxt][nf][0] = bs[(wn0 + nf * 16 + g) * SB + (s4 + 1) * 4 + tig];
                    fa[nxt][nf][1] = bs[(wn0 + nf * 16 + g + 8) * SB + (s4 + 1) * 4 + tig];
                }
#pragma unroll
                for (int mf = 0; mf < 8; ++mf) fb[nxt][mf] = as[((s4 + 1) * 4 + tig) * SA + wm0 + mf * 8 + g];
            } else if (ci + 1 < total) {
#pragma unroll
                for (int nf = 0; nf < 2; ++nf) {
                    fa[nxt][nf][0] = bsn[(wn0 + nf * 16 + g) * SB + tig];
                    fa[nxt][nf][1] = bsn[(wn0 + nf * 16 + g + 8) * SB + tig];
                }
#pragma unroll
                for (int mf = 0; mf < 8; ++mf) fb[nxt][mf] = asn[tig * SA + wm0 + mf * 8 + g];
            }
#pragma unroll
            for (int nf = 0; nf < 2; ++nf)
#pragma unroll
                for (int mf = 0; mf < 8; ++mf) dmma_16x8x4(acc[nf][mf], fa[cur][nf][0], fa[cur][nf][1], fb[cur][mf]);
        }
        if (++kt == KT) {
            kt = 0; ++lt;
#pragma unroll
            for (int nf = 0; nf < 2; ++nf) {
#pragma unroll
                for (int mh = 0; mh < 2; ++mh) {                 // batches of 8 x 16 B
                    double2 cv[2][4];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        int64_t n = n0 + wn0 + nf * 16 + g + h * 8;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            int64_t m = m0 + wm0 + (mh * 4 + q) * 8 + 2 * tig;
                            if (VEC == 2 && n < N && m + 1 < M) cv[h][q] = *reinterpret_cast<const double2 *>(C + m + n * ldc);
                            else if (n < N && m < M) cv[h][q] = make_double2(C[m + n * ldc], (m + 1 < M) ? C[m + 1 + n * ldc] : 0.0);
                            else cv[h][q] = make_double2(0.0, 0.0);
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        int64_t n = n0 + wn0 + nf * 16 + g + h * 8;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            int64_t m = m0 + wm0 + (mh * 4 + q) * 8 + 2 * tig;
                            double2 c = cv[h][q];
                            c.x -= acc[nf][mh * 4 + q][2 * h]; c.y -= acc[nf][mh * 4 + q][2 * h + 1];
                            if (VEC == 2 && n < N && m + 1 < M) *reinterpret_cast<double2 *>(C + m + n * ldc) = c;
                            else if (n < N && m < M) { C[m + n * ldc] = c.x; if (m + 1 < M) C[m + 1 + n * ldc] = c.y; }
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int v = 0; v < 4; ++v) acc[i][j][v] = 0.0;
        }
    }
    cp_async_wait<0>();
}

// ---- complex: C -= A*B with interleaved (re,im); four real DMMAs per complex MMA ---------------------
// Tiling: CTA 64(m) x 64(n) x 16(k) complex, 8 warps as 2(m) x 4(n), warp tile 32 x 16.
constexpr int ZBM = 64, ZBN = 64, ZBK = 16;
constexpr int ZSA = ZBM + 2;            // complex elements; 16 B each -> stride 66*16 B
constexpr int ZSB = ZBK + 4;            // 16-byte banks: (g*4 + tig) mod 8 distinct over a quarter warp
constexpr int ZAS_STAGE = ZBK * ZSA, ZBS_STAGE = ZBN * ZSB;
constexpr int ZSTAGES = 3;

__device__ __forceinline__ void zload_stage(zcomplex *As, zcomplex *Bs, const zcomplex *__restrict__ A, int64_t lda,
                                            const zcomplex *__restrict__ B, int64_t ldb, int64_t m0, int64_t n0, int k0,
                                            int64_t M, int64_t N, int K, int tid)
{
#pragma unroll
    for (int i = 0; i < (ZBK * ZBM) / NTHREADS; ++i) {
        int c = tid + i * NTHREADS;
        int k = c >> 6, mc = c & 63;
        int64_t m = m0 + mc; int kk = k0 + k;
        int bytes = (kk < K && m < M) ? 16 : 0;
        const zcomplex *src = bytes ? (A + m + (int64_t)kk * lda) : A;
        cp_async16(As + k * ZSA + mc, src, bytes);
    }
#pragma unroll
    for (int i = 0; i < (ZBN * ZBK) / NTHREADS; ++i) {
        int c = tid + i * NTHREADS;
        int n = c >> 4, kc = c & 15;
        int64_t nn = n0 + n; int kk = k0 + kc;
        int bytes = (nn < N && kk < K) ? 16 : 0;
        const zcomplex *src = bytes ? (B + kk + nn * ldb) : B;
        cp_async16(Bs + n * ZSB + kc, src, bytes);
    }
}

__global__ void __launch_bounds__(NTHREADS, 1)
zgemm_minus_kernel(int64_t M, int64_t N, int K, const zcomplex *__restrict__ A, int64_t lda, const zcomplex *__restrict__ B,
                   int64_t ldb, zcomplex *__restrict__ C, int64_t ldc, int tiles_m, int tiles_n)
{
    extern __shared__ __align__(16) double smem[];
    zcomplex *As = reinterpret_cast<zcomplex *>(smem);
    zcomplex *Bs = As + ZSTAGES * ZAS_STAGE;

    int64_t t = blockIdx.x;
    int group_sz = GROUP_M * tiles_n;
    int grp = (int)(t / group_sz);
    int first_m = grp * GROUP_M;
    int gm = min(GROUP_M, tiles_m - first_m);
    int r = (int)(t % group_sz);
    int tm = first_m + r % gm, tn = r / gm;
    const int64_t m0 = (int64_t)tm * ZBM, n0 = (int64_t)tn * ZBN;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int wm0 = (warp & 1) * 32, wn0 = (warp >> 1) * 16;

    double accr[4][4], acci[4][4];       // one n-frag (16) x four m-frags (8)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int v = 0; v < 4; ++v) { accr[j][v] = 0.0; acci[j][v] = 0.0; }

    const int KT = (K + ZBK - 1) / ZBK;
#pragma unroll
    for (int s = 0; s < ZSTAGES - 1; ++s) {
        if (s < KT) zload_stage(As + s * ZAS_STAGE, Bs + s * ZBS_STAGE, A, lda, B, ldb, m0, n0, s * ZBK, M, N, K, tid);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<ZSTAGES - 2>();
        __syncthreads();
        {
            int nk = kt + ZSTAGES - 1;
            if (nk < KT) { int st = nk % ZSTAGES; zload_stage(As + st * ZAS_STAGE, Bs + st * ZBS_STAGE, A, lda, B, ldb, m0, n0, nk * ZBK, M, N, K, tid); }
            cp_async_commit();
        }
        const zcomplex *as = As + (kt % ZSTAGES) * ZAS_STAGE;
        const zcomplex *bs = Bs + (kt % ZSTAGES) * ZBS_STAGE;
#pragma unroll
        for (int k4 = 0; k4 < ZBK; k4 += 4) {
            zcomplex b0 = bs[(wn0 + g) * ZSB + k4 + tig];          // MMA-A operand rows n
            zcomplex b1 = bs[(wn0 + g + 8) * ZSB + k4 + tig];
#pragma unroll
            for (int mf = 0; mf < 4; ++mf) {
                zcomplex a = as[(k4 + tig) * ZSA + wm0 + mf * 8 + g];
                // (br + i bi)(ar + i ai): real += br*ar - bi*ai ; imag += br*ai + bi*ar
                dmma_16x8x4(accr[mf], b0.x, b1.x, a.x);
                dmma_16x8x4(accr[mf], -b0.y, -b1.y, a.y);
                dmma_16x8x4(acci[mf], b0.x, b1.x, a.y);
                dmma_16x8x4(acci[mf], b0.y, b1.y, a.x);
            }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int64_t n = n0 + wn0 + g + h * 8;
        if (n >= N) continue;
#pragma unroll
        for (int mf = 0; mf < 4; ++mf) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int64_t m = m0 + wm0 + mf * 8 + 2 * tig + e;
                if (m >= M) continue;
                zcomplex *p = C + m + n * ldc;
                zcomplex c = *p;
                c.x -= accr[mf][2 * h + e]; c.y -= acci[mf][2 * h + e];
                *p = c;
            }
        }
    }
}

}  // namespace

bool dgemm_takes_packed(int64_t M, int K, int flags)
{
    // SLB200_GEMM_VARIANT=7: cp.async kernel only (options are read on every call so slb200_set_option takes effect)
    return (flags & GEMM_MAIN) && K >= 16 && opt("gemm_variant", 9) == 9 && M >= opt("gemm_packed_min", 3072);
}

void launch_dgemm_minus(int64_t M, int64_t N, int K, const double *A, int64_t lda, const double *B, int64_t ldb,
                        double *C, int64_t ldc, cudaStream_t s, int chunk, int flags)
{
    if (M <= 0 || N <= 0 || K <= 0) return;
    if (dgemm_takes_packed(M, K, flags)) {
        launch_dgemm_minus_packed(M, N, K, A, lda, B, ldb, C, ldc, s, chunk, (flags & GEMM_REUSE_A) != 0);
        return;
    }
    constexpr size_t smem_bytes = (size_t)STAGES * (AS_STAGE + BS_STAGE) * sizeof(double);
    static bool attr_done = false;
    if (!attr_done) {
        SLB_CUDA(cudaFuncSetAttribute(dgemm_minus_p8b<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        SLB_CUDA(cudaFuncSetAttribute(dgemm_minus_p8b<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        attr_done = true;
    }
    const int tiles_m = (int)((M + BM - 1) / BM), tiles_n = (int)((N + BN - 1) / BN);
    const int64_t ntiles = (int64_t)tiles_m * tiles_n;
    if (ntiles > 0x7fffffffLL) fatal("dgemm: too many tiles");
    const bool aligned = (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) == 0 && (lda % 2 == 0) && (ldb % 2 == 0) && (ldc % 2 == 0);
    unsigned grid = (unsigned)(ntiles < rt().sm_count ? ntiles : rt().sm_count);
    if (chunk > 0 && ntiles > rt().sm_count) grid = (unsigned)((ntiles + chunk - 1) / chunk); else chunk = 0;
    if (aligned)
        dgemm_minus_p8b<2><<<grid, NTHREADS, smem_bytes, s>>>(M, N, K, A, lda, B, ldb, C, ldc, tiles_m, tiles_n, chunk);
    else
        dgemm_minus_p8b<1><<<grid, NTHREADS, smem_bytes, s>>>(M, N, K, A, lda, B, ldb, C, ldc, tiles_m, tiles_n, chunk);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
    counter_add("gemm_launches", 1);
}

bool zgemm_takes_packed(int64_t M, int K, int flags)
{
    return (flags & GEMM_MAIN) && K >= 16 && opt("zgemm_packed", 1) != 0 && M >= opt("zgemm_packed_min", 2048);
}

void launch_zgemm_minus(int64_t M, int64_t N, int K, const zcomplex *A, int64_t lda, const zcomplex *B, int64_t ldb,
                        zcomplex *C, int64_t ldc, cudaStream_t s, int chunk, int flags)
{
    if (M <= 0 || N <= 0 || K <= 0) return;
    if (zgemm_takes_packed(M, K, flags)) {
        launch_zgemm_minus_packed(M, N, K, A, lda, B, ldb, C, ldc, s, chunk, (flags & GEMM_REUSE_A) != 0);
        return;
    }
    static bool attr_done = false;
    const size_t smem_bytes = (size_t)ZSTAGES * (ZAS_STAGE + ZBS_STAGE) * sizeof(zcomplex);
    if (!attr_done) {
        SLB_CUDA(cudaFuncSetAttribute(zgemm_minus_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        attr_done = true;
    }
    int tiles_m = (int)((M + ZBM - 1) / ZBM), tiles_n = (int)((N + ZBN - 1) / ZBN);
    int64_t ntiles = (int64_t)tiles_m * tiles_n;
    if (ntiles > 0x7fffffffLL) fatal("zgemm: too many tiles");
    zgemm_minus_kernel<<<(unsigned)ntiles, NTHREADS, smem_bytes, s>>>(M, N, K, A, lda, B, ldb, C, ldc, tiles_m, tiles_n);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
    counter_add("gemm_launches", 1);
}

}  // namespace slb
