// gemm_packed.cu -- trailing-matrix update, generation 9: fragment-ordered operands + bulk-copy ring + mbarriers.
//
// Same contraction as gemm.cu (C[MxN] -= A[MxK] B[KxN], FP64 DMMA, 128 x 128 x 16 CTA tile, 8 consumer warps of
// 64 x 32), replaces the one local dgemm_ per LU step (PBLAS/SRC/PTOOLS/PB_CpgemmAB.c:345 from SRC/pdgetrf.f:288).
// What changed, and why (ncu on v7, profiles/r01_gemm_v7_ncu.md: DMMA pipe 83 % busy):
//   * 10 % of the issue samples were the 227 address/LDGSTS instructions every warp executes per k-stage, 5 % the
//     epilogue during which BOTH warps of a scheduler leave the pipe idle, plus one block barrier per stage.
//   * Here the L21 panel and the U12 block row are first re-ordered ("packed") by two HBM-bound kernels into
//     16 KB blocks (128 rows x 16 k) whose element order is exactly the DMMA fragment order of the consumer
//     warps.  A k-stage is then TWO cp.async.bulk copies issued by one producer thread (no per-thread address
//     math, no bounds checks: blocks are zero padded), every fragment load is a conflict-free LDS.128 at an
//     immediate offset, and stages are handed over with full/empty mbarriers instead of __syncthreads().
//   * Without block barriers the warps drift apart, so one warp's epilogue falls into the main loop of the other
//     warp of its scheduler and the DMMA pipe keeps a warp to issue from during epilogues; the 6-deep ring bounds the
//     drift.  (SLB200_GEMM_LAG=cycles starts the second warp of each scheduler late to force the de-phasing; measured:
//     no difference, the warps de-phase by themselves -- profiles/r01_gemm_ab_v7_v9.txt.  SLB200_GEMM_EPI=1 selects a
//     red.global.add epilogue: also no difference.)
// Per-element arithmetic (k order, one DMMA chain per element, C - acc) is identical to gemm.cu: results are
// bit-identical to every other variant.
#include "kernels.cuh"
#include "common.h"

namespace slb {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int BLK = BM * BK;                     // doubles per packed block (16 KB)
constexpr int PSTAGES = 6;
constexpr int CONSUMERS = 8;                     // consumer warps: 2 (m) x 4 (n), 64 x 32 each
constexpr int PK_THREADS = (CONSUMERS + 4) * 32; // + 1 producer warpgroup (registers are re-balanced with setmaxnreg)
constexpr int GROUP_M = 16;
constexpr size_t PK_SMEM = (size_t)PSTAGES * 2 * BLK * sizeof(double) + 2 * PSTAGES * sizeof(uint64_t) + 16;

// ---- packed layouts ---------------------------------------------------------------------------------------
// A block (tile_m, kt): element (m, k) of the 128 x 16 tile, m = wm*64 + (2p+e)*8 + g, k = s4*4 + tig, lives at
//   (((wm*4 + s4)*4 + p)*32 + g*4 + tig)*2 + e                  -> one LDS.128 per (p): fb[2p], fb[2p+1]
// B block (tile_n, kt): element (k, n), n = wn*32 + nf*16 + h*8 + g, k = s4*4 + tig, lives at
//   (((wn*4 + s4)*2 + nf)*32 + g*4 + tig)*2 + h                 -> one LDS.128 per (nf): a0, a1
__device__ __forceinline__ int a_slot(int m, int k)
{
    int wm = m >> 6, r = m & 63, mf = r >> 3, g = r & 7, s4 = k >> 2, tig = k & 3;
    return ((((wm * 4 + s4) * 4 + (mf >> 1)) * 32 + g * 4 + tig) << 1) + (mf & 1);
}
__device__ __forceinline__ int b_slot(int n, int k)
{
    int wn = n >> 5, r = n & 31, nf = r >> 4, h = (r >> 3) & 1, g = r & 7, s4 = k >> 2, tig = k & 3;
    return ((((wn * 4 + s4) * 2 + nf) * 32 + g * 4 + tig) << 1) + h;
}

// one CTA per block; A is M x K column-major (m contiguous), B is K x N column-major (k contiguous)
__global__ void __launch_bounds__(256) pack_a_kernel(int64_t M, int K, const double *__restrict__ A, int64_t lda, double *__restrict__ Ap, int KT)
{
    __shared__ double t[BLK];
    const int kt = blockIdx.x % KT; const int64_t tm = blockIdx.x / KT;
    const int64_t m0 = tm * BM; const int k0 = kt * BK;
    for (int i = threadIdx.x; i < BLK; i += 256) {
        int m = i & 127, k = i >> 7;
        double v = 0.0;
        if (m0 + m < M && k0 + k < K) v = A[m0 + m + (int64_t)(k0 + k) * lda];
        t[a_slot(m, k)] = v;
    }
    __syncthreads();
    double *dst = Ap + (int64_t)blockIdx.x * BLK;
    for (int i = threadIdx.x; i < BLK; i += 256) dst[i] = t[i];
}
__global__ void __launch_bounds__(256) pack_b_kernel(int64_t N, int K, const double *__restrict__ B, int64_t ldb, double *__restrict__ Bp, int KT)
{
    __shared__ double t[BLK];
    const int kt = blockIdx.x % KT; const int64_t tn = blockIdx.x / KT;
    const int64_t n0 = tn * BN; const int k0 = kt * BK;
    for (int i = threadIdx.x; i < BLK; i += 256) {
        int k = i & 15, n = i >> 4;
        double v = 0.0;
        if (n0 + n < N && k0 + k < K) v = B[k0 + k + (n0 + n) * ldb];
        t[b_slot(n, k)] = v;
    }
    __syncthreads();
    double *dst = Bp + (int64_t)blockIdx.x * BLK;
    for (int i = threadIdx.x; i < BLK; i += 256) dst[i] = t[i];
}

// ---- complex operands (PZGEMM, SRC/pzgetrf.f:288) on the same real kernel ------------------------------------------------
// C -= A B with A = Ar + i Ai, B = Br + i Bi becomes ONE real product with K doubled and the columns of C interleaved:
//     [Cr | Ci] -= [Ar | Ai] * [ Br  Bi ]
//                              [-Bi  Br ]
// A' = [Ar | Ai] is M x 2K' (K' = K rounded up to the 16-wide k-tile, zero padded); B' has 2N "virtual" columns, column
// 2n = (Br(:,n); -Bi(:,n)), column 2n+1 = (Bi(:,n); Br(:,n)); virtual column n' of C is the real (n' even) or imaginary (n' odd)
// plane of complex column n'/2, i.e. exactly the interleaved COMPLEX*16 storage with row stride 2 (CMODE = 1 in the kernel).
__global__ void __launch_bounds__(256) pack_a_z_kernel(int64_t M, int K, const double2 *__restrict__ A, int64_t lda, double *__restrict__ Ap, int KTh)
{
    __shared__ double t[BLK];
    const int KT = 2 * KTh;
    const int kt = blockIdx.x % KT; const int64_t tm = blockIdx.x / KT;
    const bool imag = kt >= KTh;
    const int64_t m0 = tm * BM; const int k0 = (imag ? kt - KTh : kt) * BK;
    for (int i = threadIdx.x; i < BLK; i += 256) {
        int m = i & 127, k = i >> 7;
        double v = 0.0;
        if (m0 + m < M && k0 + k < K) { const double2 z = A[m0 + m + (int64_t)(k0 + k) * lda]; v = imag ? z.y : z.x; }
        t[a_slot(m, k)] = v;
    }
    __syncthreads();
    double *dst = Ap + (int64_t)blockIdx.x * BLK;
    for (int i = threadIdx.x; i < BLK; i += 256) dst[i] = t[i];
}
// N2 = 2 N virtual columns
__global__ void __launch_bounds__(256) pack_b_z_kernel(int64_t N2, int K, const double2 *__restrict__ B, int64_t ldb, double *__restrict__ Bp, int KTh)
{
    __shared__ double t[BLK];
    const int KT = 2 * KTh;
    const int kt = blockIdx.x % KT; const int64_t tn = blockIdx.x / KT;
    const bool lower = kt >= KTh;                                      // the half of B' that multiplies Ai
    const int64_t n0 = tn * BN; const int k0 = (lower ? kt - KTh : kt) * BK;
    for (int i = threadIdx.x; i < BLK; i += 256) {
        int k = i & 15, n = i >> 4;
        double v = 0.0;
        const int64_t nv = n0 + n;
        if (nv < N2 && k0 + k < K) {
            const double2 z = B[k0 + k + (nv >> 1) * ldb];
            const bool im_col = nv & 1;
            v = !lower ? (im_col ? z.y : z.x) : (im_col ? z.x : -z.y);
        }
        t[b_slot(n, k)] = v;
    }
    __syncthreads();
    double *dst = Bp + (int64_t)blockIdx.x * BLK;
    for (int i = threadIdx.x; i < BLK; i += 256) dst[i] = t[i];
}

// ---- PTX helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra.uni WAIT_DONE;\nbra.uni WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void dmma_16x8x4(double (&d)[4], double a0, double a1, double b0)
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a0), "d"(a1), "d"(b0));
}
__device__ __forceinline__ double2 lds128(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int &tm, int &tn)
{
    int group_sz = GROUP_M * tiles_n;
    int grp = t / group_sz;
    int first_m = grp * GROUP_M;
    int gm = min(GROUP_M, tiles_m - first_m);
    int r = t - grp * group_sz;
    tm = first_m + r % gm;
    tn = r / gm;
}

// EPI: 0 = load / subtract / store (16-byte when VEC == 2), 1 = fire-and-forget red.global.add.f64 of -acc
// CMODE: 0 = C is a real column-major matrix (ldc in doubles); 1 = C is an interleaved COMPLEX*16 matrix seen as 2N virtual real
// columns (see pack_b_z_kernel): element (m, n') lives at 2 m + (n' & 1) + (n' >> 1) * ldc with ldc = 2 x the complex leading dimension
template <int CMODE>
__device__ __forceinline__ int64_t c_index(int64_t m, int64_t n, int64_t ldc)
{ return CMODE ? 2 * m + (n & 1) + (n >> 1) * ldc : m + n * ldc; }

template <int VEC, int EPI, int CMODE = 0>
__global__ void __launch_bounds__(PK_THREADS, 1)
dgemm_minus_packed(int64_t M, int64_t N, int KT, const double *__restrict__ Ap, const double *__restrict__ Bp,
                   double *__restrict__ C, int64_t ldc, int tiles_m, int tiles_n, int tn_off, int chunk, int lag)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned sbase = smem_u32(smem_raw);
    const unsigned bars = sbase + PSTAGES * 2 * BLK * 8;              // full[PSTAGES], empty[PSTAGES]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int ntiles = tiles_m * tiles_n;
    const int t_first = chunk > 0 ? (int)blockIdx.x * chunk : (int)blockIdx.x;
    const int t_stride = chunk > 0 ? 1 : (int)gridDim.x;
    const int my_tiles = chunk > 0 ? max(0, min(chunk, ntiles - t_first)) : (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * KT;

    if (tid == 0) {
        for (int s = 0; s < PSTAGES; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (PSTAGES + s), CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    if (warp >= CONSUMERS) {
        // ================= producer warpgroup: gives its registers to the consumers; warp 8 lane 0 streams the packed
        // blocks, all lanes of warp 8 prefetch the C tiles into L2 ====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (warp > CONSUMERS) return;
        int st = 0; unsigned ph = 0;
        for (int lt = 0; lt < my_tiles; ++lt) {
            int tm, tn; tile_coords(t_first + lt * t_stride, tiles_m, tiles_n, tm, tn);
            {   // C tile -> L2: 128 columns x 8 lines of 128 B
                const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;
                if (CMODE == 0) {
                    for (int i = lane; i < BN * 8; i += 32) {
                        int64_t n = n0 + (i >> 3), m = m0 + (i & 7) * 16;
                        if (n < N && m < M) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(C + m + n * ldc));
                    }
                } else {                                                // 64 complex columns x 16 lines of 128 B (8 complex rows)
                    for (int i = lane; i < (BN / 2) * 16; i += 32) {
                        int64_t n = n0 + 2 * (i >> 4), m = m0 + (i & 15) * 8;
                        if (n < N && m < M) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(C + c_index<1>(m, n, ldc)));
                    }
                }
            }
            if (lane == 0) {
                const double *ap = Ap + (int64_t)tm * KT * BLK;
                const double *bp = Bp + (int64_t)(tn + tn_off) * KT * BLK;
                for (int kt = 0; kt < KT; ++kt) {
                    const unsigned full = bars + 8 * st, empty = bars + 8 * (PSTAGES + st);
                    mbar_wait(empty, ph ^ 1);                          // first pass: passes immediately
                    mbar_expect_tx(full, 2 * BLK * 8);
                    bulk_g2s(sbase + st * (2 * BLK * 8), ap + (int64_t)kt * BLK, BLK * 8, full);
                    bulk_g2s(sbase + st * (2 * BLK * 8) + BLK * 8, bp + (int64_t)kt * BLK, BLK * 8, full);
                    if (++st == PSTAGES) { st = 0; ph ^= 1; }
                }
            }
            __syncwarp();
        }
        return;
    }

    // ================= consumer warps (two warpgroups, 232 registers per thread) =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    const int g = lane >> 2, tig = lane & 3;
    const int wm = warp & 1, wn = warp >> 1;
    const int wm0 = wm * 64, wn0 = wn * 32;
    // byte offsets of this warp's fragments inside a stage (A block first, then B block)
    const unsigned offA = (unsigned)(wm * 4 * 4 * 32 * 2 * 8) + lane * 16;               // + s4*2048 + p*512
    const unsigned offB = (unsigned)(BLK * 8 + wn * 4 * 2 * 32 * 2 * 8) + lane * 16;     // + s4*1024 + nf*512

    if (lag > 0 && warp >= 4) {                                        // second warp of each scheduler starts late
        long long t0 = clock64();
        while (clock64() - t0 < lag) { }
    }

    double acc[2][8][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[i][j][v] = 0.0;

    double2 fa[2][2], fb[2][4];
    int st = 0; unsigned ph = 0;
    if (total > 0) {
        mbar_wait(bars, 0);
#pragma unroll
        for (int nf = 0; nf < 2; ++nf) fa[0][nf] = lds128(sbase + offB + nf * 512);
#pragma unroll
        for (int p = 0; p < 4; ++p) fb[0][p] = lds128(sbase + offA + p * 512);
    }
    int kt = 0, lt = 0;
    for (int ci = 0; ci < total; ++ci) {
        const unsigned sb = sbase + st * (2 * BLK * 8);
        int nst = st + 1; unsigned nph = ph;
        if (nst == PSTAGES) { nst = 0; nph ^= 1; }
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            const int cur = s4 & 1, nxt = cur ^ 1;
            if (s4 < 3) {
#pragma unroll
                for (int nf = 0; nf < 2; ++nf) fa[nxt][nf] = lds128(sb + offB + (s4 + 1) * 1024 + nf * 512);
#pragma unroll
                for (int p = 0; p < 4; ++p) fb[nxt][p] = lds128(sb + offA + (s4 + 1) * 2048 + p * 512);
            } else if (ci + 1 < total) {
                mbar_wait(bars + 8 * nst, nph);
                const unsigned sn = sbase + nst * (2 * BLK * 8);
#pragma unroll
                for (int nf = 0; nf < 2; ++nf) fa[nxt][nf] = lds128(sn + offB + nf * 512);
#pragma unroll
                for (int p = 0; p < 4; ++p) fb[nxt][p] = lds128(sn + offA + p * 512);
            }
#pragma unroll
            for (int nf = 0; nf < 2; ++nf)
#pragma unroll
                for (int mf = 0; mf < 8; ++mf)
                    dmma_16x8x4(acc[nf][mf], fa[cur][nf].x, fa[cur][nf].y, (mf & 1) ? fb[cur][mf >> 1].y : fb[cur][mf >> 1].x);
        }
        // every fragment of stage st has been consumed by an issued DMMA: hand the buffer back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (PSTAGES + st));
        st = nst; ph = nph;

        if (++kt == KT) {
            kt = 0;
            int tm, tn; tile_coords(t_first + lt * t_stride, tiles_m, tiles_n, tm, tn);
            ++lt;
            const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;
            if (EPI == 1) {
#pragma unroll
                for (int nf = 0; nf < 2; ++nf)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int64_t n = n0 + wn0 + nf * 16 + g + h * 8;
#pragma unroll
                        for (int mf = 0; mf < 8; ++mf) {
                            const int64_t m = m0 + wm0 + mf * 8 + 2 * tig;
                            if (n < N && m < M) asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(C + c_index<CMODE>(m, n, ldc)), "d"(-acc[nf][mf][2 * h]) : "memory");
                            if (n < N && m + 1 < M) asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(C + c_index<CMODE>(m + 1, n, ldc)), "d"(-acc[nf][mf][2 * h + 1]) : "memory");
                        }
                    }
            } else {
#pragma unroll
                for (int nf = 0; nf < 2; ++nf) {
#pragma unroll
                    for (int mh = 0; mh < 2; ++mh) {                 // batches of 8 x 16 B
                        double2 cv[2][4];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int64_t n = n0 + wn0 + nf * 16 + g + h * 8;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int64_t m = m0 + wm0 + (mh * 4 + q) * 8 + 2 * tig;
                                if (VEC == 2 && n < N && m + 1 < M) cv[h][q] = *reinterpret_cast<const double2 *>(C + m + n * ldc);
                                else if (n < N && m < M) cv[h][q] = make_double2(C[c_index<CMODE>(m, n, ldc)], (m + 1 < M) ? C[c_index<CMODE>(m + 1, n, ldc)] : 0.0);
                                else cv[h][q] = make_double2(0.0, 0.0);
                            }
                        }
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int64_t n = n0 + wn0 + nf * 16 + g + h * 8;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int64_t m = m0 + wm0 + (mh * 4 + q) * 8 + 2 * tig;
                                double2 c = cv[h][q];
                                c.x -= acc[nf][mh * 4 + q][2 * h]; c.y -= acc[nf][mh * 4 + q][2 * h + 1];
                                if (VEC == 2 && n < N && m + 1 < M) *reinterpret_cast<double2 *>(C + m + n * ldc) = c;
                                else if (n < N && m < M) { C[c_index<CMODE>(m, n, ldc)] = c.x; if (m + 1 < M) C[c_index<CMODE>(m + 1, n, ldc)] = c.y; }
                            }
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
}

}  // namespace

// Packs B (and A unless reuse_a) into the per-stream workspaces and runs the packed kernel.
void launch_dgemm_minus_packed(int64_t M, int64_t N, int K, const double *A, int64_t lda, const double *B, int64_t ldb,
                               double *C, int64_t ldc, cudaStream_t s, int chunk, bool reuse_a)
{
    if (M <= 0 || N <= 0 || K <= 0) return;
    static bool attr_done = false;
    const int epi = (int)opt("gemm_epi", 0), lag = (int)opt("gemm_lag", 0);
    if (!attr_done) {
        SLB_CUDA(cudaFuncSetAttribute(dgemm_minus_packed<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM));
        SLB_CUDA(cudaFuncSetAttribute(dgemm_minus_packed<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM));
        SLB_CUDA(cudaFuncSetAttribute(dgemm_minus_packed<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM));
        attr_done = true;
    }
    const int KT = (K + BK - 1) / BK;
    const int tiles_m = (int)((M + BM - 1) / BM), tiles_n = (int)((N + BN - 1) / BN);
    const int64_t ntiles = (int64_t)tiles_m * tiles_n;
    if (ntiles * KT > 0x7fffffffLL) fatal("dgemm(packed): too many tile stages");
    double *Ap = (double *)workspace("gemm_Apack", (size_t)tiles_m * KT * BLK * sizeof(double));
    double *Bp = (double *)workspace("gemm_Bpack", (size_t)tiles_n * KT * BLK * sizeof(double));
    if (!reuse_a) {
        pack_a_kernel<<<(unsigned)(tiles_m * KT), 256, 0, s>>>(M, K, A, lda, Ap, KT);
        counter_add("kernel_launches", 1);
    }
    pack_b_kernel<<<(unsigned)(tiles_n * KT), 256, 0, s>>>(N, K, B, ldb, Bp, KT);
    const bool aligned = (((uintptr_t)C) & 15) == 0 && (ldc % 2 == 0);
    unsigned grid = (unsigned)(ntiles < rt().sm_count ? ntiles : rt().sm_count);
    if (chunk > 0 && ntiles > rt().sm_count) grid = (unsigned)((ntiles + chunk - 1) / chunk); else chunk = 0;
    if (epi == 1)
        dgemm_minus_packed<1, 1><<<grid, PK_THREADS, PK_SMEM, s>>>(M, N, KT, Ap, Bp, C, ldc, tiles_m, tiles_n, 0, chunk, lag);
    else if (aligned)
        dgemm_minus_packed<2, 0><<<grid, PK_THREADS, PK_SMEM, s>>>(M, N, KT, Ap, Bp, C, ldc, tiles_m, tiles_n, 0, chunk, lag);
    else
        dgemm_minus_packed<1, 0><<<grid, PK_THREADS, PK_SMEM, s>>>(M, N, KT, Ap, Bp, C, ldc, tiles_m, tiles_n, 0, chunk, lag);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 2);
    counter_add("gemm_launches", 1);
}

// PZGEMM 'N','N' (alpha = -1, beta = 1) through the packed real kernel: one launch, K doubled, C interleaved (see pack_b_z_kernel).
void launch_zgemm_minus_packed(int64_t M, int64_t N, int K, const zcomplex *A, int64_t lda, const zcomplex *B, int64_t ldb,
                               zcomplex *C, int64_t ldc, cudaStream_t s, int chunk, bool reuse_a)
{
    if (M <= 0 || N <= 0 || K <= 0) return;
    static bool attr_done = false;
    if (!attr_done) {
        SLB_CUDA(cudaFuncSetAttribute(dgemm_minus_packed<1, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM));
        attr_done = true;
    }
    const int KTh = (K + BK - 1) / BK, KT = 2 * KTh;
    const int64_t N2 = 2 * N;
    const int tiles_m = (int)((M + BM - 1) / BM), tiles_n = (int)((N2 + BN - 1) / BN);
    const int64_t ntiles = (int64_t)tiles_m * tiles_n;
    if (ntiles * KT > 0x7fffffffLL) fatal("zgemm(packed): too many tile stages");
    double *Ap = (double *)workspace("zgemm_Apack", (size_t)tiles_m * KT * BLK * sizeof(double));
    double *Bp = (double *)workspace("zgemm_Bpack", (size_t)tiles_n * KT * BLK * sizeof(double));
    if (!reuse_a) {
        pack_a_z_kernel<<<(unsigned)(tiles_m * KT), 256, 0, s>>>(M, K, A, lda, Ap, KTh);
        counter_add("kernel_launches", 1);
    }
    pack_b_z_kernel<<<(unsigned)(tiles_n * KT), 256, 0, s>>>(N2, K, B, ldb, Bp, KTh);
    unsigned grid = (unsigned)(ntiles < rt().sm_count ? ntiles : rt().sm_count);
    if (chunk > 0 && ntiles > rt().sm_count) grid = (unsigned)((ntiles + chunk - 1) / chunk); else chunk = 0;
    dgemm_minus_packed<1, 0, 1><<<grid, PK_THREADS, PK_SMEM, s>>>(M, N2, KT, Ap, Bp, reinterpret_cast<double *>(C), 2 * ldc, tiles_m, tiles_n, 0, chunk, 0);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 2);
    counter_add("gemm_launches", 1);
}

}  // namespace slb
