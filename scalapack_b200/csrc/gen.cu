// gen.cu -- device-side test-driver helpers: the reference's PDMATGEN generator (closed form of
// TESTING/traditional/LIN/pdmatgen.f:448-510 + pmatgeninc.f), the 64-bit LCG generator used beyond
// PDMATGEN's 2^31 period, a generated-matrix mat-vec for the solve residual of pdlaschk.f:187,296, and
// the micro-benchmarks that give the roofline denominators (FP64 DMMA peak, FP64 FMA peak, copy GB/s).
#include "kernels.cuh"
#include "common.h"

namespace slb {

namespace {

struct Gen31 {
    static constexpr unsigned long long A = 1103515245ULL, C = 12345ULL, MASK = 0x7fffffffULL;
    __host__ __device__ static inline unsigned long long step(unsigned long long x) { return (A * x + C) & MASK; }
    __host__ __device__ static inline double val(unsigned long long x) { return 1.0 - 2.0 * ((double)x / 2147483648.0); }
};
struct Gen64 {
    static constexpr unsigned long long A = 6364136223846793005ULL, C = 1ULL, MASK = ~0ULL;
    __host__ __device__ static inline unsigned long long step(unsigned long long x) { return A * x + C; }
    __host__ __device__ static inline double val(unsigned long long x) { return (double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5; }
};

// (a^k, c (a^k - 1)/(a - 1)) by repeated squaring
template <typename G>
__host__ __device__ inline void lcg_jump(unsigned long long k, unsigned long long &ak, unsigned long long &ck)
{
    unsigned long long a = G::A, c = G::C, ra = 1, rc = 0;
    while (k) {
        if (k & 1) { rc = (a * rc + c) & G::MASK; ra = (a * ra) & G::MASK; }
        c = ((a + 1) * c) & G::MASK; a = (a * a) & G::MASK;
        k >>= 1;
    }
    ak = ra; ck = rc;
}

constexpr int RCH = 16;     // rows per thread
constexpr int CCH = 64;     // local columns per thread

// One thread = RCH consecutive rows of one row block x CCH local columns.
// MODE 0: write A.  MODE 1: accumulate r += A*x and rowabs += |A| (real only).
template <typename G, int MODE>
__global__ void __launch_bounds__(128)
gen_kernel(long long M, int mb, int nb, long long mloc, long long nloc, int myrow_rel, int mycol_rel, int nprow, int npcol,
           unsigned long long seed, int cplx, double *__restrict__ a, long long lda, const double *__restrict__ x,
           double *__restrict__ r, double *__restrict__ rowabs)
{
    const int chunks_per_blk = (mb + RCH - 1) / RCH;
    const long long nrblk = (mloc + mb - 1) / mb;
    const long long nrch = nrblk * chunks_per_blk;
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long rc = tid % nrch, cc = tid / nrch;
    long long jl0 = cc * CCH;
    if (jl0 >= nloc) return;
    long long lb = rc / chunks_per_blk; int ch = (int)(rc % chunks_per_blk);
    long long il0 = lb * mb + (long long)ch * RCH;
    int nr = min(RCH, mb - ch * RCH);
    if (il0 + nr > mloc) nr = (int)(mloc - il0);
    if (nr <= 0) return;
    const long long ig0 = (lb * nprow + myrow_rel) * mb + (long long)ch * RCH;
    const unsigned long long mul = cplx ? 2ULL : 1ULL;

    unsigned long long st[RCH];
    unsigned long long aM, cM;  lcg_jump<G>(mul * (unsigned long long)M, aM, cM);           // next column
    double accr[RCH], acca[RCH];
    if (MODE == 1) {
#pragma unroll
        for (int e = 0; e < RCH; ++e) { accr[e] = 0.0; acca[e] = 0.0; }
    }
    long long jl1 = min(nloc, jl0 + CCH);
    for (long long jl = jl0; jl < jl1; ++jl) {
        long long jg = ((jl / nb) * npcol + mycol_rel) * nb + jl % nb;
        if (jl == jl0 || jl % nb == 0) {
            unsigned long long ak, ck;
            lcg_jump<G>(1ULL + mul * ((unsigned long long)ig0 + (unsigned long long)jg * (unsigned long long)M), ak, ck);
            unsigned long long s = (ak * seed + ck) & G::MASK;
#pragma unroll
            for (int e = 0; e < RCH; ++e) { st[e] = s; s = G::step(s); if (cplx) s = G::step(s); }
        }
        if (MODE == 0) {
            if (!cplx) {
#pragma unroll
                for (int e = 0; e < RCH; ++e) if (e < nr) a[il0 + e + jl * lda] = G::val(st[e]);
            } else {
#pragma unroll
                for (int e = 0; e < RCH; ++e) if (e < nr) {
                    a[2 * (il0 + e + jl * lda)] = G::val(st[e]);
                    a[2 * (il0 + e + jl * lda) + 1] = G::val(G::step(st[e]));
                }
            }
        } else {
            double xj = x[jl];
#pragma unroll
            for (int e = 0; e < RCH; ++e) { double v = G::val(st[e]); accr[e] = fma(v, xj, accr[e]); acca[e] += fabs(v); }
        }
#pragma unroll
        for (int e = 0; e < RCH; ++e) st[e] = (aM * st[e] + cM) & G::MASK;
    }
    if (MODE == 1) {
#pragma unroll
        for (int e = 0; e < RCH; ++e) if (e < nr) { atomicAdd(r + il0 + e, accr[e]); atomicAdd(rowabs + il0 + e, acca[e]); }
    }
}

template <typename G, int MODE>
void launch_gen(long long M, int mb, int nb, long long mloc, long long nloc, int myrow_rel, int mycol_rel, int nprow, int npcol,
                unsigned long long seed, int cplx, double *a, long long lda, const double *x, double *r, double *rowabs,
                cudaStream_t s)
{
    if (mloc <= 0 || nloc <= 0) return;
    long long chunks_per_blk = (mb + RCH - 1) / RCH, nrblk = (mloc + mb - 1) / mb;
    long long nthreads = nrblk * chunks_per_blk * ((nloc + CCH - 1) / CCH);
    long long grid = (nthreads + 127) / 128;
    if (grid > 0x7fffffffLL) fatal("generator grid too large");
    gen_kernel<G, MODE><<<(unsigned)grid, 128, 0, s>>>(M, mb, nb, mloc, nloc, myrow_rel, mycol_rel, nprow, npcol, seed, cplx, a,
                                                       lda, x, r, rowabs);
    SLB_CUDA(cudaGetLastError());
    counter_add("kernel_launches", 1);
}

// ---------------- micro-benchmarks ----------------
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double *out)
{
    // 2 x 4 grid of independent DMMA.8x8x4 accumulators with operands in registers: the register tiling that
    // scripts/dmma_probe2.cu measured at the pipe's limit (37.0 TFLOP/s = 64 FMA/clk/SM at 1.965 GHz)
    double acc[2][4][2];
    double a[2], b[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) a[i] = out[64 + threadIdx.x + i];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = out[512 + threadIdx.x + j];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(acc[i][j][0]), "+d"(acc[i][j][1]) : "d"(a[i]), "d"(b[j]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double *out)
{
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}
__global__ void copy_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[i];
}

}  // namespace

void launch_pdmatgen_local(int m, int n, int mb, int nb, double *a, int64_t lda, int iarow, int iacol, int iseed,
                           int myrow, int mycol, int nprow, int npcol, cudaStream_t s)
{
    long long mloc = numroc(m, mb, myrow, iarow, nprow), nloc = numroc(n, nb, mycol, iacol, npcol);
    launch_gen<Gen31, 0>(m, mb, nb, mloc, nloc, (nprow + myrow - iarow) % nprow, (npcol + mycol - iacol) % npcol, nprow, npcol,
                         (unsigned long long)iseed, 0, a, lda, nullptr, nullptr, nullptr, s);
}

static long long numroc64(long long n, long long nb, int iproc, int isrc, int nprocs)
{
    long long mydist = (nprocs + iproc - isrc) % nprocs, nblocks = n / nb, r = (nblocks / nprocs) * nb, extra = nblocks % nprocs;
    if (mydist < extra) r += nb; else if (mydist == extra) r += n % nb;
    return r;
}

void launch_matgen64_local(int64_t m, int64_t n, int mb, int nb, double *a, int64_t lda, int iarow, int iacol,
                           uint64_t seed, int myrow, int mycol, int nprow, int npcol, int is_complex, cudaStream_t s)
{
    long long mloc = numroc64(m, mb, myrow, iarow, nprow), nloc = numroc64(n, nb, mycol, iacol, npcol);
    launch_gen<Gen64, 0>(m, mb, nb, mloc, nloc, (nprow + myrow - iarow) % nprow, (npcol + mycol - iacol) % npcol, nprow, npcol,
                         seed, is_complex, a, lda, nullptr, nullptr, nullptr, s);
}

void launch_gen_matvec(int64_t n, int nb, uint64_t aseed, int gen, int myrow, int mycol, int nprow, int npcol,
                       const double *xrow, double *r, double *rowabs, cudaStream_t s)
{
    long long mloc = numroc64(n, nb, myrow, 0, nprow), nloc = numroc64(n, nb, mycol, 0, npcol);
    if (gen == 31)
        launch_gen<Gen31, 1>(n, nb, nb, mloc, nloc, myrow, mycol, nprow, npcol, aseed, 0, nullptr, 0, xrow, r, rowabs, s);
    else
        launch_gen<Gen64, 1>(n, nb, nb, mloc, nloc, myrow, mycol, nprow, npcol, aseed, 0, nullptr, 0, xrow, r, rowabs, s);
}

static double time_kernel_ms(void (*fn)(void *), void *arg, int reps)
{
    cudaEvent_t e0, e1; SLB_CUDA(cudaEventCreate(&e0)); SLB_CUDA(cudaEventCreate(&e1));
    fn(arg);                                      // warm-up
    SLB_CUDA(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        SLB_CUDA(cudaEventRecord(e0, 0)); fn(arg); SLB_CUDA(cudaEventRecord(e1, 0));
        SLB_CUDA(cudaEventSynchronize(e1));
        float ms; SLB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

struct PeakArg { int iters; double *out; int blocks; };
static void run_dmma(void *p) { PeakArg *a = (PeakArg *)p; dmma_peak_kernel<<<a->blocks, 256>>>(a->iters, a->out); }
static void run_dfma(void *p) { PeakArg *a = (PeakArg *)p; dfma_peak_kernel<<<a->blocks, 256>>>(a->iters, a->out); }

double bench_dmma_peak_tflops(int iters)
{
    Runtime &r = rt();
    double *out = (double *)workspace("bench_out", 16384, true);
    PeakArg a{ iters, out, r.sm_count * 2 };
    double ms = time_kernel_ms(run_dmma, &a, 5);
    // per warp per iteration: 8 DMMA.8x8x4 x (8*8*4) FMAs x 2 flops; 8 warps per CTA
    double flops = (double)a.blocks * 8.0 * iters * 8.0 * 256.0 * 2.0;
    return flops / (ms * 1e-3) / 1e12;
}
double bench_dfma_peak_tflops(int iters)
{
    Runtime &r = rt();
    double *out = (double *)workspace("bench_out", 16384, true);
    PeakArg a{ iters, out, r.sm_count * 8 };
    double ms = time_kernel_ms(run_dfma, &a, 5);
    double flops = (double)a.blocks * 256.0 * iters * 16.0 * 2.0;
    return flops / (ms * 1e-3) / 1e12;
}
struct CopyArg { const double2 *s; double2 *d; size_t n; int blocks; };
static void run_copy(void *p) { CopyArg *a = (CopyArg *)p; copy_kernel<<<a->blocks, 512>>>(a->s, a->d, a->n); }
double bench_copy_gbs(size_t bytes)
{
    Runtime &r = rt();
    char *buf = (char *)workspace("bench_copy", 2 * bytes);
    CopyArg a{ (const double2 *)buf, (double2 *)(buf + bytes), bytes / 16, r.sm_count * 8 };
    double ms = time_kernel_ms(run_copy, &a, 5);
    return 2.0 * bytes / (ms * 1e-3) / 1e9;
}

}  // namespace slb
