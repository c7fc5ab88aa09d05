// panel.cu -- panel factorisation (replaces PDGETF2, SRC/pdgetf2.f:207-237, and its per-column PBLAS calls
//   PDAMAX (PBLAS/SRC/pdamax_.c:404-487), PDSWAP, PDSCAL (pdgetf2.f:224), PDGER, IGEBS2D of the pivots).
//
// The m x jb panel (rows in GLOBAL order; on a P>1 grid the process column's panel has been gathered on one
// GPU, lu.cu) is factored recursively (Toledo/Gustavson):
//     rlu(c0,c1): if c1-c0 <= W: leaf kernel   else: rlu(left); swap+TRSM+GEMM on the right; rlu(right); swap left
// so all but O(m*jb*W) flops run in the DMMA GEMM of gemm.cu and the panel is streamed log2(jb/W) times instead
// of jb/W times.  Pivot choice is the reference's (unblocked rule) in exact arithmetic: first max |a| (|Re|+|Im|
// complex), ties to the lowest process row then the lowest row (idamax + the 1-tree combine's strict '<',
// pdamax_.c:457); a zero pivot records INFO and skips swap + scale (pdgetf2.f:214-227).
//
// Leaf kernel = ONE persistent launch for W (<= 32) columns, rows x W slab of every CTA resident in SHARED
// MEMORY.  Per column there is no grid barrier: every CTA publishes {|v|, row, tag} + its candidate row in a
// global mailbox and polls the other CTAs' 16-byte headers (one header per thread, one L2 round trip) until all
// carry this column's tag -- the data is the barrier; the pivot row travels with the candidate, so the swap needs
// no second exchange.  Then each CTA applies swap / reciprocal scale / rank-1 update to its slab and derives its
// candidate for the next column in the same sweep (warp-shuffle max-loc).
#include "kernels.cuh"
#include "devmath.cuh"
#include "common.h"

namespace slb {

namespace {

constexpr int PT = 256;        // threads per CTA
constexpr int MAXG = 160;      // max CTAs (>= SM count)

// 16-byte self-validating mailbox packet {value, tag, aux}: written / read as ONE vector access, so a reader that sees
// the tag it waits for also sees the value -- no fence and no second round trip ("the data is the flag").
struct __align__(16) Pkt { double v; unsigned tag; unsigned aux; };
__device__ __forceinline__ void st_pkt(Pkt *p, double a, unsigned tag, unsigned aux)
{
    unsigned lo = (unsigned)(__double_as_longlong(a) & 0xffffffffLL), hi = (unsigned)((unsigned long long)__double_as_longlong(a) >> 32);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(lo), "r"(hi), "r"(tag), "r"(aux) : "memory");
}
__device__ __forceinline__ void ld_pkt(const Pkt *p, double &a, unsigned &tag, unsigned &aux)
{
    unsigned lo, hi;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(lo), "=r"(hi), "=r"(tag), "=r"(aux) : "l"(p) : "memory");
    a = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

constexpr int G1 = 32;         // up to G1 CTAs: one-phase exchange (every CTA reads every candidate row)

struct VMap {
    PanelRowMap m;
    __device__ __forceinline__ int global_row(int v) const { return m.g0 + v; }
    // tie-break order of the reference: lower process row first, then lower global row
    __device__ __forceinline__ long long key(int v) const
    {
        int g = m.g0 + v;
        int prow = (m.rsrc + g / m.nb) % m.nprow;
        return ((long long)prow << 32) | (unsigned)g;
    }
};

// Mailbox layout in `work` (global): cand[2][MAXG][WN + 1] packets (packet 0 = {|v|, tag, vrow}, 1..WN = the candidate
// row as doubles), rowj[2][WN] packets (the current row jj, published by its owner); WN = W doubles (2 W for complex).
template <typename T, int W>
__global__ void __launch_bounds__(PT, 1)
panel_leaf_kernel(int m, int w, T *__restrict__ Wp, int64_t ldw, PanelRowMap map_, int *__restrict__ ipiv_out,
                  int *__restrict__ info_out, int info_offset, unsigned char *__restrict__ work, int rpb, unsigned tagbase,
                  unsigned long long *__restrict__ dbg)
{
    constexpr int LS = W + 1;
    constexpr int NE = sizeof(T) / sizeof(double);
    constexpr int WN = W * NE, WN1 = WN + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *crow = reinterpret_cast<double *>(smem_raw);          // [G1][WN] candidate rows (one-phase) / [WN] winner row
    double *jrow_d = crow + G1 * WN;                              // [WN] old row jj
    double *cs_a = jrow_d + WN;                                   // [G1]
    double *red_abs = cs_a + G1;                                  // [16]
    int *cs_v = reinterpret_cast<int *>(red_abs + 16);            // [G1]
    int *red_v = cs_v + G1;                                       // [16]
    int *red_b = red_v + 16;                                      // [16]
    int *misc = red_b + 16;                                       // [8]: [0]=winner cta [1]=winner vrow [2]=best local v
    T *S = reinterpret_cast<T *>(misc + 8);                       // [rpb][LS]
    const double *Sd = reinterpret_cast<const double *>(S);

    Pkt *candp = reinterpret_cast<Pkt *>(work + 512);             // [2][MAXG][WN1]
    Pkt *rowjp = candp + (size_t)2 * MAXG * WN1;                  // [2][WN]

    VMap vm; vm.m = map_;
    auto stamp = [&](int jj, int k) {
        if (dbg != nullptr && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
            unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            dbg[((size_t)(blockIdx.x == 0 ? 0 : 1) * 4096 + jj) * 8 + k] = t;
        }
    };
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    const int base = b * rpb;
    const int nrows = max(0, min(rpb, m - base));
    int parity = 0;

    // ---------------- load my slab ----------------
    for (int c = 0; c < w; ++c)
        for (int i = tid; i < nrows; i += PT) S[i * LS + c] = ld_cg(Wp + (base + i) + (int64_t)c * ldw);
    for (int i = tid; i < nrows; i += PT)
        for (int c = w; c < W; ++c) S[i * LS + c] = t_zero(T());     // unused columns of a narrow leaf: defined values
    __syncthreads();

    double babs = -1.0; int bv = -1;
    auto consider = [&](double a, int v) {
        if (bv < 0 || a > babs || (a == babs && vm.key(v) < vm.key(bv))) { babs = a; bv = v; }
    };
    for (int i = tid; i < nrows; i += PT) consider(t_abs1(S[i * LS]), base + i);

    for (int j = 0; j < w; ++j) {
        const int jj = j;
        const unsigned want = tagbase + (unsigned)jj + 1u;
        stamp(jj, 0);
        // ---- block reduction of the per-thread candidates ----
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double oa = __shfl_xor_sync(0xffffffffu, babs, off);
            int ov = __shfl_xor_sync(0xffffffffu, bv, off);
            if (ov >= 0 && (bv < 0 || oa > babs || (oa == babs && vm.key(ov) < vm.key(bv)))) { babs = oa; bv = ov; }
        }
        if (lane == 0) { red_abs[warp] = babs; red_v[warp] = bv; }
        __syncthreads();
        if (warp == 0) {
            double a = lane < PT / 32 ? red_abs[lane] : -1.0;
            int v = lane < PT / 32 ? red_v[lane] : -1;
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) {
                double oa = __shfl_xor_sync(0xffffffffu, a, off);
                int ov = __shfl_xor_sync(0xffffffffu, v, off);
                if (ov >= 0 && (v < 0 || oa > a || (oa == a && vm.key(ov) < vm.key(v)))) { a = oa; v = ov; }
            }
            if (lane == 0) { misc[2] = v; red_abs[8] = a; }
        }
        __syncthreads();
        stamp(jj, 1);
        // ---- publish: candidate row + header, and row jj from its owner -- plain packet stores, no fence ----
        const int myv = misc[2];
        const bool own_jj = jj >= base && jj < base + nrows;
        {
            Pkt *mine = candp + ((size_t)parity * MAXG + b) * WN1;
            if (tid < WN) st_pkt(mine + 1 + tid, myv >= 0 ? Sd[(size_t)(myv - base) * LS * NE + tid] : 0.0, want, 0u);
            else if (tid == WN) st_pkt(mine, myv >= 0 ? red_abs[8] : -1.0, want, (unsigned)myv);
            if (own_jj && tid >= 128 && tid < 128 + WN) st_pkt(rowjp + parity * WN + (tid - 128), Sd[(size_t)(jj - base) * LS * NE + (tid - 128)], want, 0u);
        }
        stamp(jj, 2);
        int wb, pv;
        const double *prow_d;
        if (G <= G1) {
            // ---- one-phase gather: every packet of every candidate (+ row jj) in ONE L2 round trip ----
            const int ncand = G * WN1, total = ncand + WN;
            for (int p = tid; p < total; p += PT) {
                const Pkt *src = p < ncand ? candp + ((size_t)parity * MAXG + p / WN1) * WN1 + p % WN1 : rowjp + parity * WN + (p - ncand);
                double v; unsigned t, aux;
                do { ld_pkt(src, v, t, aux); } while (t != want);
                if (p < ncand) {
                    const int q = p / WN1, e = p % WN1;
                    if (e == 0) { cs_a[q] = v; cs_v[q] = (int)aux; } else crow[q * WN + e - 1] = v;
                } else jrow_d[p - ncand] = v;
            }
            __syncthreads();
            if (warp == 0) {
                double a = -1.0; int v = -1, qb = -1;
                if (lane < G) { a = cs_a[lane]; v = cs_v[lane]; qb = lane; if (v < 0) { a = -1.0; qb = -1; } }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    double oa = __shfl_xor_sync(0xffffffffu, a, off);
                    int ov = __shfl_xor_sync(0xffffffffu, v, off);
                    int ob = __shfl_xor_sync(0xffffffffu, qb, off);
                    if (ov >= 0 && (v < 0 || oa > a || (oa == a && vm.key(ov) < vm.key(v)))) { a = oa; v = ov; qb = ob; }
                }
                if (lane == 0) { misc[0] = qb; misc[1] = v; }
            }
            __syncthreads();
            wb = misc[0]; pv = misc[1];
            prow_d = crow + (wb < 0 ? 0 : wb) * WN;
        } else {
            // ---- two-phase gather: thread q polls header q, block reduce, then the winner's row packets ----
            double a = -1.0; int v = -1, qb = -1;
            if (tid < G) {
                unsigned qt, aux;
                do { ld_pkt(candp + ((size_t)parity * MAXG + tid) * WN1, a, qt, aux); } while (qt != want);
                v = (int)aux; qb = tid;
                if (v < 0) { a = -1.0; qb = -1; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                double oa = __shfl_xor_sync(0xffffffffu, a, off);
                int ov = __shfl_xor_sync(0xffffffffu, v, off);
                int ob = __shfl_xor_sync(0xffffffffu, qb, off);
                if (ov >= 0 && (v < 0 || oa > a || (oa == a && vm.key(ov) < vm.key(v)))) { a = oa; v = ov; qb = ob; }
            }
            if (lane == 0) { red_abs[warp] = a; red_v[warp] = v; red_b[warp] = qb; }
            __syncthreads();
            if (warp == 0) {
                a = lane < PT / 32 ? red_abs[lane] : -1.0;
                v = lane < PT / 32 ? red_v[lane] : -1;
                qb = lane < PT / 32 ? red_b[lane] : -1;
#pragma unroll
                for (int off = 4; off > 0; off >>= 1) {
                    double oa = __shfl_xor_sync(0xffffffffu, a, off);
                    int ov = __shfl_xor_sync(0xffffffffu, v, off);
                    int ob = __shfl_xor_sync(0xffffffffu, qb, off);
                    if (ov >= 0 && (v < 0 || oa > a || (oa == a && vm.key(ov) < vm.key(v)))) { a = oa; v = ov; qb = ob; }
                }
                if (lane == 0) { misc[0] = qb; misc[1] = v; }
            }
            __syncthreads();
            wb = misc[0]; pv = misc[1];
            if (tid < WN) {
                double x; unsigned t, aux;
                do { ld_pkt(candp + ((size_t)parity * MAXG + (wb < 0 ? 0 : wb)) * WN1 + 1 + tid, x, t, aux); } while (t != want);
                crow[tid] = x;
            } else if (tid >= 128 && tid < 128 + WN) {
                double x; unsigned t, aux;
                do { ld_pkt(rowjp + parity * WN + (tid - 128), x, t, aux); } while (t != want);
                jrow_d[tid - 128] = x;
            }
            __syncthreads();
            prow_d = crow;
        }
        stamp(jj, 4);
        const T *prow_s = reinterpret_cast<const T *>(prow_d);
        const T *jrow_s = reinterpret_cast<const T *>(jrow_d);
        const T pivot = prow_s[j];
        const bool nonzero = wb >= 0 && !t_iszero(pivot);
        if (!nonzero) pv = jj;                       // AMAX == 0 => INDX = IX (pdamax_.c:486); no swap
        if (b == 0 && tid == 0) {
            ipiv_out[jj] = vm.global_row(pv) + 1;
            if (!nonzero && *info_out == 0) *info_out = info_offset + jj + 1;
        }
        if (nonzero && pv != jj) {
            if (pv >= base && pv < base + nrows && tid < w) S[(pv - base) * LS + tid] = jrow_s[tid];
            if (own_jj && tid < w) S[(jj - base) * LS + tid] = prow_s[tid];
            __syncthreads();
        }
        babs = -1.0; bv = -1;
        if (nonzero) {
            const T rinv = t_recip(pivot);
            for (int i = tid; i < nrows; i += PT) {
                if (base + i <= jj) continue;
                T *row = S + i * LS;
                T l = t_mul(row[j], rinv);
                row[j] = l;
                for (int c = j + 1; c < w; ++c) row[c] = t_fnma(l, prow_s[c], row[c]);
                if (j + 1 < w) consider(t_abs1(row[j + 1]), base + i);
            }
        } else if (j + 1 < w) {
            for (int i = tid; i < nrows; i += PT)
                if (base + i > jj) consider(t_abs1(S[i * LS + j + 1]), base + i);
        }
        parity ^= 1;
        __syncthreads();
        stamp(jj, 5);
    }

    // ---------------- write the factored slab back ----------------
    for (int c = 0; c < w; ++c)
        for (int i = tid; i < nrows; i += PT) Wp[(base + i) + (int64_t)c * ldw] = S[i * LS + c];
}

inline unsigned &leaf_epoch() { static unsigned e = 0; return e; }

template <typename T, int W>
size_t leaf_smem_bytes(int rpb)
{
    constexpr int WN = W * (int)(sizeof(T) / sizeof(double));
    return (size_t)(G1 * WN + WN + G1 + 16) * sizeof(double) + (size_t)(G1 + 16 + 16 + 8) * sizeof(int) + (size_t)rpb * (W + 1) * sizeof(T) + 32;
}

// rows per CTA for an m-row leaf of width W; 0 if it does not fit
template <typename T, int W>
int leaf_rpb(int m, int gmax)
{
    Runtime &r = rt();
    int nsm = r.sm_count < MAXG ? r.sm_count : MAXG;
    if (gmax > 0 && gmax < nsm) nsm = gmax;
    int rpb = (m + nsm - 1) / nsm;
    if (rpb < 128) rpb = 128;
    rpb = (rpb + 7) & ~7;
    return leaf_smem_bytes<T, W>(rpb) <= r.smem_optin ? rpb : 0;
}

template <typename T, int W>
void launch_leaf(int m, int w, T *Wp, int64_t ldw, const PanelRowMap &map, int *ipiv_out, int *info_out, int info_offset,
                 void *work, cudaStream_t s, int gmax, bool coop)
{
    Runtime &r = rt();
    int rpb = leaf_rpb<T, W>(m, gmax);
    if (rpb == 0) fatal("panel leaf of %d rows does not fit in shared memory", m);
    int G = (m + rpb - 1) / rpb;
    static bool attr_done = false;
    if (!attr_done) {
        SLB_CUDA(cudaFuncSetAttribute(panel_leaf_kernel<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r.smem_optin));
        attr_done = true;
    }
    unsigned char *wk = (unsigned char *)work;
    // mailbox tags = (launch epoch << 8) + column + 1: never equal to a stale tag of an earlier launch.  ONE counter for
    // every instantiation (real / complex, every leaf width): they all share the mailbox.
    unsigned &epoch = leaf_epoch();
    if ((++epoch & 0x7fffffu) == 0) { SLB_CUDA(cudaMemsetAsync(wk, 0, panel_work_bytes(0), s)); ++epoch; }
    unsigned tagbase = (epoch & 0x7fffffu) << 8;
    unsigned long long *dbg = opt("panel_debug", 0) ? (unsigned long long *)workspace("panel_dbg", 2 * 4096 * 8 * 8, true) : nullptr;
    size_t smem = leaf_smem_bytes<T, W>(rpb);
    // every CTA must be resident at the same time (they wait on each other's mailbox entries): G <= #SMs at one
    // CTA per SM; launched as a cooperative grid so the runtime checks exactly that.
    PanelRowMap mp = map;
    void *args[] = { &m, &w, &Wp, &ldw, &mp, &ipiv_out, &info_out, &info_offset, &wk, &rpb, &tagbase, &dbg };
    if (!coop) {        // look-ahead: plain launch; the CTAs become resident as the update's chunked CTAs retire
        panel_leaf_kernel<T, W><<<G, PT, smem, s>>>(m, w, Wp, ldw, mp, ipiv_out, info_out, info_offset, wk, rpb, tagbase, dbg);
        SLB_CUDA(cudaGetLastError());
    } else
        SLB_CUDA(cudaLaunchCooperativeKernel((void *)panel_leaf_kernel<T, W>, dim3(G), dim3(PT), args, smem, s));
    counter_add("kernel_launches", 1);
    counter_add("panel_launches", 1);
}

template <typename T> struct LeafCfg;
template <> struct LeafCfg<double> { static constexpr int W0 = 32; };
template <> struct LeafCfg<zcomplex> { static constexpr int W0 = 16; };

template <typename T>
int pick_leaf_width(int m, int gmax)
{
    int forced = (int)opt("panel_width", 0);
    if (forced == 8 || forced == 16 || (forced == 32 && LeafCfg<T>::W0 == 32)) return forced;
    if constexpr (LeafCfg<T>::W0 == 32) { if (leaf_rpb<T, 32>(m, gmax)) return 32; }
    if (leaf_rpb<T, 16>(m, gmax)) return 16;
    if (leaf_rpb<T, 8>(m, gmax)) return 8;
    fatal("panel of %d rows does not fit the shared-memory slabs", m);
}

template <typename T>
void leaf_dispatch(int W, int m, int w, T *Wp, int64_t ldw, const PanelRowMap &map, int *ipiv, int *info, int off, void *work, cudaStream_t s, int gmax, bool coop)
{
    if (W == 32) { if constexpr (LeafCfg<T>::W0 == 32) launch_leaf<T, 32>(m, w, Wp, ldw, map, ipiv, info, off, work, s, gmax, coop); }
    else if (W == 16) launch_leaf<T, 16>(m, w, Wp, ldw, map, ipiv, info, off, work, s, gmax, coop);
    else launch_leaf<T, 8>(m, w, Wp, ldw, map, ipiv, info, off, work, s, gmax, coop);
}

inline void gemm_minus(int64_t M, int64_t N, int K, const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc, cudaStream_t s)
{ launch_dgemm_minus(M, N, K, A, lda, B, ldb, C, ldc, s); }
inline void gemm_minus(int64_t M, int64_t N, int K, const zcomplex *A, int64_t lda, const zcomplex *B, int64_t ldb, zcomplex *C, int64_t ldc, cudaStream_t s)
{ launch_zgemm_minus(M, N, K, A, lda, B, ldb, C, ldc, s); }
inline void trsm_llnu(int jb, int64_t n, const double *L, int64_t ldl, double *B, int64_t ldb, cudaStream_t s) { launch_dtrsm_llnu(jb, n, L, ldl, B, ldb, s); }
inline void trsm_llnu(int jb, int64_t n, const zcomplex *L, int64_t ldl, zcomplex *B, int64_t ldb, cudaStream_t s) { launch_ztrsm_llnu(jb, n, L, ldl, B, ldb, s); }

template <typename T>
struct PanelCtx {
    int m, W; T *Wp; int64_t ldw; PanelRowMap map; int *ipiv; int *info; int info_offset; void *work;
    SwapPlan plan; T *U; T *O; cudaStream_t s; int gmax; bool coop;
};

// interchanges of panel columns [p0,p1) (pivots ipiv[p0..p1)) applied to panel columns [c0,c1); the permuted top
// block lands in Wp[p0:p1, c0:c1)
template <typename T>
void panel_swap(PanelCtx<T> &c, int p0, int p1, int c0, int c1)
{
    const int jb = p1 - p0, nc = c1 - c0;
    if (jb <= 0 || nc <= 0) return;
    RowDist rd{ 1 << 30, 1, 0, 0, c.map.g0 };            // rows of Wp are global rows g0, g0+1, ...
    const int j0 = c.map.g0 + p0;
    launch_swap_plan(j0, jb, c.ipiv + p0, c.plan, c.s);
    launch_swap_pack<T>(jb, j0, c.plan, rd, c.Wp, c.ldw, c0, c1, c.U, jb, c.O, jb, c.s);
    launch_swap_unpack_out<T>(jb, c.plan, rd, c.Wp, c.ldw, c0, c1, c.O, jb, c.s);
    launch_copy2d<T>(jb, nc, c.U, jb, c.Wp + p0 + (int64_t)c0 * c.ldw, c.ldw, c.s);
}

template <typename T>
void panel_rec(PanelCtx<T> &c, int c0, int c1)
{
    const int n = c1 - c0;
    if (n <= c.W) {
        PanelRowMap mp = c.map; mp.g0 += c0;
        leaf_dispatch<T>(c.W, c.m - c0, n, c.Wp + c0 + (int64_t)c0 * c.ldw, c.ldw, mp, c.ipiv + c0, c.info, c.info_offset + c0, c.work, c.s, c.gmax, c.coop);
        return;
    }
    const int half = ((n / 2 + c.W - 1) / c.W) * c.W;
    const int mid = c0 + half;
    panel_rec(c, c0, mid);
    panel_swap(c, c0, mid, mid, c1);                                                            // PDLASWP on the right half
    trsm_llnu(mid - c0, c1 - mid, c.Wp + c0 + (int64_t)c0 * c.ldw, c.ldw, c.Wp + c0 + (int64_t)mid * c.ldw, c.ldw, c.s);
    if (c.m > mid)
        gemm_minus(c.m - mid, c1 - mid, mid - c0, c.Wp + mid + (int64_t)c0 * c.ldw, c.ldw, c.Wp + c0 + (int64_t)mid * c.ldw, c.ldw,
                   c.Wp + mid + (int64_t)mid * c.ldw, c.ldw, c.s);
    panel_rec(c, mid, c1);
    panel_swap(c, mid, c1, c0, mid);                                                            // and on the left half
}

template <typename T>
void launch_panel(int m, int jb, T *Wp, int64_t ldw, const PanelRowMap &map, int *ipiv_out, int *info_out,
                  int info_offset, void *work, cudaStream_t s, int gmax)
{
    if (m <= 0 || jb <= 0) return;
    PanelCtx<T> c;
    // The CTA count (and with it the leaf width W and the recursion tree, i.e. the ARITHMETIC of the panel) is a function of m
    // and the panel_gmax option only -- never of whether this panel happens to overlap a trailing update: a host-resident
    // caller's schedule follows the arrival of its column slabs (lu.cu), and the factors must not depend on that.
    const bool lookahead = gmax > 0;
    gmax = (int)opt("panel_gmax", 32);
    if (gmax <= 0) gmax = G1;
    // if the slabs of so few CTAs cannot hold m rows, allow more (0 = whole GPU)
    while (gmax > 0 && leaf_rpb<T, 8>(m, gmax) == 0) gmax = (2 * gmax >= rt().sm_count) ? 0 : 2 * gmax;
    // A capped look-ahead panel is launched plainly and its CTAs become resident as the chunked update's CTAs retire; a panel
    // that needs (nearly) every SM cannot rely on that: it is launched cooperatively, so the runtime guarantees co-residency
    // (it then starts once the kernels ahead of it have drained).
    c.coop = !lookahead || gmax == 0 || gmax > rt().sm_count / 2;
    // a panel that fits the slabs of G1 CTAs at the widest leaf uses no more: the one-phase exchange (one L2 round
    // trip per column) needs G <= G1, and per-column latency, not bandwidth, bounds a panel of this size
    if ((gmax == 0 || gmax > G1) && leaf_rpb<T, LeafCfg<T>::W0>(m, G1) != 0) gmax = G1;
    c.gmax = gmax;
    c.m = m; c.W = pick_leaf_width<T>(m, gmax); c.Wp = Wp; c.ldw = ldw; c.map = map; c.ipiv = ipiv_out; c.info = info_out;
    c.info_offset = info_offset; c.work = work; c.s = s;
    int *pm = (int *)workspace("panel_plan", (size_t)3 * jb * sizeof(int));
    c.plan = SwapPlan{ pm, pm + jb, pm + 2 * jb };
    c.U = (T *)workspace("panel_U", (size_t)jb * jb * sizeof(T));
    c.O = (T *)workspace("panel_O", (size_t)jb * jb * sizeof(T));
    panel_rec(c, 0, jb);
}

}  // namespace

size_t panel_work_bytes(int jb)
{
    (void)jb;   // cand[2][MAXG][WN + 1] + rowj[2][WN] packets, WN <= 32 doubles (W = 32 real, W = 16 complex)
    return 512 + (size_t)2 * MAXG * 33 * sizeof(Pkt) + (size_t)2 * 32 * sizeof(Pkt) + 256;
}

void launch_dpanel(int m, int jb, double *W, int64_t ldw, const PanelRowMap &map, int *ipiv_out, int *info_out,
                   int info_offset, void *work, cudaStream_t s, int gmax)
{ launch_panel<double>(m, jb, W, ldw, map, ipiv_out, info_out, info_offset, work, s, gmax); }
void launch_zpanel(int m, int jb, zcomplex *W, int64_t ldw, const PanelRowMap &map, int *ipiv_out, int *info_out,
                   int info_offset, void *work, cudaStream_t s, int gmax)
{ launch_panel<zcomplex>(m, jb, W, ldw, map, ipiv_out, info_out, info_offset, work, s, gmax); }

}  // namespace slb
