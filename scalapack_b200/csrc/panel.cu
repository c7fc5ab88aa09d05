// panel.cu -- panel factorisation: the m x jb block column is factored with partial pivoting by ONE
// persistent cooperative kernel.  Replaces PDGETF2 (SRC/pdgetf2.f:207-237), i.e. per column
//   PDAMAX (pivot search, PBLAS/SRC/pdamax_.c:404-487)  -> warp-shuffle max-loc + one grid barrier
//   PDSWAP (row interchange inside the panel)            -> exchange through a 2-slot global mailbox
//   PDSCAL (reciprocal scale, pdgetf2.f:224)             -> fused
//   PDGER  (rank-1 update)                               -> fused, panel rows resident in shared memory
// and the jb-pivot broadcast that ends it.
//
// Structure (right-looking over sub-panels of W columns, W = 32/16/8 chosen so a CTA's row slab fits):
//   F  each CTA keeps its slab (rpb rows x W columns) of the sub-panel in shared memory.  Per column: local
//      arg-max -> publish {|v|, key, the whole candidate row} -> grid barrier -> every CTA reduces the
//      candidates redundantly, applies the swap / scale / rank-1 update to its slab.  ONE grid barrier per
//      column; the pivot row travels with the candidate so no second exchange is needed.
//   S  the W interchanges are applied to the other columns of the panel as a net permutation (one thread
//      per column), fused with the W x W unit-lower solve that produces the U rows of the columns on the
//      right.
//   G  each CTA updates its slab rows of the columns on the right (rank-W update, L from shared memory,
//      U staged through shared memory).
// Pivot rule = the reference's: max |a| (|Re|+|Im| complex), ties to the lowest process row, then to the
// lowest global row (idamax first index + the 1-tree combine's strict '<', pdamax_.c:457).
// Zero pivot: INFO records the first one, swap and scale are skipped (pdgetf2.f:214-227).
#include "kernels.cuh"
#include "devmath.cuh"
#include "common.h"

#include <cooperative_groups.h>

namespace slb {

namespace {

constexpr int PT = 256;        // threads per CTA
constexpr int UCH = 64;        // columns of U staged per chunk in phase G
constexpr int MAXG = 160;      // max CTAs (>= SM count)

struct __align__(16) CandHdr { double absval; int vrow; unsigned tag; };

// 16-byte mailbox header: written / read as one vector access so {value,row,tag} are observed together
__device__ __forceinline__ void st_hdr(CandHdr *p, double a, int v, unsigned tag)
{
    unsigned lo = (unsigned)(__double_as_longlong(a) & 0xffffffffLL), hi = (unsigned)((unsigned long long)__double_as_longlong(a) >> 32);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(lo), "r"(hi), "r"((unsigned)v), "r"(tag) : "memory");
}
__device__ __forceinline__ void ld_hdr(const CandHdr *p, double &a, int &v, unsigned &tag)
{
    unsigned lo, hi, uv;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(lo), "=r"(hi), "=r"(uv), "=r"(tag) : "l"(p) : "memory");
    a = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo)); v = (int)uv;
}

__device__ __forceinline__ void grid_barrier(unsigned *count, volatile unsigned *gen, unsigned nblocks)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned g = *gen;
        __threadfence();
        if (atomicAdd(count, 1u) == nblocks - 1) {
            *count = 0;
            __threadfence();
            atomicAdd((unsigned *)gen, 1u);
        } else {
            while (*gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}

struct VMap {
    PanelRowMap m;
    __device__ __forceinline__ void locate(int v, int &prow, int &lrow) const
    {
        int s = 0;
#pragma unroll
        for (int i = 1; i < 8; ++i) if (i < m.nseg && v >= m.seg_v0[i]) s = i;
        prow = m.seg_prow[s];
        lrow = m.seg_lr0[s] + (v - m.seg_v0[s]);
    }
    __device__ __forceinline__ int global_row(int v) const
    {
        int prow, l; locate(v, prow, l);
        return ((l / m.nb) * m.nprow + ((prow - m.rsrc + m.nprow) % m.nprow)) * m.nb + l % m.nb;
    }
    __device__ __forceinline__ long long key(int v) const
    {
        int prow, l; locate(v, prow, l);
        int g = ((l / m.nb) * m.nprow + ((prow - m.rsrc + m.nprow) % m.nprow)) * m.nb + l % m.nb;
        return ((long long)prow << 32) | (unsigned)g;
    }
};

// candidate a better than b ?
__device__ __forceinline__ bool better(double aa, long long ka, double ab, long long kb)
{
    if (aa != ab) return aa > ab;
    return ka < kb;
}

template <typename T, int W>
__global__ void __launch_bounds__(PT, 1)
panel_kernel(int m, int jb, T *__restrict__ Wp, int64_t ldw, PanelRowMap map_, int *__restrict__ ipiv_out,
             int *__restrict__ info_out, int info_offset, unsigned char *__restrict__ work, int rpb, unsigned tagbase)
{
    constexpr int LS = W + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *S = reinterpret_cast<T *>(smem_raw);            // [rpb][LS]
    T *prow_s = S + (size_t)rpb * LS;                  // [W] pivot row
    T *jrow_s = prow_s + W;                            // [W] old row jj
    T *Us = jrow_s + W;                                // [W][UCH]
    T *Ls = Us + W * UCH;                              // [W][W]  Ls[k*W + i] = L11[i][k]
    double *red_abs = reinterpret_cast<double *>(Ls + W * W);    // [8]
    int *red_v = reinterpret_cast<int *>(red_abs + 16);          // [8]   (red_abs[8] = block best)
    int *plan = red_v + 8;                             // top_src[W], out_dst[W], out_src[W]
    int *misc = plan + 3 * W;                          // [0]=winner cta [1]=winner vrow [2]=best local v

    // global work area
    unsigned *bar_count = reinterpret_cast<unsigned *>(work);
    volatile unsigned *bar_gen = reinterpret_cast<volatile unsigned *>(work + 128);
    unsigned *jtag = reinterpret_cast<unsigned *>(work + 256);                  // [2] tags of the published row jj (128 B apart)
    CandHdr *cand = reinterpret_cast<CandHdr *>(work + 512);                    // [2][MAXG]
    T *candrow = reinterpret_cast<T *>(work + 512 + 2 * MAXG * sizeof(CandHdr));   // [2][MAXG][W]
    T *rowj = candrow + 2 * MAXG * W;                                            // [2][W]
    int *piv_v = reinterpret_cast<int *>(rowj + 2 * W);                          // [jb]

    VMap vm; vm.m = map_;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    const int base = b * rpb;
    const int nrows = max(0, min(rpb, m - base));
    int parity = 0;

    for (int s0 = 0; s0 < jb; s0 += W) {
        const int w = min(W, jb - s0);
        // ---------------- load my slab of the sub-panel (rows >= s0 only) ----------------
        for (int c = 0; c < w; ++c)
            for (int i = tid; i < nrows; i += PT)
                if (base + i >= s0) S[i * LS + c] = ld_cg(Wp + (base + i) + (int64_t)(s0 + c) * ldw);
        __syncthreads();

        // ---------------- phase F: factor the sub-panel column by column ----------------
        // Per column: block arg-max -> publish {|v|, row, tag} + candidate row in the mailbox -> every CTA polls
        // the G headers until all carry this column's tag (the data IS the barrier: no counter, no second
        // round trip) -> redundant reduction -> fetch pivot row + old row jj -> swap / scale / rank-1 update,
        // which also produces each thread's candidate for the next column.
        double babs = -1.0; int bv = -1;
        auto consider = [&](double a, int v) {
            if (bv < 0 || a > babs || (a == babs && vm.key(v) < vm.key(bv))) { babs = a; bv = v; }
        };
        for (int i = tid; i < nrows; i += PT)
            if (base + i >= s0) consider(t_abs1(S[i * LS]), base + i);
        for (int j = 0; j < w; ++j) {
            const int jj = s0 + j;
            const unsigned want = tagbase + (unsigned)jj + 1u;
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
            // ---- publish ----
            const int myv = misc[2];
            const bool own_jj = jj >= base && jj < base + nrows;
            if (myv >= 0 && tid < w) candrow[((size_t)parity * MAXG + b) * W + tid] = S[(myv - base) * LS + tid];
            if (own_jj && tid < w) rowj[parity * W + tid] = S[(jj - base) * LS + tid];
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                st_hdr(&cand[parity * MAXG + b], myv >= 0 ? red_abs[8] : -1.0, myv, want);
                if (own_jj) *reinterpret_cast<volatile unsigned *>(&jtag[parity * 32]) = want;
            }
            // ---- gather: poll all headers (warp 0), reduce ----
            if (warp == 0) {
                double a = -1.0; int v = -1, wb = -1;
                for (int q = lane; q < G; q += 32) {
                    double qa; int qv; unsigned qt;
                    do { ld_hdr(&cand[parity * MAXG + q], qa, qv, qt); } while (qt != want);
                    if (qv >= 0 && (v < 0 || qa > a || (qa == a && vm.key(qv) < vm.key(v)))) { a = qa; v = qv; wb = q; }
                }
                if (lane == 0) { while (*reinterpret_cast<volatile unsigned *>(&jtag[parity * 32]) != want) { } }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    double oa = __shfl_xor_sync(0xffffffffu, a, off);
                    int ov = __shfl_xor_sync(0xffffffffu, v, off);
                    int ob = __shfl_xor_sync(0xffffffffu, wb, off);
                    if (ov >= 0 && (v < 0 || oa > a || (oa == a && vm.key(ov) < vm.key(v)))) { a = oa; v = ov; wb = ob; }
                }
                __threadfence();
                if (lane == 0) { misc[0] = wb; misc[1] = v; }
            }
            __syncthreads();
            const int wb = misc[0];
            int pv = misc[1];
            if (tid < w) {
                prow_s[tid] = ld_cg(candrow + ((size_t)parity * MAXG + wb) * W + tid);
                jrow_s[tid] = ld_cg(rowj + parity * W + tid);
            }
            __syncthreads();
            const T pivot = prow_s[j];
            const bool nonzero = !t_iszero(pivot);
            if (!nonzero) pv = jj;                       // AMAX == 0 => INDX = IX (pdamax_.c:486); no swap
            if (b == 0 && tid == 0) {
                piv_v[jj] = pv;
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
        }

        // ---------------- write the factored slab back ----------------
        for (int c = 0; c < w; ++c)
            for (int i = tid; i < nrows; i += PT)
                if (base + i >= s0) Wp[(base + i) + (int64_t)(s0 + c) * ldw] = S[i * LS + c];
        grid_barrier(bar_count, bar_gen, G);

        // ---------------- phase S: interchanges on the other panel columns (+ solve on the right) ----------------
        const int nother = jb - w;
        if (nother > 0) {
            // net permutation of the w interchanges (rows are virtual rows of the panel)
            if (tid < w) {
                int pos = s0 + tid;
                for (int s = w - 1; s >= 0; --s) {
                    int r = s0 + s, p = ld_cg(piv_v + r);
                    if (pos == r) pos = p; else if (pos == p) pos = r;
                }
                plan[tid] = pos;
                int p = ld_cg(piv_v + s0 + tid), dst = -1, src = 0;
                if (p >= s0 + w) {
                    bool first = true;
                    for (int s = 0; s < tid; ++s) if (ld_cg(piv_v + s0 + s) == p) { first = false; break; }
                    if (first) {
                        dst = p; int q = p;
                        for (int s = w - 1; s >= 0; --s) {
                            int r = s0 + s, pp = ld_cg(piv_v + r);
                            if (q == r) q = pp; else if (q == pp) q = r;
                        }
                        src = q - s0;
                    }
                }
                plan[W + tid] = dst; plan[2 * W + tid] = src;
            }
            for (int e = tid; e < w * w; e += PT) {
                int i = e % w, k = e / w;
                Ls[k * W + i] = (i > k) ? ld_cg(Wp + (s0 + i) + (int64_t)(s0 + k) * ldw) : t_zero(T());
            }
            __syncthreads();
            for (int q = b + G * tid; q < nother; q += G * PT) {
                const int col = q < s0 ? q : q + w;          // skip the sub-panel's own columns
                T *cp = Wp + (int64_t)col * ldw;
                T x[W], o[W];
#pragma unroll
                for (int t = 0; t < W; ++t) if (t < w) x[t] = ld_cg(cp + plan[t]);
#pragma unroll
                for (int t = 0; t < W; ++t) if (t < w && plan[W + t] >= 0) o[t] = ld_cg(cp + s0 + plan[2 * W + t]);
#pragma unroll
                for (int t = 0; t < W; ++t) if (t < w && plan[W + t] >= 0) cp[plan[W + t]] = o[t];
                if (col >= s0 + w) {
#pragma unroll
                    for (int k = 0; k < W - 1; ++k) {
                        if (k < w - 1) {
                            T xk = x[k];
#pragma unroll
                            for (int i = k + 1; i < W; ++i) if (i < w) x[i] = t_fnma(Ls[k * W + i], xk, x[i]);
                        }
                    }
                }
#pragma unroll
                for (int t = 0; t < W; ++t) if (t < w) cp[s0 + t] = x[t];
            }
        }
        const int nright = jb - s0 - w;
        if (nright <= 0) {
            // last sub-panel: nothing to update; the kernel ends (left-column swaps need no further sync)
            break;
        }
        grid_barrier(bar_count, bar_gen, G);

        // ---------------- phase G: my slab rows of the columns on the right -= L_slab * U ----------------
        {
            const int c_begin = s0 + w;
            for (int cc0 = 0; cc0 < nright; cc0 += UCH) {
                const int nch = min(UCH, nright - cc0);
                __syncthreads();
                for (int e = tid; e < w * nch; e += PT) {
                    int k = e % w, c = e / w;
                    Us[k * UCH + c] = ld_cg(Wp + (s0 + k) + (int64_t)(c_begin + cc0 + c) * ldw);
                }
                __syncthreads();
                for (int i = tid; i < nrows; i += PT) {
                    const int v = base + i;
                    if (v < s0 + w) continue;
                    T l[W];
#pragma unroll
                    for (int k = 0; k < W; ++k) l[k] = (k < w) ? S[i * LS + k] : t_zero(T());
                    T *rp = Wp + v + (int64_t)(c_begin + cc0) * ldw;
                    int c = 0;
                    for (; c + 4 <= nch; c += 4) {
                        T a0 = ld_cg(rp + (int64_t)(c + 0) * ldw), a1 = ld_cg(rp + (int64_t)(c + 1) * ldw);
                        T a2 = ld_cg(rp + (int64_t)(c + 2) * ldw), a3 = ld_cg(rp + (int64_t)(c + 3) * ldw);
#pragma unroll
                        for (int k = 0; k < W; ++k) {
                            a0 = t_fnma(l[k], Us[k * UCH + c + 0], a0);
                            a1 = t_fnma(l[k], Us[k * UCH + c + 1], a1);
                            a2 = t_fnma(l[k], Us[k * UCH + c + 2], a2);
                            a3 = t_fnma(l[k], Us[k * UCH + c + 3], a3);
                        }
                        rp[(int64_t)(c + 0) * ldw] = a0; rp[(int64_t)(c + 1) * ldw] = a1;
                        rp[(int64_t)(c + 2) * ldw] = a2; rp[(int64_t)(c + 3) * ldw] = a3;
                    }
                    for (; c < nch; ++c) {
                        T a0 = ld_cg(rp + (int64_t)c * ldw);
#pragma unroll
                        for (int k = 0; k < W; ++k) a0 = t_fnma(l[k], Us[k * UCH + c], a0);
                        rp[(int64_t)c * ldw] = a0;
                    }
                }
            }
        }
        __syncthreads();
        // no grid barrier needed here: the next phase F touches only this CTA's own slab rows (written above
        // by this CTA) and the mailbox; other CTAs' rows are next touched in phase S, W grid barriers later.
    }
}

template <typename T, int W>
size_t panel_smem_bytes(int rpb)
{
    return ((size_t)rpb * (W + 1) + 2 * W + (size_t)W * UCH + (size_t)W * W) * sizeof(T) + 16 * sizeof(double) +
           8 * sizeof(int) + 3 * W * sizeof(int) + 8 * sizeof(int) + 64;
}

template <typename T, int W>
bool try_launch_panel(int m, int jb, T *Wp, int64_t ldw, const PanelRowMap &map, int *ipiv_out, int *info_out,
                      int info_offset, void *work, cudaStream_t s)
{
    Runtime &r = rt();
    int nsm = r.sm_count < MAXG ? r.sm_count : MAXG;
    int rpb = (m + nsm - 1) / nsm;
    if (rpb < 128) rpb = 128;
    rpb = (rpb + 7) & ~7;
    size_t smem = panel_smem_bytes<T, W>(rpb);
    if (smem > r.smem_optin) return false;
    int G = (m + rpb - 1) / rpb;
    static bool attr_done = false;
    if (!attr_done) {
        SLB_CUDA(cudaFuncSetAttribute(panel_kernel<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r.smem_optin));
        attr_done = true;
    }
    unsigned char *wk = (unsigned char *)work;
    PanelRowMap mp = map;
    // mailbox tags = (launch epoch << 16) + column + 1: never equal to a stale tag of an earlier launch
    static unsigned epoch = 0;
    if ((++epoch & 0x7fffu) == 0) { SLB_CUDA(cudaMemsetAsync(wk + 256, 0, 256 + 2 * MAXG * sizeof(CandHdr), s)); ++epoch; }
    if (jb >= 65535) fatal("panel wider than 65534 columns");
    unsigned tagbase = (epoch & 0x7fffu) << 16;
    void *args[] = { &m, &jb, &Wp, &ldw, &mp, &ipiv_out, &info_out, &info_offset, &wk, &rpb, &tagbase };
    SLB_CUDA(cudaLaunchCooperativeKernel((void *)panel_kernel<T, W>, dim3(G), dim3(PT), args, smem, s));
    counter_add("kernel_launches", 1);
    counter_add("panel_launches", 1);
    return true;
}

template <typename T>
void launch_panel(int m, int jb, T *Wp, int64_t ldw, const PanelRowMap &map, int *ipiv_out, int *info_out,
                  int info_offset, void *work, cudaStream_t s)
{
    if (m <= 0 || jb <= 0) return;
    int forced = (int)opt("panel_width", 0);
    if ((forced == 0 || forced == 32) && sizeof(T) == 8 && try_launch_panel<T, 32>(m, jb, Wp, ldw, map, ipiv_out, info_out, info_offset, work, s)) return;
    if ((forced == 0 || forced == 16 || forced == 32) && try_launch_panel<T, 16>(m, jb, Wp, ldw, map, ipiv_out, info_out, info_offset, work, s)) return;
    if (try_launch_panel<T, 8>(m, jb, Wp, ldw, map, ipiv_out, info_out, info_offset, work, s)) return;
    fatal("panel of %d rows does not fit the shared-memory slabs", m);
}

}  // namespace

size_t panel_work_bytes(int jb)
{
    return 512 + 2 * MAXG * sizeof(CandHdr) + (size_t)(2 * MAXG * 32 + 2 * 32) * sizeof(zcomplex) + (size_t)(jb + 64) * sizeof(int) + 256;
}

void launch_dpanel(int m, int jb, double *W, int64_t ldw, const PanelRowMap &map, int *ipiv_out, int *info_out,
                   int info_offset, void *work, cudaStream_t s)
{ launch_panel<double>(m, jb, W, ldw, map, ipiv_out, info_out, info_offset, work, s); }
void launch_zpanel(int m, int jb, zcomplex *W, int64_t ldw, const PanelRowMap &map, int *ipiv_out, int *info_out,
                   int info_offset, void *work, cudaStream_t s)
{ launch_panel<zcomplex>(m, jb, W, ldw, map, ipiv_out, info_out, info_offset, work, s); }

}  // namespace slb
