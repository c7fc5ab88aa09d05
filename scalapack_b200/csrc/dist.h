// dist.h -- host-side helpers of the SURVEY 8(f) rows: the local window of ANY sub-matrix (block-aligned or not), a blocking
// device view of a caller's local array, and small distributed matrices (right-hand sides, scale factors) replicated on the
// host in global order.  The O(N^2) data stays distributed on the GPUs; only O(N NRHS) vectors take this route.
#pragma once
#include "common.h"
#include "entry.h"

namespace slb {

// sub(A) = A(ia:ia+m-1, ja:ja+n-1), any ia / ja (TOOLS/infog2l.f + the NUMROC( M+IROFF ) idiom of e.g. SRC/pdlange.f:184-192).
// Local rows [loff_r, loff_r + mloc) x local columns [loff_c, loff_c + nloc) of the local array hold my part of sub(A).
struct AnyWindow {
    int64_t loff_r, loff_c, mloc, nloc;
    int ia, ja, mb, nb, rsrc0, csrc0, P, Q, myrow, mycol;     // rsrc0 / csrc0: the DESCRIPTOR's source process
    int arow, acol;                                           // process owning A(ia, ja)
    // global row (0-based, relative to sub(A)) of local window row il; same for columns
    int64_t grow(int64_t il) const { return (int64_t)indxl2g((int)(loff_r + il + 1), mb, myrow, rsrc0, P) - ia; }
    int64_t gcol(int64_t jl) const { return (int64_t)indxl2g((int)(loff_c + jl + 1), nb, mycol, csrc0, Q) - ja; }
};
inline AnyWindow any_window(int m, int n, int ia, int ja, const int *desc, int P, int Q, int myrow, int mycol)
{
    AnyWindow w;
    w.ia = ia; w.ja = ja; w.mb = desc[MB_]; w.nb = desc[NB_]; w.rsrc0 = desc[RSRC_]; w.csrc0 = desc[CSRC_];
    w.P = P; w.Q = Q; w.myrow = myrow; w.mycol = mycol;
    w.arow = indxg2p(ia, w.mb, w.rsrc0, P); w.acol = indxg2p(ja, w.nb, w.csrc0, Q);
    w.loff_r = numroc(ia - 1, w.mb, myrow, w.rsrc0, P);
    w.loff_c = numroc(ja - 1, w.nb, mycol, w.csrc0, Q);
    w.mloc = (int64_t)numroc(ia - 1 + m, w.mb, myrow, w.rsrc0, P) - w.loff_r;
    w.nloc = (int64_t)numroc(ja - 1 + n, w.nb, mycol, w.csrc0, Q) - w.loff_c;
    return w;
}

// Device view of the rows x cols window at `p` (leading dimension lld) of a caller's local array: in place when the caller's
// array is device-resident, else one blocking 2-D copy into a named workspace (and back with download()).
template <typename T>
struct StageMat {
    T *dev = nullptr; int64_t ld = 1; bool staged = false;
    T *host = nullptr; int64_t hld = 1, rows = 0, cols = 0;
    StageMat(const char *name, const T *base, int64_t lld, int64_t loff_r, int64_t loff_c, int64_t rows_, int64_t cols_, bool upload = true)
    {
        T *win = const_cast<T *>(base) + loff_r + loff_c * lld;
        rows = rows_; cols = cols_;
        if (rows > 0 && cols > 0 && is_device_ptr(base)) { dev = win; ld = lld; return; }
        if (rows <= 0 || cols <= 0) {                          // nothing of sub(A) lives here: a valid device address that is never read
            dev = (T *)workspace("stage_empty", 256); ld = lld > 0 ? lld : 1; rows = cols = 0;
            return;
        }
        staged = true; host = win; hld = lld;
        ld = (rows + 1) & ~(int64_t)1;                        // even: 16-byte alignment of every column for the update's epilogue
        dev = (T *)workspace(name, (size_t)ld * (size_t)cols * sizeof(T));
        if (upload) {
            SLB_CUDA(cudaMemcpy2DAsync(dev, (size_t)ld * sizeof(T), host, (size_t)hld * sizeof(T), (size_t)rows * sizeof(T), (size_t)cols,
                                       cudaMemcpyHostToDevice, rt().s_main));
            SLB_CUDA(cudaStreamSynchronize(rt().s_main));
            counter_add("h2d_bytes", (int64_t)(rows * cols * (int64_t)sizeof(T)));
        }
    }
    void download()
    {
        if (!staged) return;
        SLB_CUDA(cudaMemcpy2DAsync(host, (size_t)hld * sizeof(T), dev, (size_t)ld * sizeof(T), (size_t)rows * sizeof(T), (size_t)cols,
                                   cudaMemcpyDeviceToHost, rt().s_main));
        SLB_CUDA(cudaStreamSynchronize(rt().s_main));
        counter_add("d2h_bytes", (int64_t)(rows * cols * (int64_t)sizeof(T)));
    }
};

// every process contributes a vector of `len` doubles; all receive the element-wise sum ('+'), max ('M') or min ('m') taken
// in process order (identical bits everywhere).  Host control plane: these are O(N) vectors.
inline void grid_combine(Grid *g, char scope, double *v, size_t len, char op)
{
    const int np = grid_scope_size(g, scope);
    if (np <= 1 || len == 0) return;
    const size_t CHUNK = (size_t)1 << 21;                       // 16 MiB per process and exchange: the control plane bounds a message
    std::vector<double> all((len < CHUNK ? len : CHUNK) * (size_t)np);
    for (size_t c0 = 0; c0 < len; c0 += CHUNK) {
        const size_t cl = len - c0 < CHUNK ? len - c0 : CHUNK;
        grid_allgather(g, scope, v + c0, all.data(), cl * sizeof(double));
        for (size_t e = 0; e < cl; ++e) {
            double a = all[e];
            for (int p = 1; p < np; ++p) {
                const double b = all[(size_t)p * cl + e];
                if (op == '+') a += b;
                else if (op == 'M') { if (b > a || b != b) a = b; }
                else { if (b < a || b != b) a = b; }
            }
            v[c0 + e] = a;
        }
    }
}

// sub(B) (m x n, any alignment) of a distributed matrix whose local array starts at `base` (host or device memory)
// -> the whole sub(B) in global order on every process (column-major, ld = m).
inline void gather_small(Grid *g, int m, int n, const double *base, int ib, int jb, const int *desc, std::vector<double> &out)
{
    const AnyWindow w = any_window(m, n, ib, jb, desc, g->nprow, g->npcol, g->myrow, g->mycol);
    out.assign((size_t)m * (size_t)n, 0.0);
    if (w.mloc > 0 && w.nloc > 0) {
        std::vector<double> loc((size_t)w.mloc * (size_t)w.nloc);
        const int64_t lld = desc[LLD_];
        const double *win = base + w.loff_r + w.loff_c * lld;
        if (is_device_ptr(base))
            SLB_CUDA(cudaMemcpy2D(loc.data(), (size_t)w.mloc * sizeof(double), win, (size_t)lld * sizeof(double),
                                  (size_t)w.mloc * sizeof(double), (size_t)w.nloc, cudaMemcpyDeviceToHost));
        else
            for (int64_t jl = 0; jl < w.nloc; ++jl) memcpy(loc.data() + jl * w.mloc, win + jl * lld, (size_t)w.mloc * sizeof(double));
        for (int64_t jl = 0; jl < w.nloc; ++jl) {
            const int64_t jg = w.gcol(jl);
            for (int64_t il = 0; il < w.mloc; ++il) out[(size_t)(w.grow(il) + jg * m)] = loc[(size_t)(il + jl * w.mloc)];
        }
    }
    grid_combine(g, 'A', out.data(), out.size(), '+');
}
// the inverse: my local entries of sub(B) <- the replicated global copy
inline void scatter_small(Grid *g, int m, int n, double *base, int ib, int jb, const int *desc, const std::vector<double> &in)
{
    const AnyWindow w = any_window(m, n, ib, jb, desc, g->nprow, g->npcol, g->myrow, g->mycol);
    if (w.mloc <= 0 || w.nloc <= 0) return;
    std::vector<double> loc((size_t)w.mloc * (size_t)w.nloc);
    for (int64_t jl = 0; jl < w.nloc; ++jl) {
        const int64_t jg = w.gcol(jl);
        for (int64_t il = 0; il < w.mloc; ++il) loc[(size_t)(il + jl * w.mloc)] = in[(size_t)(w.grow(il) + jg * m)];
    }
    const int64_t lld = desc[LLD_];
    double *win = base + w.loff_r + w.loff_c * lld;
    if (is_device_ptr(base))
        SLB_CUDA(cudaMemcpy2D(win, (size_t)lld * sizeof(double), loc.data(), (size_t)w.mloc * sizeof(double),
                              (size_t)w.mloc * sizeof(double), (size_t)w.nloc, cudaMemcpyHostToDevice));
    else
        for (int64_t jl = 0; jl < w.nloc; ++jl) memcpy(win + jl * lld, loc.data() + jl * w.mloc, (size_t)w.mloc * sizeof(double));
}

}  // namespace slb
