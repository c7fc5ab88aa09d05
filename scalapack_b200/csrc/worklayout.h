// worklayout.h -- an operand in the WORKING LAYOUT of the SURVEY 8(f) routines: square nb x nb blocks on the caller's grid, first
// block on process (rs, 0), in a named device workspace, filled from / written back to a sub-matrix with ANY alignment, blocking
// or transposition by the redistribution engine of redist.cu.
#pragma once
#include "common.h"
#include "dist.h"
#include "launch.h"

namespace slb {

// sub(B) <- sub(A) (or its transpose; rowmap: row k of sub(B) <- row rowmap[k] of sub(A)), redist.cu
template <typename T>
void gemr2d_core(int m, int n, const T *a, int ia, int ja, const int *desca, T *b, int ib, int jb, const int *descb, int gctxt, bool tr,
                 const int *rowmap);

namespace {

// M <- f M on a rows x cols block (f == 0 stores zeros, so NaNs in an overwritten operand do not propagate, like the BLAS)
__global__ void __launch_bounds__(256)
scale_block_kernel(int64_t rows, int64_t cols, double *__restrict__ M, int64_t ld, double f)
{
    const int64_t total = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % rows, j = e / rows;
        M[i + j * ld] = f == 0.0 ? 0.0 : f * M[i + j * ld];
    }
}

// An m x n matrix in the working layout on grid g: nb x nb blocks from process (0, 0), local array in a named device workspace
struct Work {
    double *dev = nullptr; int64_t ld = 2, mloc = 0, nloc = 0; int desc[9]; int m = 0, n = 0;
    // rs: the process row that holds the first row block (0 for the PBLAS entry points; the factors' own for PDGETRS)
    Work(const char *name, Grid *g, int m_, int n_, int nb, int rs = 0)
    {
        m = m_; n = n_;
        mloc = numroc(m, nb, g->myrow, rs, g->nprow); nloc = numroc(n, nb, g->mycol, 0, g->npcol);
        ld = ((mloc > 0 ? mloc : 1) + 1) & ~(int64_t)1;
        dev = (double *)workspace(name, (size_t)ld * (size_t)(nloc > 0 ? nloc : 1) * sizeof(double));
        const int d[9] = { 1, g->ctxt, m, n, nb, nb, rs, 0, (int)ld };
        memcpy(desc, d, sizeof(d));
    }
    // this <- op(sub(S)) with sub(S) = S(is:, js:) of shape (tr ? n x m : m x n); rowmap: my row k <- row rowmap[k] of sub(S)
    void load(const double *S, int is, int js, const int *descs, bool tr, const int *rowmap = nullptr)
    { gemr2d_core<double>(tr ? n : m, tr ? m : n, S, is, js, descs, dev, 1, 1, desc, desc[CTXT_], tr, rowmap); }
    // sub(D) (shape tr ? n x m : m x n) <- op(this); rowmap: row k of sub(D) <- my row rowmap[k]
    void store(double *D, int id, int jd, const int *descd, bool tr, const int *rowmap = nullptr) const
    { gemr2d_core<double>(m, n, dev, 1, 1, desc, D, id, jd, descd, desc[CTXT_], tr, rowmap); }
    void scale(double f) const
    {
        if (mloc > 0 && nloc > 0 && f != 1.0) SLB_LAUNCH(scale_block_kernel, grid1d(mloc * nloc), 256, rt().s_main, mloc, nloc, dev, ld, f);
    }
};

}  // namespace

}  // namespace slb
