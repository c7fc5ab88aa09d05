// entry.h -- helpers shared by the reference-facing entry points (api.cu and the files of the "next" rows of SURVEY 8f):
// where an operand lives, the local window of a block-aligned sub-matrix, the row-distributed IPIV <-> the global pivot vector.
#pragma once
#include "common.h"

namespace slb {

inline bool is_device_ptr(const void *p)
{
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    if (at.type == cudaMemoryTypeDevice && at.device != rt().device)
        fatal("a device-resident operand lives on GPU %d but this BLACS process drives GPU %d (LOCAL_RANK): allocate it on the "
              "process's own GPU", at.device, rt().device);
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

inline void xerbla(int ictxt, const char *name, int info) { int p = -info; pxerbla_(&ictxt, name, &p); }

// Local window of a block-aligned sub-matrix sub(A) = A(IA:IA+M-1, JA:JA+N-1) (SRC/pdgetrf.f:178-186, TOOLS/infog2l.f): every
// process holds a contiguous mloc x nloc window of its local array starting at (loff_r, loff_c), and sub(A) is itself a
// block-cyclic matrix whose first block lives on process (rsrc, csrc).
struct Window { int64_t loff_r, loff_c, mloc, nloc; int rsrc, csrc; };
inline Window window(int m, int n, int ia, int ja, const int *desc, int P, int Q, int myrow, int mycol)
{
    Window w;
    const int mb = desc[MB_], nb = desc[NB_];
    w.loff_r = numroc(ia - 1, mb, myrow, desc[RSRC_], P);
    w.loff_c = numroc(ja - 1, nb, mycol, desc[CSRC_], Q);
    w.rsrc = indxg2p(ia, mb, desc[RSRC_], P);
    w.csrc = indxg2p(ja, nb, desc[CSRC_], Q);
    w.mloc = numroc(m, mb, myrow, w.rsrc, P);
    w.nloc = numroc(n, nb, mycol, w.csrc, Q);
    return w;
}

// local IPIV(loff + il) <- global pivot (as a row index of A, not of sub(A)) of each locally owned row < mn (SRC/pdgetrf.f:118-121)
inline void fill_local_ipiv(const std::vector<int> &ipg, int mn, int nb, int rsrc, int P, int myrow, int *ipiv, int row0)
{
    for (int gi = 0; gi < mn; ++gi) {
        if (indxg2p(gi + 1, nb, rsrc, P) != myrow) continue;
        ipiv[indxg2l(gi + 1, nb, P) - 1] = ipg[gi] + row0;
    }
}

// pivot vector of sub(A) (1-based, relative to sub(A)) from the row-distributed IPIV (each process row holds the entries of its rows)
inline void gather_global_ipiv(Grid *g, int n, int nb, int rsrc, const int *ipiv_local, int row0, std::vector<int> &ipg)
{
    const int P = g->nprow;
    ipg.assign((size_t)n, 0);
    std::vector<int> mine((size_t)n, 0);
    for (int gi = 0; gi < n; ++gi)
        if (indxg2p(gi + 1, nb, rsrc, P) == g->myrow) mine[gi] = ipiv_local[indxg2l(gi + 1, nb, P) - 1] - row0;
    if (P == 1) ipg = mine;
    else {
        std::vector<int> all((size_t)n * P);
        grid_allgather(g, 'C', mine.data(), all.data(), (size_t)n * sizeof(int));
        for (int p = 0; p < P; ++p) for (int gi = 0; gi < n; ++gi) if (all[(size_t)p * n + gi] > ipg[gi]) ipg[gi] = all[(size_t)p * n + gi];
    }
    // what PDGETRF writes is a pivot sequence: row i is exchanged with a row at or below it, inside sub(A).  Anything else (an IPIV
    // of another sub-matrix, an uninitialised array) would index outside the permutation below -- silently in the reference's PDLAPIV;
    // here it stops the run with a message.
    for (int gi = 0; gi < n; ++gi)
        if (ipg[gi] < gi + 1 || ipg[gi] > n)
            fatal("IPIV: the entry for row %d of sub(A) is %d (expected %d .. %d): not the pivots PDGETRF returned for this sub-matrix", gi + 1,
                  ipg[gi], gi + 1, n);
}

// the right-hand sides sub(B) = B(IB:IB+N-1, JB:JB+NRHS-1): rows aligned with sub(A) (checked), columns anywhere in B
struct RhsWindow { int64_t loff_r, nloc_all; };
inline RhsWindow rhs_window(int ib, const int *descb, int P, int Q, int myrow, int mycol)
{
    RhsWindow w;
    w.loff_r = numroc(ib - 1, descb[MB_], myrow, descb[RSRC_], P);
    w.nloc_all = numroc(descb[N_], descb[NB_], mycol, descb[CSRC_], Q);
    return w;
}

}  // namespace slb
