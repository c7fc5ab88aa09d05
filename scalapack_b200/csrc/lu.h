// lu.h -- device-level drivers of the LU path (templated on double / zcomplex).
#pragma once
#include "common.h"
#include "kernels.cuh"

namespace slb {

struct LuStats {
    double factor_ms = 0, solve_ms = 0;
    double update_ms = 0, update_flops = 0;
    int64_t update_launches = 0;
    bool host_written = false;   // the factorisation already wrote the factors to the caller's host array (link)
};
extern LuStats g_last_lu;

// A: device pointer to the local block-cyclic array (lld x LOCc(N)), IA = JA = 1.
// ipiv_glob_host: min(M,N) ints, 1-based global pivot rows (replicated on every process).
// link (optional): A is the (uninitialised) HBM staging copy of a HOST-resident caller's array; the factorisation uploads
// it through `link` (column slabs that join the sweep as they arrive) and writes the factors back (block rows as they
// become final); on return the caller's host array holds the factors and g_last_lu.host_written is set.
class HostLink;
template <typename T>
int getrf_device(Grid *g, int M, int N, T *A, int64_t lld, int nb, int rsrc, int csrc, int *ipiv_glob_host, int *info_host,
                 HostLink *link = nullptr);

// PDGETRS (SRC/pdgetrs.f:244-286).  trans = 'N': sub(A) X = sub(B); 'T': sub(A)^T X = sub(B); 'C': sub(A)^H X = sub(B).
// A: device pointer to the local window of the factors (lld x LOCc(N)); B: device pointer to the first local row of sub(B)
// in the local array of B (lldb x nlocB_all columns of the WHOLE B, column blocking nbb, source csrcb); sub(B) occupies the
// global columns [jb0, jb0 + nrhs) of B.  ipiv_glob_host: N ints, 1-based, relative to sub(A).
template <typename T>
int getrs_device(Grid *g, char trans, int N, int nrhs, const T *A, int64_t lld, int nb, int rsrc, int csrc, const int *ipiv_glob_host,
                 T *B, int64_t lldb, int nbb, int csrcb, int jb0, int64_t nlocB_all,
                 const T *Xrep_in = nullptr, T *Xrep_out = nullptr);
// Xrep_in / Xrep_out (device, N x nrhs, ld = N, global row order, identical on every process): when given, the right-hand
// sides are taken from / the solutions are left in these replicated blocks instead of the block-cyclic B (which may be null).

// 1 x 1 grid, TRANS = 'N', nb <= 512, few right-hand sides: the two sweeps at HBM speed (solve_fast.cu)
bool getrs_fast_applies(int P, int Q, char trans, int nb, int nrhs);
void getrs_fast_device(int N, int nrhs, const double *A, int64_t lld, int nb, double *Xg);
double solve_fast_probe(int which, int nb, int64_t nr, const double *A, int64_t lld, int N, int reps);

template <typename T>
void launch_gather_rows(int64_t n, const int *perm, const T *src, int64_t lds, T *dst, int64_t ldd, int nrhs, cudaStream_t s,
                        bool scatter = false);

}  // namespace slb
