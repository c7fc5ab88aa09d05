// lu.h -- device-level drivers of the LU path (templated on double / zcomplex).
#pragma once
#include "common.h"
#include "kernels.cuh"

namespace slb {

struct LuStats {
    double factor_ms = 0, solve_ms = 0;
    double update_ms = 0, update_flops = 0;
    int64_t update_launches = 0;
    bool host_written = false;   // the factorisation already wrote the factors to the caller's host array (e2e_overlap)
};
extern LuStats g_last_lu;

// A: device pointer to the local block-cyclic array (lld x LOCc(N)), IA = JA = 1.
// ipiv_glob_host: min(M,N) ints, 1-based global pivot rows (replicated on every process).
// host_out (optional): the caller's host-resident array A was staged from; with SLB200_E2E_OVERLAP=1 (experimental, 1x1
// grid) finished block rows are written back to it during the factorisation and g_last_lu.host_written is set.
template <typename T>
int getrf_device(Grid *g, int M, int N, T *A, int64_t lld, int nb, int rsrc, int csrc, int *ipiv_glob_host, int *info_host,
                 T *host_out = nullptr);

// B: device pointer to the local block-cyclic right-hand sides (lldb x LOCc(NRHS)), row blocking nb, column
// blocking nbb, sources (rsrc, csrcb).  ipiv_glob_host: N ints (1-based global).  trans: 'N' only.
template <typename T>
int getrs_device(Grid *g, int N, int nrhs, const T *A, int64_t lld, int nb, int rsrc, int csrc, const int *ipiv_glob_host,
                 T *B, int64_t lldb, int nbb, int csrcb);

template <typename T>
void launch_gather_rows(int64_t n, const int *perm, const T *src, int64_t lds, T *dst, int64_t ldd, int nrhs, cudaStream_t s);

}  // namespace slb
