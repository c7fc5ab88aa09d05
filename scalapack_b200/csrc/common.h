// common.h -- shared declarations of the B200-native ScaLAPACK LU library.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/scalapack_b200.h"

namespace slb {

// ---- descriptor field indices (0-based; TOOLS/descinit.f:130-134) -----------
enum { DTYPE_ = 0, CTXT_ = 1, M_ = 2, N_ = 3, MB_ = 4, NB_ = 5, RSRC_ = 6, CSRC_ = 7, LLD_ = 8 };

// ---- fatal error helpers -----------------------------------------------------
[[noreturn]] void fatal(const char *fmt, ...);
void vlog(int level, const char *fmt, ...);
int verbose();

#define SLB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            ::slb::fatal("CUDA error %s at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

// ---- integer index algebra (TOOLS/*.f), 1-based like the reference -----------
int numroc(int n, int nb, int iproc, int isrc, int nprocs);
inline int indxg2p(int ig, int nb, int isrc, int nprocs) { return (isrc + (ig - 1) / nb) % nprocs; }
inline int indxg2l(int ig, int nb, int nprocs) { return nb * ((ig - 1) / (nb * nprocs)) + (ig - 1) % nb + 1; }
inline int indxl2g(int il, int nb, int iproc, int isrc, int nprocs)
{ return nprocs * nb * ((il - 1) / nb) + (il - 1) % nb + ((nprocs + iproc - isrc) % nprocs) * nb + 1; }
void infog2l(int gr, int gc, const int *desc, int nprow, int npcol, int myrow, int mycol, int *lr, int *lc,
             int *rsrc, int *csrc);
void chk1mat(int ma, int mapos0, int na, int napos0, int ia, int ja, const int *desc, int descpos0, int *info);

// ---- host control plane (replaces MPI for setup / tiny combines) -------------
struct HostComm;
HostComm *hostcomm();                 // lazily bootstrapped singleton
int  hc_rank();
int  hc_size();
// all-gather `len` bytes among `nmembers` participants identified by `group` (any 64-bit key that all
// members agree on); `index` = my position in the group.  out must hold nmembers*len bytes.
void hc_allgather(uint64_t group, int nmembers, int index, const void *in, void *out, size_t len);
void hc_shutdown();

// ---- BLACS grid context --------------------------------------------------------
struct NcclComms;   // ncclw.h
struct Grid {
    bool valid = false;
    int ctxt = -1;
    int uid = 0;                  // creation counter: identical on all members (gridinit is collective)
    int nprow = 0, npcol = 0, myrow = -1, mycol = -1;
    std::vector<int> pmap;        // pmap[r * npcol + c] = world rank
    uint64_t seq_all = 0, seq_row = 0, seq_col = 0;   // per-scope collective sequence numbers
    NcclComms *nccl = nullptr;    // created lazily on first multi-GPU factor/solve
    bool in_grid() const { return valid && myrow >= 0; }
};
Grid *grid_of(int ictxt);         // nullptr if invalid
// scoped host collectives over a grid: scope 'A','R','C'
void grid_allgather(Grid *g, char scope, const void *in, void *out, size_t len);
int  grid_scope_size(Grid *g, char scope);
int  grid_scope_index(Grid *g, char scope);
void grid_barrier(Grid *g, char scope);
int  grid_imin(Grid *g, char scope, int v);
int  grid_imax(Grid *g, char scope, int v);

// ---- device / runtime ------------------------------------------------------------
struct Runtime {
    bool cuda_ok = false;
    int device = -1;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t s_main = nullptr;     // trailing update / swaps
    cudaStream_t s_panel = nullptr;    // look-ahead panel stream (high priority)
    cudaStream_t s_copy = nullptr;     // staging copies
    cudaStream_t s_d2h = nullptr;      // copy engine, device -> host: finished block rows of a host-resident caller (stage.cu)
    cudaStream_t s_h2d = nullptr;      // copy engine, host -> device: column slabs of a host-resident caller (stage.cu)
    cudaStream_t s_aux = nullptr;      // carries no work: joins several events into one (cudaStreamWaitEvent x n + cudaEventRecord)
    cudaStream_t s_prep = nullptr;     // row interchanges + U12 solve of the next column half (above the update, below the panel)
};
Runtime &rt();                      // initialises CUDA lazily; fatal()s if no device (no CPU fallback)
bool cuda_available();
// named device workspace, grown on demand and kept across calls
void *workspace(const char *name, size_t bytes, bool zero_on_alloc = false);
void workspace_release_all();

// ---- options & counters -------------------------------------------------------------
int64_t opt(const char *key, int64_t dflt);
void counter_add(const char *key, int64_t v);

}  // namespace slb
