// ncclw.h -- NCCL loaded at run time (dlopen) so the library also loads on a CPU-only box and on
// 1x1 grids that never communicate.  Communicators mirror the three BLACS scopes of a grid
// (BLACS/SRC/blacs_map_.c:106-118): all, row (color=myrow,key=mycol), column (color=mycol,key=myrow).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace slb {

struct Grid;
typedef struct ncclComm *ncclComm_t_;

struct NcclComms {
    ncclComm_t_ all = nullptr, row = nullptr, col = nullptr;
    ncclComm_t_ colp = nullptr;     // second column communicator: panel phase on the look-ahead stream
};

enum NcclType { NT_U8 = 1, NT_I32 = 2, NT_F64 = 8 };   // values of ncclDataType_t

NcclComms *nccl_create(Grid *g);        // collective over the grid; fatal()s if NCCL cannot be loaded
void nccl_destroy(NcclComms *c);

// thin wrappers; rank arguments are ranks inside the given communicator
void nccl_group_start();
void nccl_group_end();
void nccl_bcast(ncclComm_t_ comm, void *buf, size_t count, NcclType t, int root, cudaStream_t s);
void nccl_send(ncclComm_t_ comm, const void *buf, size_t count, NcclType t, int peer, cudaStream_t s);
void nccl_recv(ncclComm_t_ comm, void *buf, size_t count, NcclType t, int peer, cudaStream_t s);
void nccl_allgather(ncclComm_t_ comm, const void *send, void *recv, size_t sendcount, NcclType t, cudaStream_t s);
void nccl_allreduce_min_i32(ncclComm_t_ comm, const void *send, void *recv, size_t count, cudaStream_t s);
void nccl_allreduce_sum_f64(ncclComm_t_ comm, const void *send, void *recv, size_t count, cudaStream_t s);
void nccl_allreduce_max_f64(ncclComm_t_ comm, const void *send, void *recv, size_t count, cudaStream_t s);
// personalised all-to-all in BYTES (PDGEMR2D): peer p receives send[sdispl[p], sdispl[p] + scount[p]) and its part lands in
// recv[rdispl[p], rdispl[p] + rcount[p]); one group of ncclSend / ncclRecv, the own part is a device copy.  np / me: size of the
// communicator and my rank in it.
void nccl_alltoallv(ncclComm_t_ comm, int np, int me, const void *send, const size_t *scount, const size_t *sdispl, void *recv,
                    const size_t *rcount, const size_t *rdispl, cudaStream_t s);
const char *nccl_version_string();

}  // namespace slb
