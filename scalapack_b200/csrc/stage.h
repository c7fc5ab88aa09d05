// stage.h -- asynchronous 2-D transfers between a HOST-resident local array (the drop-in case: an unmodified
// Fortran caller owns A in host memory, SURVEY.md 8b "Ownership") and its staging copy in HBM.
//
//   * page-locked caller memory (cudaHostAlloc / cudaHostRegister): cudaMemcpy2DAsync straight between the
//     caller's array and HBM on the copy-engine streams;
//   * pageable caller memory: a ring of page-locked bounce buffers; worker threads move caller <-> bounce with
//     parallel memcpy while the copy engine moves bounce <-> HBM (CUDA's own pageable path is one thread).
// Every transfer is a TICKET: the factorisation polls / waits for upload tickets (column slabs join the
// right-looking sweep as they arrive, lu.cu) and hands finished block rows to download tickets.
#pragma once
#include "common.h"

#include <atomic>
#include <memory>

namespace slb {

struct HostMat {             // a column-major host array: element (i, j) at p + (i + j*ld) * elem
    void *p = nullptr;
    int64_t ld = 0;          // leading dimension in ELEMENTS
    int64_t rows = 0, cols = 0;
    size_t elem = 8;
    bool pinned = false;     // page-locked (DMA-able in place)
};

bool host_ptr_is_pinned(const void *p);

class HostLink {
public:
    // dev: device array with leading dimension ldd (elements) mirroring `h` (same rows x cols window)
    HostLink(const HostMat &h, void *dev, int64_t ldd);
    ~HostLink();
    // rows [r0, r1) x columns [c0, c1) host -> device.  Returns a ticket.
    int upload(int64_t r0, int64_t r1, int64_t c0, int64_t c1);
    // device -> host once `after` (an event the caller has ALREADY recorded; may be null) has completed
    int download(int64_t r0, int64_t r1, int64_t c0, int64_t c1, cudaEvent_t after);
    bool done(int ticket);                       // non-blocking: the transfer has completed
    void wait(int ticket);                       // block the calling host thread
    void stream_wait(int ticket, cudaStream_t s);   // make stream s wait (host blocks briefly until the ticket's event exists)
    void finish();                               // wait for every ticket issued so far
    int64_t bytes_up() const { return up_bytes_; }
    int64_t bytes_down() const { return down_bytes_; }
    bool pinned() const { return h_.pinned; }
private:
    struct Impl;
    std::unique_ptr<Impl> im_;
    HostMat h_;
    void *dev_; int64_t ldd_;
    int64_t up_bytes_ = 0, down_bytes_ = 0;
};

}  // namespace slb
