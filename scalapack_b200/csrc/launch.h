// launch.h -- one spelling for a kernel launch in the files of the SURVEY 8(f) rows (refine.cu, chol.cu, inverse.cu, redist.cu).
//
// Product build (nvcc): a plain <<< >>> launch + error check + launch counter.
// Host-logic test build (tests/emul, g++ with a stub cuda_runtime.h that defines SLB_EMUL): the same kernel body is run
// thread by thread in a serial loop, so the index arithmetic of these kernels and of the drivers above them is checked
// on a CPU-only machine.  Kernels launched through SLB_LAUNCH therefore use no shared memory, barriers, shuffles or
// atomics, and no thread reads an element another thread of the same launch writes; SLB_LAUNCH_SYNC additionally allows
// __syncthreads() and static __shared__ arrays (small blocks: the emulation spends one OS thread per CUDA thread).  The emulation is test
// infrastructure only: the product library contains no CPU path and aborts without a GPU (runtime.cu).
#pragma once
#include "common.h"

#ifndef SLB_EMUL
#define SLB_LAUNCH(kernel, grid, block, stream, ...)                      \
    do {                                                                  \
        kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);            \
        SLB_CUDA(cudaGetLastError());                                     \
        ::slb::counter_add("kernel_launches", 1);                         \
    } while (0)
// a kernel that uses __syncthreads() / static __shared__ arrays: the emulation runs one OS thread per CUDA thread of a block
#define SLB_LAUNCH_SYNC SLB_LAUNCH
#endif

namespace slb {
// 1-D grid of 256-thread blocks covering n items, capped so that very long arrays are walked with a grid-stride loop
inline unsigned grid1d(int64_t n, int block = 256, unsigned cap = 148 * 32)
{
    int64_t b = (n + block - 1) / block;
    if (b < 1) b = 1;
    return (unsigned)(b < (int64_t)cap ? b : cap);
}
}  // namespace slb
