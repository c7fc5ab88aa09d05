// cuda_runtime.h (STUB) -- test infrastructure, never part of the product.
//
// tests/emul builds the HOST logic of the SURVEY 8(f) files (refine.cu, chol.cu, inverse.cu, redist.cu, ...) with g++ against
// this header instead of the CUDA runtime: "device" memory is host memory, streams are synchronous, and a kernel launched
// through SLB_LAUNCH (csrc/launch.h) is run thread by thread in a serial loop.  That checks, on a CPU-only machine, the index
// arithmetic of those kernels, the block loops of the drivers above them and the order of their collectives on P x Q grids.
// It deliberately offers no shared memory, barriers, shuffles or atomics: a kernel that needs them cannot go through SLB_LAUNCH.
// The kernels of the LU hot path are NOT emulated; tests/emul/backend.cpp stands in for them by their documented contract.
#pragma once
#define SLB_EMUL 1
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct emul_uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
inline thread_local emul_uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T> static inline T __ldg(const T *p) { return *p; }

typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef struct emul_stream *cudaStream_t;
typedef struct emul_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; };
enum { cudaStreamNonBlocking = 1 };

static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { a->type = cudaMemoryTypeUnregistered; a->device = 0; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
static inline cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind)
{ for (size_t j = 0; j < h; ++j) memmove((char *)d + j * dp, (const char *)s + j * sp, w); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t)
{ return cudaMemcpy2D(d, dp, s, sp, w, h, k); }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }

// serial execution of a barrier-free kernel: every block, every thread, one after the other
#define SLB_LAUNCH(kernel, grid, block, stream, ...)                                                         \
    do {                                                                                                     \
        const dim3 g_ = (grid), b_ = (block);                                                                \
        (void)(stream);                                                                                      \
        gridDim = g_; blockDim = b_;                                                                         \
        for (unsigned bz_ = 0; bz_ < g_.z; ++bz_) for (unsigned by_ = 0; by_ < g_.y; ++by_) for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) { \
            blockIdx.x = bx_; blockIdx.y = by_; blockIdx.z = bz_;                                            \
            for (unsigned tz_ = 0; tz_ < b_.z; ++tz_) for (unsigned ty_ = 0; ty_ < b_.y; ++ty_) for (unsigned tx_ = 0; tx_ < b_.x; ++tx_) { \
                threadIdx.x = tx_; threadIdx.y = ty_; threadIdx.z = tz_;                                     \
                kernel(__VA_ARGS__);                                                                         \
            }                                                                                                \
        }                                                                                                    \
        ::slb::counter_add("kernel_launches", 1);                                                            \
    } while (0)

// ---- kernels with block-wide barriers and static shared memory (SLB_LAUNCH_SYNC) -------------------------------------------
// One OS thread per CUDA thread of a block, blocks one after the other; __syncthreads() is a real barrier between them and a
// `__shared__` array is a function-static one (shared by the threads of the running block).  Meant for SMALL blocks in tests.
#define __shared__ static
struct emul_barrier {
    std::mutex mu; std::condition_variable cv; unsigned n = 1, waiting = 0, phase = 0;
    void wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned ph = phase;
        if (++waiting == n) { waiting = 0; ++phase; cv.notify_all(); }
        else cv.wait(lk, [&] { return phase != ph; });
    }
};
inline emul_barrier emul_block_barrier;
static inline void __syncthreads() { emul_block_barrier.wait(); }
template <typename F>
static inline void emul_run_block(const dim3 &g, const dim3 &b, unsigned bx, unsigned by, unsigned bz, F body)
{
    const unsigned nt = b.x * b.y * b.z;
    emul_block_barrier.n = nt; emul_block_barrier.waiting = 0;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([=] {
            gridDim = g; blockDim = b; blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
            threadIdx.x = t % b.x; threadIdx.y = (t / b.x) % b.y; threadIdx.z = t / (b.x * b.y);
            body();
        });
    for (auto &t : th) t.join();
}
#define SLB_LAUNCH_SYNC(kernel, grid, block, stream, ...)                                                    \
    do {                                                                                                     \
        const dim3 g_ = (grid), b_ = (block);                                                                \
        (void)(stream);                                                                                      \
        for (unsigned bz_ = 0; bz_ < g_.z; ++bz_) for (unsigned by_ = 0; by_ < g_.y; ++by_) for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) \
            emul_run_block(g_, b_, bx_, by_, bz_, [&] { kernel(__VA_ARGS__); });                               \
        ::slb::counter_add("kernel_launches", 1);                                                            \
    } while (0)
