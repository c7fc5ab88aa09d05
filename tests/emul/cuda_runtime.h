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
#include <initializer_list>

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

// ---- "device" memory registry: what cudaMalloc returned is device memory, everything else is host memory.  Copies check their
// kind against it, kernel launches and the emulated NCCL check every pointer they are given, cudaPointerGetAttributes answers
// from it -- so a host pointer handed to a kernel, or a copy in the wrong direction, fails here as it would on a GPU.
#include <cstdio>
#include <map>
struct emul_registry {
    std::mutex mu; std::map<const char *, size_t> blocks;
    void add(void *p, size_t n) { std::lock_guard<std::mutex> lk(mu); blocks[(const char *)p] = n; }
    void del(void *p) { std::lock_guard<std::mutex> lk(mu); blocks.erase((const char *)p); }
    bool has(const void *p)
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = blocks.upper_bound((const char *)p);
        if (it == blocks.begin()) return false;
        --it;
        return (const char *)p < it->first + it->second;
    }
};
inline emul_registry emul_devmem;
[[noreturn]] static inline void emul_die(const char *what) { fprintf(stderr, "CUDA emulation: %s\n", what); fflush(stderr); abort(); }
static inline void emul_need_dev(const void *p, const char *what) { if (p && !emul_devmem.has(p)) emul_die(what); }
static inline void emul_need_host(const void *p, const char *what) { if (p && emul_devmem.has(p)) emul_die(what); }
static inline void emul_check_kind(const void *d, const void *s, int kind)
{
    if (kind == 1) { emul_need_dev(d, "HostToDevice copy: the destination is not device memory"); emul_need_host(s, "HostToDevice copy: the source is device memory"); }
    else if (kind == 2) { emul_need_dev(s, "DeviceToHost copy: the source is not device memory"); emul_need_host(d, "DeviceToHost copy: the destination is device memory"); }
    else if (kind == 3) { emul_need_dev(d, "DeviceToDevice copy: the destination is not device memory"); emul_need_dev(s, "DeviceToDevice copy: the source is not device memory"); }
    else if (kind == 0) { emul_need_host(d, "HostToHost copy of device memory"); emul_need_host(s, "HostToHost copy of device memory"); }
}
// every pointer argument of a kernel must be device memory (or null)
static inline void emul_check_arg(...) {}
template <typename T> static inline void emul_check_arg(T *p) { emul_need_dev((const void *)p, "a kernel was given a pointer that is not device memory"); }
template <typename... A> static inline void emul_check_args(A... a) { (void)std::initializer_list<int>{ (emul_check_arg(a), 0)... }; }
static inline void emul_check_geometry(const dim3 &g, const dim3 &b)
{
    if (g.x < 1 || g.y < 1 || g.z < 1 || b.x < 1 || b.y < 1 || b.z < 1) emul_die("kernel launch with an empty grid or block (invalid configuration)");
    if ((uint64_t)b.x * b.y * b.z > 1024 || g.y > 65535 || g.z > 65535 || g.x > 2147483647u) emul_die("kernel launch geometry out of range");
}

static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p)
{ a->type = emul_devmem.has(p) ? cudaMemoryTypeDevice : cudaMemoryTypeUnregistered; a->device = 0; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); if (*p) { memset(*p, 0xA5, n ? n : 1); emul_devmem.add(*p, n ? n : 1); } return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaFree(void *p) { if (p) { emul_devmem.del(p); free(p); } return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k) { if (n) { emul_check_kind(d, s, (int)k); memmove(d, s, n); } return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
static inline cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind k)
{
    if (w == 0 || h == 0) return cudaSuccess;
    if (dp < w || sp < w) emul_die("cudaMemcpy2D: a pitch is smaller than the row width");
    emul_check_kind(d, s, (int)k);
    for (size_t j = 0; j < h; ++j) memmove((char *)d + j * dp, (const char *)s + j * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t)
{ return cudaMemcpy2D(d, dp, s, sp, w, h, k); }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) { emul_need_dev(d, "cudaMemset of host memory"); memset(d, v, n); } return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { return cudaMemset(d, v, n); }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
// events are tokens only: the emulation executes everything in program order, so every recorded event has completed
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t *fr, size_t *tot) { *fr = (size_t)64 << 30; *tot = (size_t)128 << 30; return cudaSuccess; }
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
        emul_check_geometry(g_, b_); emul_check_args(__VA_ARGS__);                                           \
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
        emul_check_geometry(g_, b_); emul_check_args(__VA_ARGS__);                                           \
        for (unsigned bz_ = 0; bz_ < g_.z; ++bz_) for (unsigned by_ = 0; by_ < g_.y; ++by_) for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) \
            emul_run_block(g_, b_, bx_, by_, bz_, [&] { kernel(__VA_ARGS__); });                               \
        ::slb::counter_add("kernel_launches", 1);                                                            \
    } while (0)
