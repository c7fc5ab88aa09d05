// runtime.cu -- device binding, streams, workspace cache.  There is NO CPU fallback: every compute entry
// point goes through rt(), which aborts loudly when no sm_100 device is usable.
#include "common.h"
#include "kernels.cuh"

#include <map>
#include <mutex>

namespace slb {

static Runtime g_rt;
static std::mutex g_rtmu;

bool cuda_available()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return n > 0;
}

Runtime &rt()
{
    std::lock_guard<std::mutex> lk(g_rtmu);
    if (g_rt.cuda_ok) {
        // a caller's other thread may not have this process's GPU current yet
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cur != g_rt.device) SLB_CUDA(cudaSetDevice(g_rt.device));
        return g_rt;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        fatal("no CUDA device: the LU path runs only on the GPU (hand-written sm_100a kernels); there is no CPU fallback");
    int dev = 0;
    const char *lr = getenv("LOCAL_RANK");
    if (!lr) lr = getenv("OMPI_COMM_WORLD_LOCAL_RANK");
    if (lr && *lr) dev = atoi(lr) % n; else dev = hc_rank() % n;
    SLB_CUDA(cudaSetDevice(dev));
    cudaDeviceProp p; SLB_CUDA(cudaGetDeviceProperties(&p, dev));
    if (p.major < 10) fatal("device %d is sm_%d%d; this library contains sm_100a code only", dev, p.major, p.minor);
    g_rt.device = dev;
    g_rt.sm_count = p.multiProcessorCount;
    g_rt.smem_optin = p.sharedMemPerBlockOptin;
    int lo = 0, hi = 0; SLB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SLB_CUDA(cudaStreamCreateWithPriority(&g_rt.s_main, cudaStreamNonBlocking, lo));
    SLB_CUDA(cudaStreamCreateWithPriority(&g_rt.s_panel, cudaStreamNonBlocking, hi));
    SLB_CUDA(cudaStreamCreateWithPriority(&g_rt.s_copy, cudaStreamNonBlocking, hi < lo - 1 ? hi + 1 : hi));
    SLB_CUDA(cudaStreamCreateWithPriority(&g_rt.s_d2h, cudaStreamNonBlocking, lo));
    SLB_CUDA(cudaStreamCreateWithPriority(&g_rt.s_h2d, cudaStreamNonBlocking, lo));
    SLB_CUDA(cudaStreamCreateWithPriority(&g_rt.s_aux, cudaStreamNonBlocking, hi));
    SLB_CUDA(cudaStreamCreateWithPriority(&g_rt.s_prep, cudaStreamNonBlocking, hi < lo - 1 ? hi + 1 : hi));
    g_rt.cuda_ok = true;
    return g_rt;
}

// ---- workspace cache: device buffers keyed by name, grown on demand, kept across calls ----------------
struct WsEntry { void *p = nullptr; size_t bytes = 0; };
static std::map<std::string, WsEntry> g_ws;

void *workspace(const char *name, size_t bytes, bool zero_on_alloc)
{
    rt();
    std::lock_guard<std::mutex> lk(g_rtmu);
    WsEntry &e = g_ws[name];
    if (e.bytes < bytes) {
        if (e.p) { SLB_CUDA(cudaDeviceSynchronize()); SLB_CUDA(cudaFree(e.p)); }
        size_t nb = bytes + (bytes >> 3) + 256;
        cudaError_t err = cudaMalloc(&e.p, nb);
        if (err != cudaSuccess) fatal("cudaMalloc(%zu bytes) for workspace '%s' failed: %s", nb, name, cudaGetErrorString(err));
        e.bytes = nb;
        if (zero_on_alloc) SLB_CUDA(cudaMemset(e.p, 0, nb));
    }
    return e.p;
}

void workspace_release_all()
{
    std::lock_guard<std::mutex> lk(g_rtmu);
    for (auto &kv : g_ws) if (kv.second.p) cudaFree(kv.second.p);
    g_ws.clear();
}

}  // namespace slb

extern "C" int slb200_has_cuda(void)
{
    if (!slb::cuda_available()) return 0;
    int dev = 0; cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return p.major >= 10 ? 1 : 0;
}
extern "C" int slb200_device(void) { return slb::cuda_available() ? slb::rt().device : -1; }
