// dmma_probe3.cu -- does feeding DMMA operands from shared memory (LDS.64 fragments, GEMM-like 64x32 warp tile) lower the pipe rate?
#include <cstdio>
#include <cuda_runtime.h>
constexpr int SA = 132, SB = 20, BK = 16;
template <int MODE, int WARPS>   // MODE 0: operands in registers (loop invariant); 1: reloaded from smem every k4, double buffered; 2: as 1 + __syncthreads per 4 k4
__global__ void __launch_bounds__(WARPS * 32, 1) probe(int iters, double *out)
{
    extern __shared__ double sm[];
    double *as = sm, *bs = sm + 4 * BK * SA;
    for (int i = threadIdx.x; i < 4 * BK * SA + 4 * 128 * SB; i += blockDim.x) sm[i] = 1e-3 * (i % 7);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    const int wm0 = (warp & 1) * 64, wn0 = ((warp >> 1) & 3) * 32;
    double acc[2][8][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[i][j][v] = 0;
    double fa[2][2][2], fb[2][8];
    auto load = [&](int buf, int st, int k4) {
        const double *a = as + st * BK * SA, *b = bs + st * 128 * SB;
#pragma unroll
        for (int nf = 0; nf < 2; ++nf) { fa[buf][nf][0] = b[(wn0 + nf * 16 + g) * SB + k4 + tig]; fa[buf][nf][1] = b[(wn0 + nf * 16 + g + 8) * SB + k4 + tig]; }
#pragma unroll
        for (int mf = 0; mf < 8; ++mf) fb[buf][mf] = a[(k4 + tig) * SA + wm0 + mf * 8 + g];
    };
    load(0, 0, 0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            const int cur = s4 & 1, nxt = cur ^ 1;
            if (MODE >= 1) load(nxt, (it + (s4 == 3)) & 3, ((s4 + 1) & 3) * 4);
            if (MODE == 2 && s4 == 1) __syncthreads();
#pragma unroll
            for (int nf = 0; nf < 2; ++nf)
#pragma unroll
                for (int mf = 0; mf < 8; ++mf)
                    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                                 : "+d"(acc[nf][mf][0]), "+d"(acc[nf][mf][1]), "+d"(acc[nf][mf][2]), "+d"(acc[nf][mf][3])
                                 : "d"(fa[MODE ? cur : 0][nf][0]), "d"(fa[MODE ? cur : 0][nf][1]), "d"(fb[MODE ? cur : 0][mf]));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j][0] + acc[i][j][1] + acc[i][j][2] + acc[i][j][3];
    if (s == 123.456) out[0] = s;
}
template <int MODE, int WARPS> void run(int nsm, double *out)
{
    size_t smem = (4 * BK * SA + 4 * 128 * SB) * 8;
    cudaFuncSetAttribute(probe<MODE, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE, WARPS><<<nsm, WARPS * 32, smem>>>(50, out); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); probe<MODE, WARPS><<<nsm, WARPS * 32, smem>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double fl = (double)nsm * WARPS * iters * 4.0 * 16 * 512 * 2;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, probe<MODE, WARPS>);
    printf("MODE=%d warps=%d regs=%d: %.3f ms %.2f TF  (%s)\n", MODE, WARPS, fa.numRegs, best, fl / best / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int nsm = p.multiProcessorCount;
    double *out; cudaMalloc(&out, 64);
    run<0, 8>(nsm, out); run<1, 8>(nsm, out); run<2, 8>(nsm, out); run<0, 4>(nsm, out); run<1, 4>(nsm, out);
    return 0;
}
