// dmma_probe.cu -- FP64 throughput probes on B200: DMMA.8x8x4 vs DFMA vs both interleaved,
// as a function of warps per SM and independent chains per warp.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int NMMA, int NFMA>
__global__ void __launch_bounds__(1024) probe(int iters, double *out)
{
    double acc[NMMA > 0 ? NMMA : 1][2];
    double f[NFMA > 0 ? NFMA : 1];
#pragma unroll
    for (int i = 0; i < (NMMA > 0 ? NMMA : 1); ++i) { acc[i][0] = 0; acc[i][1] = 0; }
#pragma unroll
    for (int i = 0; i < (NFMA > 0 ? NFMA : 1); ++i) f[i] = threadIdx.x + i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9, c = 1e-9, d = 1.0000001;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NMMA; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
#pragma unroll
            for (int i = 0; i < NFMA; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;\n" : "+d"(f[i]) : "d"(d), "d"(c));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < (NMMA > 0 ? NMMA : 1); ++i) s += acc[i][0] + acc[i][1];
#pragma unroll
    for (int i = 0; i < (NFMA > 0 ? NFMA : 1); ++i) s += f[i];
    if (s == 123.456) out[0] = s;
}

template <int NMMA, int NFMA>
void run(int threads, int blocks_per_sm, int nsm, double *out)
{
    int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<NMMA, NFMA><<<nsm * blocks_per_sm, threads>>>(100, out);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); probe<NMMA, NFMA><<<nsm * blocks_per_sm, threads>>>(iters, out); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double warps = (double)nsm * blocks_per_sm * threads / 32;
    double mma_fl = warps * iters * 4.0 * NMMA * 256 * 2, fma_fl = warps * iters * 4.0 * NFMA * 32 * 2;
    printf("NMMA=%2d NFMA=%2d warps/SM=%2d : %8.3f ms  dmma %6.2f TF  dfma %6.2f TF  total %6.2f TF\n", NMMA, NFMA,
           blocks_per_sm * threads / 32, best, mma_fl / best / 1e9, fma_fl / best / 1e9, (mma_fl + fma_fl) / best / 1e9);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    double *out; cudaMalloc(&out, 64);
    printf("%s SMs=%d clock=%d kHz\n", p.name, nsm, p.clockRate);
    for (int w : {4, 8, 16, 32}) { run<1, 0>(w * 32, 1, nsm, out); run<2, 0>(w * 32, 1, nsm, out); run<4, 0>(w * 32, 1, nsm, out); run<8, 0>(w * 32, 1, nsm, out); run<16, 0>(w * 32, 1, nsm, out); }
    for (int w : {4, 8, 16, 32}) { run<0, 4>(w * 32, 1, nsm, out); run<0, 8>(w * 32, 1, nsm, out); run<0, 16>(w * 32, 1, nsm, out); }
    for (int w : {8, 16, 32}) { run<8, 8>(w * 32, 1, nsm, out); run<8, 16>(w * 32, 1, nsm, out); run<8, 32>(w * 32, 1, nsm, out); run<4, 32>(w * 32, 1, nsm, out); run<8, 64>(w * 32, 1, nsm, out); }
    return 0;
}
