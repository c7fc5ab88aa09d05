// dmma_probe2.cu -- DMMA.8x8x4 throughput vs register tiling (MF x NF accumulator grid per warp), operands in registers.
#include <cstdio>
#include <cuda_runtime.h>

template <int MF, int NF, int THREADS>
__global__ void __launch_bounds__(THREADS) probe(int iters, double *out, const double *in)
{
    double acc[MF][NF][2];
    double a[MF], b[NF];
#pragma unroll
    for (int i = 0; i < MF; ++i) a[i] = in[threadIdx.x + i];
#pragma unroll
    for (int j = 0; j < NF; ++j) b[j] = in[threadIdx.x + 64 + j];
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) { acc[i][j][0] = 0; acc[i][j][1] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
            for (int j = 0; j < NF; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(acc[i][j][0]), "+d"(acc[i][j][1]) : "d"(a[i]), "d"(b[j]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 123.456) out[0] = s;
}

template <int MF, int NF, int THREADS>
void run(int blocks_per_sm, int nsm, double *out, const double *in)
{
    int iters = 32768 / (MF * NF);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MF, NF, THREADS><<<nsm * blocks_per_sm, THREADS>>>(64, out, in);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); probe<MF, NF, THREADS><<<nsm * blocks_per_sm, THREADS>>>(iters, out, in); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double warps = (double)nsm * blocks_per_sm * THREADS / 32;
    double fl = warps * iters * (double)(MF * NF) * 256 * 2;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, probe<MF, NF, THREADS>);
    printf("MFxNF=%dx%d chains=%2d warps/SM=%2d regs=%3d : %8.3f ms  %6.2f TF\n", MF, NF, MF * NF, blocks_per_sm * THREADS / 32, fa.numRegs, best, fl / best / 1e9);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    double *out, *in; cudaMalloc(&out, 64); cudaMalloc(&in, 8192); cudaMemset(in, 0, 8192);
    run<1, 1, 256>(1, nsm, out, in); run<1, 2, 256>(1, nsm, out, in); run<2, 2, 256>(1, nsm, out, in); run<2, 3, 256>(1, nsm, out, in);
    run<2, 4, 256>(1, nsm, out, in); run<3, 4, 256>(1, nsm, out, in); run<4, 4, 256>(1, nsm, out, in); run<4, 8, 256>(1, nsm, out, in);
    run<8, 8, 256>(1, nsm, out, in); run<2, 16, 256>(1, nsm, out, in); run<4, 16, 256>(1, nsm, out, in);
    run<1, 2, 256>(2, nsm, out, in); run<2, 2, 256>(2, nsm, out, in); run<2, 4, 256>(2, nsm, out, in); run<4, 4, 256>(2, nsm, out, in);
    run<4, 8, 256>(2, nsm, out, in); run<2, 4, 512>(1, nsm, out, in); run<4, 4, 512>(1, nsm, out, in); run<2, 4, 1024>(1, nsm, out, in);
    run<2, 2, 1024>(1, nsm, out, in); run<4, 4, 128>(1, nsm, out, in); run<8, 8, 128>(1, nsm, out, in); run<4, 8, 128>(1, nsm, out, in);
    return 0;
}
