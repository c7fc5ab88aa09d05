// devmath.cuh -- scalar helpers shared by the real (double) and complex (double2) kernels.
#pragma once
#include <cuda_runtime.h>

namespace slb {

typedef double2 zcomplex;

__host__ __device__ __forceinline__ double t_zero(double) { return 0.0; }
__host__ __device__ __forceinline__ zcomplex t_zero(zcomplex) { return make_double2(0.0, 0.0); }

// pivot metric: |x| for real (idamax_), |Re|+|Im| for complex (izamax_, PBLAS/SRC/pzamax_.c:494-497)
__device__ __forceinline__ double t_abs1(double x) { return fabs(x); }
__device__ __forceinline__ double t_abs1(zcomplex z) { return fabs(z.x) + fabs(z.y); }
__device__ __forceinline__ bool t_iszero(double x) { return x == 0.0; }
__device__ __forceinline__ bool t_iszero(zcomplex z) { return z.x == 0.0 && z.y == 0.0; }

__device__ __forceinline__ double t_mul(double a, double b) { return a * b; }
__device__ __forceinline__ zcomplex t_mul(zcomplex a, zcomplex b)
{ return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// c - a*b
__device__ __forceinline__ double t_fnma(double a, double b, double c) { return fma(-a, b, c); }
__device__ __forceinline__ zcomplex t_fnma(zcomplex a, zcomplex b, zcomplex c)
{
    // c - a*b, each component accumulated with fused multiply-adds
    double re = fma(-a.x, b.x, c.x); re = fma(a.y, b.y, re);
    double im = fma(-a.x, b.y, c.y); im = fma(-a.y, b.x, im);
    return make_double2(re, im);
}
// reciprocal ONE/GMAX (SRC/pdgetf2.f:224); complex: Smith's division (what a Fortran (1,0)/z evaluates to)
__device__ __forceinline__ double t_recip(double x) { return 1.0 / x; }
__device__ __forceinline__ zcomplex t_recip(zcomplex z)
{
    if (fabs(z.x) >= fabs(z.y)) { double t = z.y / z.x, d = z.x + z.y * t; return make_double2(1.0 / d, -t / d); }
    double t = z.x / z.y, d = z.x * t + z.y;
    return make_double2(t / d, -1.0 / d);
}
__device__ __forceinline__ double t_add(double a, double b) { return a + b; }
__device__ __forceinline__ zcomplex t_add(zcomplex a, zcomplex b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double t_sub(double a, double b) { return a - b; }
__device__ __forceinline__ zcomplex t_sub(zcomplex a, zcomplex b) { return make_double2(a.x - b.x, a.y - b.y); }

__device__ __forceinline__ double t_conj(double a) { return a; }
__device__ __forceinline__ zcomplex t_conj(zcomplex a) { return make_double2(a.x, -a.y); }

// L1-bypassing loads for data another CTA / GPU may have just written
__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }
__device__ __forceinline__ zcomplex ld_cg(const zcomplex *p) { return __ldcg(p); }
__device__ __forceinline__ int ld_cg(const int *p) { return __ldcg(p); }

}  // namespace slb
