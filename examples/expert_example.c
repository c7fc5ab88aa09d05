/* expert_example.c -- a compiled caller of the routines around the LU, written against include/scalapack_b200.h only:
 * the flow of the reference's test driver with EST = T (TESTING/traditional/LIN/pdludriver.f:561-830: PDGESVX = equilibrate, factor,
 * PDGECON, solve, PDGERFS), then PDGEMR2D to another block size, PDGETRI on the factors and PDPOSV on A'A.
 *
 *   gcc examples/expert_example.c -Iinclude -Lscalapack_b200/lib -lscalapack_b200 -Wl,-rpath,$PWD/scalapack_b200/lib -lm -o expert_example
 *   ./expert_example                     (one process = a 1 x 1 grid; needs a B200: the library has no CPU fallback) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "scalapack_b200.h"

#define N 96
#define NB 32
#define NRHS 2

int main(void)
{
    int me, np, ictxt, what = 0, minus1 = -1, one = 1, zero = 0, n = N, nrhs = NRHS, nb = NB, nb2 = 16, lld = N, info, nprow = 1, npcol = 1, myrow, mycol;
    int desca[9], descb[9], descc[9], ipiv[N + NB], iwork[4 * N], lwork = N * NB + 16 * N, liwork = 4 * N;
    static double a[N * N], af[N * N], a0[N * N], c[N * N], b[N * NRHS], x[N * NRHS], r[N], cs[N], work[N * NB + 16 * N], ferr[NRHS], berr[NRHS], g[N * N], bb[N * NRHS];
    double rcond, anorm, err = 0.0, alpha = 1.0, beta = 0.0;
    char equed = 'N';
    unsigned long long s = 12345;
    blacs_pinfo_(&me, &np);
    blacs_get_(&minus1, &what, &ictxt);
    blacs_gridinit_(&ictxt, "Row-major", &nprow, &npcol);
    blacs_gridinfo_(&ictxt, &nprow, &npcol, &myrow, &mycol);
    if (myrow < 0) { blacs_exit_(&zero); return 0; }
    descinit_(desca, &n, &n, &nb, &nb, &zero, &zero, &ictxt, &lld, &info);
    descinit_(descb, &n, &nrhs, &nb, &nb, &zero, &zero, &ictxt, &lld, &info);
    descinit_(descc, &n, &n, &nb2, &nb2, &zero, &zero, &ictxt, &lld, &info);
    for (int i = 0; i < N * N; ++i) { s = s * 6364136223846793005ULL + 1; a[i] = a0[i] = (double)(s >> 11) / 9007199254740992.0 - 0.5; }
    for (int i = 0; i < N * NRHS; ++i) { s = s * 6364136223846793005ULL + 1; b[i] = bb[i] = (double)(s >> 11) / 9007199254740992.0 - 0.5; }

    /* expert driver: equilibrate if needed, factor, condition estimate, solve, refine */
    pdgesvx_("E", "N", &n, &nrhs, a, &one, &one, desca, af, &one, &one, desca, ipiv, &equed, r, cs, b, &one, &one, descb, x, &one, &one, descb,
             &rcond, ferr, berr, work, &lwork, iwork, &liwork, &info);
    if (info != 0) { fprintf(stderr, "PDGESVX INFO = %d\n", info); return 2; }
    for (int k = 0; k < NRHS; ++k)
        for (int i = 0; i < N; ++i) { double t = -bb[i + k * N]; for (int j = 0; j < N; ++j) t += a0[i + j * N] * x[j + k * N]; if (fabs(t) > err) err = fabs(t); }
    printf("PDGESVX: equed %c rcond %.3e ferr %.2e berr %.2e  max |A x - b| %.2e\n", equed, rcond, ferr[0], berr[0], err);
    if (!(err < 1e-10) || !(berr[0] < 1e-14)) return 3;

    /* the same number from the pieces: PDLANGE + PDGECON on the factors PDGESVX returned */
    anorm = pdlange_("1", &n, &n, a, &one, &one, desca, work);
    { double rc2; pdgecon_("1", &n, af, &one, &one, desca, &anorm, &rc2, work, &lwork, iwork, &liwork, &info);
      if (info != 0 || fabs(rc2 - rcond) > 1e-12 * rcond) { fprintf(stderr, "PDGECON %g vs %g (info %d)\n", rc2, rcond, info); return 4; } }

    /* another block size through the reference's converter, and back */
    pdgemr2d_(&n, &n, a0, &one, &one, desca, c, &one, &one, descc, &ictxt);
    for (int i = 0; i < N * N; ++i) if (c[i] != a0[i]) { fprintf(stderr, "PDGEMR2D changed element %d on a 1 x 1 grid\n", i); return 5; }

    /* inverse from the factors */
    pdgetri_(&n, af, &one, &one, desca, ipiv, work, &lwork, iwork, &liwork, &info);
    if (info != 0) { fprintf(stderr, "PDGETRI INFO = %d\n", info); return 6; }
    err = 0.0;
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) { double t = (i == j) ? -1.0 : 0.0; for (int k = 0; k < N; ++k) t += af[i + k * N] * a[k + j * N]; if (fabs(t) > err) err = fabs(t); }
    printf("PDGETRI: max |inv(A) A - I| %.2e\n", err);
    if (!(err < 1e-9)) return 7;

    /* G = A0' A0 + I through PDGEMM, then the Cholesky driver */
    for (int i = 0; i < N * N; ++i) g[i] = 0.0;
    for (int i = 0; i < N; ++i) g[i + i * N] = 1.0;
    beta = 1.0;
    pdgemm_("T", "N", &n, &n, &n, &alpha, a0, &one, &one, desca, a0, &one, &one, desca, &beta, g, &one, &one, desca);
    for (int i = 0; i < N * N; ++i) c[i] = g[i];
    for (int i = 0; i < N * NRHS; ++i) b[i] = bb[i];
    pdposv_("L", &n, &nrhs, g, &one, &one, desca, b, &one, &one, descb, &info);
    if (info != 0) { fprintf(stderr, "PDPOSV INFO = %d\n", info); return 8; }
    err = 0.0;
    for (int k = 0; k < NRHS; ++k)
        for (int i = 0; i < N; ++i) { double t = -bb[i + k * N]; for (int j = 0; j < N; ++j) t += c[i + j * N] * b[j + k * N]; if (fabs(t) > err) err = fabs(t); }
    printf("PDPOSV: max |G x - b| %.2e\n", err);
    if (!(err < 1e-10)) return 9;
    blacs_gridexit_(&ictxt);
    printf("expert example ok\n");
    return 0;
}
