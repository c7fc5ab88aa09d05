/* pdgesv_example.c -- a compiled caller of the drop-in boundary: the flow of the reference's EXAMPLE/pdscaex.f (grid set-up,
 * DESCINIT, PDGESV on the 6 x 6 tutorial system, residual check) written against include/scalapack_b200.h only.
 *
 *   gcc examples/pdgesv_example.c -Iinclude -Lscalapack_b200/lib -lscalapack_b200 -Wl,-rpath,$PWD/scalapack_b200/lib -lm -o pdgesv_example
 *   ./pdgesv_example                     (one process = a 1 x 1 grid; needs a B200: the library has no CPU fallback)
 *
 * The matrix and right-hand side are the reference's own fixture (EXAMPLE/DSCAEXMAT.dat / DSCAEXRHS.dat, column-major). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "scalapack_b200.h"

static const double AMAT[36] = {      /* column by column */
    6.0, 3.0, 0.0, 0.0, 3.0, 0.0,
    0.0, -3.0, -1.0, 1.0, 1.0, 0.0,
    -1.0, 0.0, 11.0, 0.0, 0.0, 10.0,
    0.0, 0.0, 0.0, -11.0, 0.0, 0.0,
    0.0, 0.0, 0.0, 2.0, -4.0, 0.0,
    0.0, 0.0, 0.0, 8.0, 0.0, -10.0 };
static const double BRHS[6] = { 72.0, 0.0, 160.0, 0.0, 0.0, 0.0 };

int main(void)
{
    int me, np, ictxt, sys0 = 0, what = 0, minus1 = -1, one = 1;
    int nprow = 1, npcol = 1, myrow, mycol, info, n = 6, nrhs = 1, nb = 2, lld = 6, zero = 0;
    int desca[9], descb[9], ipiv[6 + 2];
    double a[36], b[6], resid = 0.0, anorm = 0.0, xnorm = 0.0;
    blacs_pinfo_(&me, &np);
    blacs_get_(&minus1, &what, &sys0);
    ictxt = sys0;
    blacs_gridinit_(&ictxt, "Row-major", &nprow, &npcol);
    blacs_gridinfo_(&ictxt, &nprow, &npcol, &myrow, &mycol);
    if (myrow < 0) { blacs_exit_(&zero); return 0; }                 /* not in the grid (more processes than 1 x 1) */
    descinit_(desca, &n, &n, &nb, &nb, &zero, &zero, &ictxt, &lld, &info);
    descinit_(descb, &n, &nrhs, &nb, &nb, &zero, &zero, &ictxt, &lld, &info);
    for (int i = 0; i < 36; ++i) a[i] = AMAT[i];
    for (int i = 0; i < 6; ++i) b[i] = BRHS[i];
    pdgesv_(&n, &nrhs, a, &one, &one, desca, ipiv, b, &one, &one, descb, &info);
    if (info != 0) { fprintf(stderr, "PDGESV INFO = %d\n", info); return 2; }
    /* ||A x - b||_inf / (||A||_inf ||x||_inf N eps), the check of EXAMPLE/pdscaex.f:181-192 (accepts < 10) */
    for (int i = 0; i < 6; ++i) {
        double r = -BRHS[i], rs = 0.0;
        for (int j = 0; j < 6; ++j) { r += AMAT[i + 6 * j] * b[j]; rs += fabs(AMAT[i + 6 * j]); }
        if (fabs(r) > resid) resid = fabs(r);
        if (rs > anorm) anorm = rs;
        if (fabs(b[i]) > xnorm) xnorm = fabs(b[i]);
    }
    resid /= anorm * xnorm * 6.0 * ldexp(1.0, -53);
    printf("x = %.6f %.6f %.6f %.6f %.6f %.6f  ipiv = %d %d %d %d %d %d  scaled residual = %.3f\n", b[0], b[1], b[2], b[3], b[4], b[5],
           ipiv[0], ipiv[1], ipiv[2], ipiv[3], ipiv[4], ipiv[5], resid);
    blacs_gridexit_(&ictxt);
    blacs_exit_(&zero);
    return resid < 10.0 ? 0 : 1;
}
