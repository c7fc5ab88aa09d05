// lacon.h -- Hager / Higham 1-norm estimator, restated from SRC/pdlacon.f (the distributed DLACON: reverse communication, ITMAX = 5,
// final alternating-sign safeguard).  Host-only: PDGECON drives it on the replicated N-vector and the device does the solves.
//
//   est = || B ||_1 for a linear map B given only through  x <- B x  (kase 1)  and  x <- B^T x  (kase 2).
//
// What the reference's source actually returns: pdlacon.f:188-189 begins EVERY call with  EST = ZERO ; ESTWORK( 1 ) = EST , so the
// estimate found by the iteration never survives to the next call and the value compared at label 140 is zero: PDLACON returns the
// alternating-sign estimate 2 ||B x_alt||_1 / (3 N) whatever the iteration found (still a lower bound of ||B||_1, but 45x - 1500x
// smaller than Higham's on PDMATGEN matrices of order 50 - 3000).  Executing the reference's text shows it (tests/fortran_refine_runner.py).  Default here = the reference's
// result, obtained with the ONE application of B that determines it (the iteration's applications cannot change the value, so they are
// not made); option lacon_keep_estimate = 1 runs the full estimator with EST carried between the stages, as LAPACK's DLACON does.
#pragma once
#include "common.h"
#include <cmath>
#include <functional>
#include <vector>

namespace slb {

// apply(x, kase): overwrite x (n doubles) with B x (kase == 1) or B^T x (kase == 2)
inline double lacon_estimate(int n, const std::function<void(double *, int)> &apply, int *napplies = nullptr)
{
    const int ITMAX = 5;                                                  // pdlacon.f: PARAMETER ( ITMAX = 5 )
    std::vector<double> x((size_t)n), v((size_t)n);
    std::vector<int> isgn((size_t)n);
    int count = 0;
    auto sgn = [](double t) { return std::signbit(t) ? -1.0 : 1.0; };     // Fortran SIGN( ONE, t )
    auto asum = [&](const std::vector<double> &t) { double s = 0; for (double e : t) s += std::fabs(e); return s; };
    auto iamax = [&](const std::vector<double> &t) { int j = 0; for (int i = 1; i < n; ++i) if (std::fabs(t[i]) > std::fabs(t[j])) j = i; return j; };
    auto B = [&](int kase) { apply(x.data(), kase); ++count; };
    double est = 0.0;
    if (n <= 0) return 0.0;
    if (n > 1 && opt("lacon_keep_estimate", 0) == 0) {                    // the reference's result (see above): labels 120 - 150 only
        for (int k = 1; k <= n; ++k) x[k - 1] = ((k % 2 == 0) ? -1.0 : 1.0) * (1.0 + (double)(k - 1) / (double)(n - 1));
        B(1);
        const double temp = 2.0 * (asum(x) / (double)(3 * n));
        if (napplies) *napplies = count;
        return temp > 0.0 ? temp : 0.0;                                   // IF( TEMP( 1 ).GT.ESTWORK( 1 ) ) with ESTWORK( 1 ) = 0
    }
    for (int i = 0; i < n; ++i) x[i] = 1.0 / (double)n;                   // pdlacon.f:10 (KASE = 0 entry)
    B(1);
    if (n == 1) {                                                         // label 20, N = 1
        if (napplies) *napplies = count;
        return std::fabs(x[0]);
    }
    est = asum(x);                                                        // label 20
    for (int i = 0; i < n; ++i) { x[i] = sgn(x[i]); isgn[i] = (int)x[i]; }
    B(2);
    int j = iamax(x), iter = 2;                                           // label 40
    bool altsgn_stage = false;
    for (;;) {
        for (int i = 0; i < n; ++i) x[i] = 0.0;                           // label 50: x = e_j
        x[j] = 1.0;
        B(1);
        v = x;                                                            // label 70
        const double estold = est;
        est = asum(v);
        bool changed = false;
        for (int i = 0; i < n; ++i) if ((int)sgn(x[i]) != isgn[i]) { changed = true; break; }
        if (!changed || est <= estold) { altsgn_stage = true; break; }    // GO TO 120
        for (int i = 0; i < n; ++i) { x[i] = sgn(x[i]); isgn[i] = (int)x[i]; }
        B(2);
        const int jlast = j;                                              // label 110
        j = iamax(x);
        if (x[jlast] != std::fabs(x[j]) && iter < ITMAX) { ++iter; continue; }
        altsgn_stage = true;
        break;
    }
    if (altsgn_stage) {                                                   // label 120: x_k = (-1)^(k+1) (1 + (k-1)/(n-1))
        for (int k = 1; k <= n; ++k) x[k - 1] = ((k % 2 == 0) ? -1.0 : 1.0) * (1.0 + (double)(k - 1) / (double)(n - 1));
        B(1);
        const double temp = 2.0 * (asum(x) / (double)(3 * n));            // label 140
        if (temp > est) est = temp;
    }
    if (napplies) *napplies = count;
    return est;
}

}  // namespace slb
