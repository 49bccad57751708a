// The bisection of the online learner's Bregman projection (find_zero, omg/online_learner.py:18-30), as control flow
// over an evaluation functor, so that the same code runs in the kernel (learner_kernels.cuh: the functor is a warp-wide
// exp + shuffle reduction) and on the host (tests/host/bisect_check.cpp).
//
//   lrn_find_zero_seq   one evaluation per step: the reference's loop as written.
//   lrn_find_zero_two   TWO steps per round: f is evaluated at the current point and, ahead of the decision, at both
//                       points the step can go to (xx - step, xx + step); the three evaluations are independent, so
//                       their dependent exp / reduction chains overlap.  Every point visited and every comparison is
//                       the sequential loop's, in the same order -> the same result, bit for bit.  Used when few
//                       trajectories leave the SMs idle (a one-trajectory plan spends more time in the learner's
//                       bisection than in the CHOMP step).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define OMGB_HD __host__ __device__ __forceinline__
#else
#define OMGB_HD inline
#endif

namespace omgb {

// eval1(xx) -> f(xx)
template <class Eval1>
OMGB_HD double lrn_find_zero_seq(double x1, double err, Eval1 eval1) {
    double xx = (0.0 + x1) / 2, step = (x1 - 0.0) / 4;
    for (int k2 = 0; k2 < 100; ++k2) {
        const double f = eval1(xx);
        if (fabs(f) < err) break;
        const double sg = (f > 0.0) ? 1.0 : ((f < 0.0) ? -1.0 : 0.0);
        xx -= step * sg;
        step /= 2;
    }
    return xx;
}

// eval3(xx, step, f0, fm, fp): f0 = f(xx), fm = f(xx - step), fp = f(xx + step)
template <class Eval3>
OMGB_HD double lrn_find_zero_two(double x1, double err, Eval3 eval3) {
    double xx = (0.0 + x1) / 2, step = (x1 - 0.0) / 4;
    int k2 = 0;
    while (k2 < 100) {
        double f0, fm, fp;
        eval3(xx, step, f0, fm, fp);
        if (fabs(f0) < err) break;                                   // step k2
        const double sg = (f0 > 0.0) ? 1.0 : ((f0 < 0.0) ? -1.0 : 0.0);
        xx -= step * sg;                                             // (step * +-1 is exact: xx - step or xx + step)
        step /= 2;
        if (++k2 >= 100) break;
        if (sg == 0.0) continue;                                     // NaN: the point did not move; evaluate again
        const double f1 = (sg > 0.0) ? fm : fp;                      // step k2 + 1, already evaluated
        if (fabs(f1) < err) break;
        const double sg1 = (f1 > 0.0) ? 1.0 : ((f1 < 0.0) ? -1.0 : 0.0);
        xx -= step * sg1;
        step /= 2;
        ++k2;
    }
    return xx;
}

}  // namespace omgb
