// Host check of omg_planner_b200/csrc/learner_bisect.h: the two-steps-per-round bisection visits the points of the
// sequential loop (find_zero, omg/online_learner.py:18-30) and returns the same value, bit for bit -- on the Bregman
// projection's own function (sum_g sh_g exp(x + alpha_g - v_g) - target, the kernel's expression), on roots outside
// the bracket (the loop runs out of its 100 steps), on early exits at either step of a round, and on a function that
// returns NaN over part of the bracket (the point stops moving).  Prints "trials N mismatches M early_exit_second E exhausted X nan_cases Y".
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../omg_planner_b200/csrc/learner_bisect.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double urand() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) / 9007199254740992.0;
}

struct Problem {
    int G;
    double sh[64], a_minus_v[64], target;
    double nan_below;   // f is NaN for x below this (a pure function of x, like the kernel's)
    double f(double x) const {
        if (x < nan_below) return NAN;
        double part = 0.0;
        for (int g = 0; g < G; ++g) part += sh[g] * exp(x + a_minus_v[g]);
        return part - target;
    }
};

int main(int argc, char **argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 20000;
    int mismatches = 0, early_second = 0, exhausted = 0, nan_cases = 0;
    for (int t = 0; t < trials; ++t) {
        Problem P;
        P.G = 2 + (int)(urand() * 30);
        const double delta = 1.0 / (4 * P.G + 1);
        double s = 0.0, vmax = -1e300;
        for (int g = 0; g < P.G; ++g) { P.sh[g] = urand() + 1e-3; s += P.sh[g]; }
        for (int g = 0; g < P.G; ++g) {
            P.sh[g] = P.sh[g] / s + delta;
            const double v = urand() * (t % 3 == 0 ? 8.0 : 0.5), alpha = (t % 5 == 0) ? urand() * 0.3 : 0.0;
            P.a_minus_v[g] = alpha - v;
            if (1.0 + v > vmax) vmax = 1.0 + v;
        }
        P.target = 1.0 + P.G * delta;
        if (t % 11 == 0) P.target *= 1e6;          // root beyond the bracket: all 100 steps run
        P.nan_below = (t % 13 == 0) ? vmax * urand() : -1e300;
        double err = 1e-6;
        if (t % 7 == 0) err = 1e-2;                // early exits at shallow depth, at either step of a round
        const double x1 = vmax;
        // sequential
        int seq_evals = 0, seq_nan = 0;
        const double a = omgb::lrn_find_zero_seq(x1, err, [&](double x) {
            const double f = P.f(x);
            ++seq_evals;
            if (f != f) ++seq_nan;
            return f;
        });
        // two steps per round
        const double b = omgb::lrn_find_zero_two(x1, err, [&](double x, double st, double &f0, double &fm, double &fp) {
            f0 = P.f(x);
            fm = P.f(x - st);
            fp = P.f(x + st);
        });
        if (memcmp(&a, &b, sizeof(double)) != 0) {
            ++mismatches;
            if (mismatches < 5) fprintf(stderr, "trial %d: seq %.17g two %.17g (seq evals %d)\n", t, a, b, seq_evals);
        }
        if (seq_evals >= 100) ++exhausted;
        if (seq_nan) ++nan_cases;
        if (seq_evals % 2 == 0 && seq_evals < 100) ++early_second;
    }
    printf("trials %d mismatches %d early_exit_second %d exhausted %d nan_cases %d\n", trials, mismatches, early_second,
           exhausted, nan_cases);
    return mismatches ? 1 : 0;
}
