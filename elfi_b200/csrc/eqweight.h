// eqweight.h -- exact position of a quantile in NumPy's sequential cumulative sum of n equal
// weights, without the n sequential additions.
//
// weighted_sample_quantile (elfi/methods/utils.py:379-411) with weights=None builds
//   w = ones(n) / sum(ones(n))           (every weight c = fl(1/n))
//   cum = [0, cumsum(w)...], cum[-1] = 1 (np.cumsum adds left to right, one rounding per step)
// and returns the sorted sample at the first k with cum[k] < alpha <= cum[k+1].  With equal
// weights and a round alpha (SMC round 0: quantile 0.5 of 1e6 particles) alpha sits exactly on a
// cumulative weight up to rounding, so the rounding of every one of the n additions matters.
//
// The sum s_j = fl(s_{j-1} + c) is followed exactly but in O(number of binades) steps: while s
// stays inside one binade [2^e, 2^(e+1)) its ulp u is constant, s is a multiple of u, and after
// the first addition in the binade every step adds the same multiple of u (round-to-nearest-even
// of the constant c / u; in the tie case the sum is even after one step and stays even).  So the
// progression is arithmetic per binade and is advanced with integer arithmetic in units of u; the
// few additions around a binade boundary are done for real.
//
// Host-only, no CUDA: included by select.cu and by tests/harness/eqweight_harness.cpp, which
// checks it against np.cumsum for every small n and many large ones.
#pragma once

#include <cmath>
#include <cstdint>

namespace elfi {

// Smallest j in [1, n] with cum[j] >= alpha (0 < alpha <= 1), cum as defined above.
// The quantile is element j - 1 of the ascending sample.
inline int64_t equal_weight_cum_index(int64_t n, double alpha) {
    if (n <= 1) return 1;
    const double c = 1.0 / double(n);
    volatile double sv;          // volatile: every addition is rounded to fp64, never contracted
    double s = c;                // s_1
    int64_t j = 1;
    if (s >= alpha) return 1;
    double d_prev = -1.0;
    int e_prev = 0;
    while (j < n - 1) {
        sv = s + c;
        const double s1 = sv;    // s_{j+1}
        ++j;
        if (s1 >= alpha) return j;
        int e0, e1;
        std::frexp(s, &e0);
        std::frexp(s1, &e1);
        const double d = s1 - s;                  // exact (Sterbenz)
        const bool same_binade = (e0 == e1);
        s = s1;
        if (!(same_binade && e_prev == e1 && d == d_prev)) {
            d_prev = same_binade ? d : -1.0;
            e_prev = e1;
            continue;
        }
        // two consecutive equal increments inside the binade 2^(e1-1) <= s < 2^e1: the
        // progression is arithmetic until it leaves the binade.  Work in units of u = 2^(e1-53).
        const int sh = 53 - e1;
        const int64_t S = int64_t(std::ldexp(s, sh));          // exact integers < 2^53
        const int64_t D = int64_t(std::ldexp(d, sh));
        const int64_t E = int64_t(1) << 53;                    // binade end in units of u
        int64_t t = (E - 1 - S) / D;                           // steps that stay below the end
        if (t > n - 1 - j) t = n - 1 - j;
        if (alpha < std::ldexp(1.0, e1)) {                     // alpha inside this binade: exact
            const int64_t A = int64_t(std::ldexp(alpha, sh));  // multiple of u (alpha > s >= 2^(e1-1))
            const int64_t t_cross = (A - S + D - 1) / D;       // first step reaching alpha
            if (t_cross <= t) return j + t_cross;
        }
        if (t > 0) {
            s = std::ldexp(double(S + t * D), -sh);
            j += t;
        }
        d_prev = -1.0;                                         // re-establish after the boundary
    }
    return n;   // cum[n] is forced to 1.0 >= alpha
}

}  // namespace elfi
