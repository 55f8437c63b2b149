// Host build of elfi_b200/csrc/leafsum.cuh: feeds rows to the single-leaf summary structs the
// way the row-stream kernel does (16 columns per call, in order, one or two sweeps) so that the
// arithmetic and control flow of the device consumers can be checked against NumPy on a CPU.
// Columns at or beyond the row length are filled with NaN: a term that touched them would
// poison the result.  Test infrastructure only (tests/test_leafsum_host.py builds it with g++).
#include <cmath>
#include <cstdint>
#include <limits>

#include "../../elfi_b200/csrc/leafsum.cuh"
#include "../../elfi_b200/csrc/treesum.cuh"

namespace {

void load_box(const double* x, int n, int t0, double* cur) {
    for (int c = 0; c < elfi::LEAF_BOX; ++c)
        cur[c] = (t0 + c < n) ? x[t0 + c] : std::numeric_limits<double>::quiet_NaN();
}

typedef elfi::TreeSum<6> Tree;   // the depth the row-stream kernels use (rows up to 8192 terms)

template <class Sum, int LA, int LB>
void autocov_rows(const double* X, int64_t ld, int64_t B, int n, double* out) {
    for (int64_t b = 0; b < B; ++b) {
        elfi::AutocovBoxes<Sum, LA, LB> st;
        st.begin(n);
        double cur[elfi::LEAF_BOX];
        for (int t0 = 0; t0 < n; t0 += elfi::LEAF_BOX) {
            load_box(X + b * ld, n, t0, cur);
            st.box(t0, cur);
        }
        out[2 * b] = st.sum_a() / double(n - LA);
        out[2 * b + 1] = (LB >= 0) ? st.sum_b() / double(n - LB) : 0.0;
    }
}

template <class Sum>
int autocov_dispatch(const double* X, int64_t ld, int64_t B, int n, int lag_a, int lag_b,
                     double* out) {
    if (lag_a == 1 && lag_b == 2) autocov_rows<Sum, 1, 2>(X, ld, B, n, out);
    else if (lag_a == 1 && lag_b < 0) autocov_rows<Sum, 1, -1>(X, ld, B, n, out);
    else if (lag_a == 2 && lag_b < 0) autocov_rows<Sum, 2, -1>(X, ld, B, n, out);
    else if (lag_a == 3 && lag_b < 0) autocov_rows<Sum, 3, -1>(X, ld, B, n, out);
    else if (lag_a == 4 && lag_b < 0) autocov_rows<Sum, 4, -1>(X, ld, B, n, out);
    else return -1;
    return 0;
}

template <class Sum>
int meanvar_rows(const double* X, int64_t ld, int64_t B, int n, double* out) {
    for (int64_t b = 0; b < B; ++b) {
        elfi::MeanVarBoxes<Sum> st;
        st.begin(n);
        double cur[elfi::LEAF_BOX];
        for (int pass = 0; pass < 2; ++pass)
            for (int t0 = 0; t0 < n; t0 += elfi::LEAF_BOX) {
                load_box(X + b * ld, n, t0, cur);
                st.box(pass, t0, cur);
            }
        out[2 * b] = st.mean;
        out[2 * b + 1] = st.variance();
    }
    return 0;
}

template <int NBOX>
void meanvar_regs_rows(const double* X, int64_t ld, int64_t B, int n, double* out) {
    for (int64_t b = 0; b < B; ++b) {
        elfi::MeanVarRegs<NBOX> st;
        st.begin(n);
        double cur[elfi::LEAF_BOX];
        for (int g = 0; g * elfi::LEAF_BOX < n; ++g) {
            load_box(X + b * ld, n, g * elfi::LEAF_BOX, cur);
            switch (g) {                      // the kernel dispatches on the box index like this
                case 0: st.template box<0>(cur); break;
                case 1: if constexpr (NBOX > 1) st.template box<1>(cur); break;
                case 2: if constexpr (NBOX > 2) st.template box<2>(cur); break;
                default: if constexpr (NBOX > 3) st.template box<3>(cur); break;
            }
        }
        st.finish(out[2 * b], out[2 * b + 1]);
    }
}

}  // namespace

extern "C" {

// single-sweep mean / variance with the row held in registers (rows of <= 64 observations)
int harness_meanvar_regs(const double* X, int64_t ld, int64_t B, int n, double* out) {
    if (n <= 16) meanvar_regs_rows<1>(X, ld, B, n, out);
    else if (n <= 32) meanvar_regs_rows<2>(X, ld, B, n, out);
    else if (n <= 48) meanvar_regs_rows<3>(X, ld, B, n, out);
    else if (n <= 64) meanvar_regs_rows<4>(X, ld, B, n, out);
    else return -1;
    return 0;
}

// tree = 0: single-leaf accumulator (rows of <= 128 terms); tree = 1: TreeSum (up to 8192 terms)
int harness_autocov(const double* X, int64_t ld, int64_t B, int n, int lag_a, int lag_b,
                    double* out) {
    return autocov_dispatch<elfi::LeafSum>(X, ld, B, n, lag_a, lag_b, out);
}
int harness_autocov_tree(const double* X, int64_t ld, int64_t B, int n, int lag_a, int lag_b,
                         double* out) {
    return autocov_dispatch<Tree>(X, ld, B, n, lag_a, lag_b, out);
}
int harness_meanvar(const double* X, int64_t ld, int64_t B, int n, double* out) {
    return meanvar_rows<elfi::LeafSum>(X, ld, B, n, out);
}
int harness_meanvar_tree(const double* X, int64_t ld, int64_t B, int n, double* out) {
    return meanvar_rows<Tree>(X, ld, B, n, out);
}

}  // extern "C"
