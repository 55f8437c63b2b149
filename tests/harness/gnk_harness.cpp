// Host build of elfi_b200/csrc/gnkmath.cuh (test infrastructure, see tests/test_gnkmath_host.py).
#include <cstdint>

#include "../../elfi_b200/csrc/gnkmath.cuh"

extern "C" void harness_gnk_quantile(const double* prm, double c, const double* z, int64_t n,
                                     double* out) {
    for (int64_t i = 0; i < n; ++i)
        out[i] = elfi::gnk_quantile(prm[0], prm[1], prm[2], prm[3], c, z[i]);
}
