// Host build of elfi_b200/csrc/eqweight.h (test infrastructure, tests/test_eqweight_host.py).
#include "../../elfi_b200/csrc/eqweight.h"

extern "C" int64_t harness_equal_weight_cum_index(int64_t n, double alpha) {
    return elfi::equal_weight_cum_index(n, alpha);
}
