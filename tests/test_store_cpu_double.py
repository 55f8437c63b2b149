"""The device-resident output pool (elfi_b200/store.py) on the CPU
test double."""
import store_cases as cases


def test_pool_usage(cpu_double):
    cases.case_pool_usage()


def test_pool_spill(cpu_double):
    cases.case_pool_spill()
