"""Output pools on the device (row N4): stored summaries / discrepancies stay in HBM across
inferences and spill to .npy lazily.  Same bodies as tests/test_store_cpu_double.py."""
import pytest

import store_cases as cases

pytestmark = pytest.mark.gpu


def test_pool_usage_device_resident():
    cases.case_pool_usage()


def test_array_pool_write_back_cache(tmp_path):
    cases.case_array_pool(tmp_path)


def test_pool_restarts(tmp_path):
    cases.case_pool_restarts(tmp_path)
