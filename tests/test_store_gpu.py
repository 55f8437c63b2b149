"""Output pool on the device (row N4): stored summaries / discrepancies stay in HBM across
inferences and spill to host memory lazily.  Same bodies as tests/test_store_cpu_double.py."""
import pytest

import store_cases as cases

pytestmark = pytest.mark.gpu


def test_pool_usage_device_resident():
    cases.case_pool_usage()


def test_pool_spill_oldest_batches_first():
    cases.case_pool_spill()
