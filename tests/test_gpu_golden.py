"""The reference's own known-answer tests, run against the CUDA path through the C ABI."""
import pytest

import kat_runner

pytestmark = pytest.mark.gpu
CASES = kat_runner.load_cases()


@pytest.fixture(scope="module")
def backend(gpu_ctx):
    from backends import GpuBackend
    return GpuBackend(gpu_ctx)


@pytest.mark.parametrize("case", CASES, ids=[kat_runner.case_id(c) for c in CASES])
def test_reference_kat_on_gpu(case, backend):
    assert kat_runner.run_case(case, backend, allow_heavy=False) in ("ok", "skipped")
