"""Parity on the B200 for the rows widened into after the hot path: other model heads on the same
kernels.  Kept in its own file, after test_gpu_parity / test_gpu_properties, so the hot-path gate runs first."""
import pytest
import torch

import parity_checks as P

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def real_lib():
    from intel_sigir2023_b200 import _lib
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    lib = _lib.load()
    assert not _lib._allow_host_tensors
    assert lib._name.endswith("libintel_b200.so")
    yield


@pytest.mark.parametrize("name", ["gru", "bert"])
def test_awelv_int_matches_reference_golden(name):
    P.check_awelv_int(DEV, name)
