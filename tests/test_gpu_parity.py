"""Parity gate on the B200: host layer -> C ABI -> sm_100a kernels against the golden vectors of the
unmodified reference and the CPU oracle (bit-exact for index work, 1e-5 relative for fp32, NDCG to 1e-6)."""
import pytest
import torch

import parity_checks as P
from conftest import MODEL_CASES

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


def test_cpu_tensors_are_rejected():
    from intel_sigir2023_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.ptr(torch.zeros(3))


def test_linear():
    P.check_linear(DEV)


def test_gather_scatter():
    P.check_gather_scatter(DEV)


@pytest.mark.parametrize("name", MODEL_CASES)
def test_forward(name):
    P.check_forward(name, DEV)


@pytest.mark.parametrize("name", MODEL_CASES)
@pytest.mark.parametrize("kind", ["list", "bpr", "mse"])
def test_backward(name, kind):
    P.check_backward(name, kind, DEV)


def test_loss_edge_cases():
    P.check_loss_edge_cases(DEV)


@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_evaluate(tag):
    P.check_evaluate(tag, DEV)


def test_evaluate_ties():
    P.check_evaluate_ties(DEV)


def test_evaluate_intents():
    P.check_evaluate_intents(DEV)


def test_baselines():
    P.check_baselines(DEV)
