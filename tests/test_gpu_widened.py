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


@pytest.mark.parametrize("name", ["gru", "bert"])
def test_awelv_intel_matches_reference_golden(name):
    P.check_awelv_intel(DEV, name)


@pytest.mark.parametrize("name", ["h32", "h24_8"])
def test_lambdarank_scorer_matches_reference_golden(name):
    P.check_lambdarank_model(DEV, name)


def test_lambdarank_matches_reference_golden():
    P.check_lambdarank(DEV)


def test_lambdarank_properties_at_full_batch():
    """size-independent properties at the bench batch (4096 sessions x 50): pad slots are exactly 0, the lambdas of a
    session do not change when other sessions do, and flipping the sign convention (swap two items' labels and scores
    together) permutes the lambdas when the two slots share a discount - here: a session reversed twice is itself"""
    from intel_sigir2023_b200 import lambdarank, synthetic
    pred, ranking, _, slen = synthetic.eval_set(4096, 50, 25, seed=5, device=DEV)
    scores = pred.float().softmax(dim=-1)
    lam = lambdarank.compute_lambda_new(ranking, scores, slen)
    pad = torch.arange(50, device=DEV)[None, :] >= slen[:, None]
    live = ~torch.isnan(lam).any(dim=1)
    assert live.any() and not lam[live][pad[live]].any()
    half = lambdarank.compute_lambda_new(ranking[1000:3000], scores[1000:3000], slen[1000:3000])
    assert torch.equal(torch.nan_to_num(half, nan=-7.0), torch.nan_to_num(lam[1000:3000], nan=-7.0))
    # antisymmetry of the pair terms: for scores all equal, Rho = 1/2 and sum_i Lambda_i = 0 up to rounding
    flat = lambdarank.compute_lambda_new(ranking, torch.zeros_like(scores), slen)
    tot = flat[live].sum(dim=1).abs().max().item()
    assert tot < 1e-4, tot


@pytest.mark.parametrize("intent_num,kind", [(12, "list"), (2048, "list"), (12, "bpr")])
def test_config3_shapes_against_the_oracle(intent_num, kind):
    P.check_config3_shapes(DEV, intent_num, B=40, kind=kind)
