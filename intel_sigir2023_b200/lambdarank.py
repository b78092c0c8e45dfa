"""LambdaRank's per-item lambdas on the B200 path: `compute_lambda_new` of
helpers/LambdaRankRunner.py:315-344, the O(L^2) pairwise pass of every LambdaRank training step (:246), as one kernel
(`intel_lambdarank_lambdas`, one warp per session) instead of ten [B,L,L] temporaries."""
from __future__ import annotations

import torch

from . import _lib


def compute_lambda_new(true_scores: torch.Tensor, temp_scores: torch.Tensor, session_len: torch.Tensor) -> torch.Tensor:
    """Same arguments as the reference method (without `self`): `true_scores` int64 [B,L] = clamp(batch['ranking'], min=0)
    (the raw ranking works too, the kernel clamps), `temp_scores` float32 [B,L] = the detached `ens_score`,
    `session_len` int64 [B] -> Lambda float32 [B,L], to be fed to `predicted_scores.backward(Lambda)` (:259)."""
    lib = _lib.load()
    B, L = true_scores.shape
    ranking = true_scores.contiguous()
    scores = temp_scores.detach().contiguous()
    lens = session_len.contiguous()
    lambdas = torch.empty(B, L, dtype=torch.float32, device=scores.device)
    _lib.check(lib.intel_lambdarank_lambdas(B, L, _lib.ptr(ranking, torch.int64), _lib.ptr(scores, torch.float32),
                                            _lib.ptr(lens, torch.int64), _lib.ptr(lambdas), _lib.stream_ptr(scores.device)))
    return lambdas
