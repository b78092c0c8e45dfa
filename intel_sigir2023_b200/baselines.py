"""Fixed-weight list-fusion baselines of script/baselines.sh on the same kernels:
SingleSort (models/unsupervise/SingleSort.py), Borda (models/unsupervise/Borda.py) and the
random-softmax fusion of GeneralSeq.forward (models/GeneralSeq.py:23-32)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from . import _lib


def fuse(weights: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
    """ens[b,l] = sum_k weights[b,l,k] * float(scores[b,l,k])"""
    lib = _lib.load()
    B, L, K = scores.shape
    ens = torch.empty(B, L, dtype=torch.float32, device=scores.device)
    _lib.check(lib.intel_fuse_fwd(B, L, K, _lib.ptr(weights.contiguous(), torch.float32),
                                  _lib.ptr(scores, torch.float64), _lib.ptr(ens), _lib.stream_ptr(scores.device)))
    return ens


class SingleSort(nn.Module):
    reader, runner = "BaseReader", "BaseRunner"
    _COLUMN = {"pCTR": 0, "pCVR": 1}

    def __init__(self, args=None, corpus=None, choose_list: str = "pCTR"):
        super().__init__()
        self.choose_list = getattr(args, "choose_list", choose_list)

    def forward(self, data: Dict[str, object]) -> Dict[str, torch.Tensor]:
        lib = _lib.load()
        x = data['scores']
        B, L, K = x.shape
        col = self._COLUMN.get(self.choose_list, 2)
        ens = torch.empty(B, L, dtype=torch.float32, device=x.device)
        _lib.check(lib.intel_select_list(B, L, K, _lib.ptr(x, torch.float64), col, _lib.ptr(ens),
                                         _lib.stream_ptr(x.device)))
        return {"weights": torch.zeros(B, L, K, dtype=torch.float32, device=x.device), "ens_score": ens}


class Borda(nn.Module):
    reader, runner = "SeqReader", "BaseRunner"

    def __init__(self, args=None, corpus=None):
        super().__init__()

    def forward(self, data: Dict[str, object]) -> Dict[str, torch.Tensor]:
        lib = _lib.load()
        x = data['scores']
        B, L, K = x.shape
        ens = torch.empty(B, L, dtype=torch.float32, device=x.device)
        _lib.check(lib.intel_rank_lists(B, L, K, _lib.ptr(x, torch.float64), _lib.ptr(ens), _lib.stream_ptr(x.device)))
        w = torch.full((B, L, K), 1.0 / K, dtype=torch.float32, device=x.device)
        return {"weights": w, "ens_score": ens}


class RandomFusion(nn.Module):
    """GeneralSeq.forward: softmax of uniform noise as weights (the draw uses torch's generator)."""
    reader, runner = "SeqReader", "BaseRunner"

    def __init__(self, args=None, corpus=None):
        super().__init__()

    def forward(self, data: Dict[str, object], raw_weights: torch.Tensor = None) -> Dict[str, torch.Tensor]:
        x = data['scores']
        if raw_weights is None:
            raw_weights = torch.rand(x.shape, device=x.device)
        w = torch.softmax(raw_weights.float(), dim=2)
        return {"weights": w, "ens_score": fuse(w, x)}
