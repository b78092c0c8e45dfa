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


class _AWELvFn(torch.autograd.Function):
    """(user table, model table) -> (weights [B,L,K], ens_score [B,L]) through intel_awelv_fwd / intel_awelv_bwd."""

    @staticmethod
    def forward(ctx, user_table, model_table, u_id, scores):
        lib = _lib.load()
        B, L, K = scores.shape
        h = user_table.shape[1]
        dev = scores.device
        weights = torch.empty(B, L, K, dtype=torch.float32, device=dev)
        ens = torch.empty(B, L, dtype=torch.float32, device=dev)
        w_user = torch.empty(B, K, dtype=torch.float32, device=dev)
        _lib.check(lib.intel_awelv_fwd(B, L, K, h, _lib.ptr(user_table, torch.float32), _lib.ptr(model_table, torch.float32),
                                       _lib.ptr(u_id, torch.int64), _lib.ptr(scores, torch.float64), _lib.ptr(weights),
                                       _lib.ptr(ens), _lib.ptr(w_user), _lib.stream_ptr(dev)))
        ctx.save_for_backward(user_table, model_table, u_id, scores, w_user)
        return weights, ens

    @staticmethod
    def backward(ctx, d_weights, d_ens):
        lib = _lib.load()
        user_table, model_table, u_id, scores, w_user = ctx.saved_tensors
        B, L, K = scores.shape
        g_user, g_model = torch.zeros_like(user_table), torch.zeros_like(model_table)
        dw = d_weights.contiguous() if d_weights is not None else None
        de = d_ens.contiguous() if d_ens is not None else None
        _lib.check(lib.intel_awelv_bwd(B, L, K, user_table.shape[1], _lib.ptr(user_table), _lib.ptr(model_table),
                                       _lib.ptr(u_id), _lib.ptr(scores), _lib.ptr(w_user), _lib.ptr(dw), _lib.ptr(de),
                                       _lib.ptr(g_user), _lib.ptr(g_model), _lib.stream_ptr(scores.device)))
        return g_user, g_model, None, None


class aWELv(nn.Module):
    """Per-user softmax fusion weights (models/supervise/aWELv.py; script/baselines.sh:33): same parameter names
    (`uid_embeddings.weight`, `model_embeddings.weight`), flags (`--hidden_size`, `--model_num`) and output dict."""
    reader, runner = "BaseReader", "BaseRunner"
    extra_log_args: list = []

    @staticmethod
    def parse_model_args(parser):
        parser.add_argument('--hidden_size', type=int, default=32)
        parser.add_argument('--model_path', type=str, default='', help='Model save path.')
        parser.add_argument('--buffer', type=int, default=1, help='Whether to buffer feed dicts for dev/test')
        parser.add_argument('--model_num', type=int, default=2, help='Number of base models.')
        return parser

    def __init__(self, args, corpus=None, user_num: int = None):
        super().__init__()
        self.device = getattr(args, "device", None)
        self.user_num = int(user_num if user_num is not None else corpus.max_uid + 1)
        self.uid_embeddings = nn.Embedding(self.user_num, args.hidden_size)
        self.model_embeddings = nn.Embedding(args.model_num, args.hidden_size)

    def customize_parameters(self, define_dict=None) -> list:
        from .optim import customize_parameters
        return customize_parameters(self)

    def forward(self, data: Dict[str, object]) -> Dict[str, torch.Tensor]:
        w, ens = _AWELvFn.apply(self.uid_embeddings.weight, self.model_embeddings.weight, data['u_id_c'], data['scores'])
        return {"weights": w, "ens_score": ens}
