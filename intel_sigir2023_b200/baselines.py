"""Fixed-weight list-fusion baselines of script/baselines.sh on the same kernels:
SingleSort (models/unsupervise/SingleSort.py), Borda (models/unsupervise/Borda.py) and the
random-softmax fusion of GeneralSeq.forward (models/GeneralSeq.py:23-32); and the learned per-user / per-session
softmax fusions of SURVEY.md 8f-4: aWELv (models/supervise/aWELv.py), aWELv_Int (models/supervise/aWELv_Int.py)."""
from __future__ import annotations

import argparse
from typing import Dict, Optional

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


class _AWELvIntFn(torch.autograd.Function):
    """(parameters...) -> (weights [B,L,K], ens_score [B,L], intents [B,I]) for aWELv_Int: intel_intent_fwd, then the
    per-session context row [E_u[u] || intent_embeddings(intent)] (intel_gather_fwd + intel_linear_fwd) through the aWELv
    kernel with one table row per session.  Backward walks the same entry points in reverse; `intent_embeddings`,
    `uid_embeddings` are shared between the head and the intent predictor, so every *_bwd call accumulates into one
    zero-filled gradient buffer."""

    @staticmethod
    def forward(ctx, model: "aWELv_Int", batch: Dict[str, object], *params: torch.Tensor):
        from .IntEL import _take_ws, _give_ws
        lib = _lib.load()
        cfg = model.cfg
        dev = params[0].device
        tensors = dict(zip(model._param_names, params))
        scores = batch["scores"]
        B, L, K = scores.shape
        I, du, dint = cfg.intent_num, cfg.u_emb_size, cfg.intent_emb_size
        hid = du + dint
        dims = _lib.make_dims(cfg, B, L, batch["his_context_mh"].shape[1], batch["his_item_id"].shape[1])
        P = _lib.make_tensors(cfg, tensors)
        bt = _lib.make_batch(batch, cfg)
        stream = _lib.stream_ptr(dev)
        ws_int = _take_ws(model, lib.intel_intent_workspace_bytes(dims), dev)
        intents = torch.empty(B, I, dtype=torch.float32, device=dev)
        _lib.check(lib.intel_intent_fwd(dims, P, bt, _lib.ptr(intents), _lib.ptr(ws_int), ws_int.numel(), stream))
        h_ctx = torch.empty(B, hid, dtype=torch.float32, device=dev)
        h_int = torch.empty(B, dint, dtype=torch.float32, device=dev)
        _lib.check(lib.intel_gather_fwd(B, du, _lib.ptr(tensors["uid_embeddings.weight"]), _lib.ptr(batch["u_id_c"], torch.int64),
                                        _lib.ptr(h_ctx), hid, 0, stream))
        _lib.check(lib.intel_linear_fwd(B, dint, I, _lib.ptr(intents), _lib.ptr(tensors["intent_embeddings.weight"]),
                                        _lib.ptr(tensors["intent_embeddings.bias"]), _lib.ptr(h_int), stream))
        h_ctx[:, du:].copy_(h_int)
        rows = torch.arange(B, dtype=torch.int64, device=dev)        # one context row per session
        weights = torch.empty(B, L, K, dtype=torch.float32, device=dev)
        ens = torch.empty(B, L, dtype=torch.float32, device=dev)
        w_sess = torch.empty(B, K, dtype=torch.float32, device=dev)
        _lib.check(lib.intel_awelv_fwd(B, L, K, hid, _lib.ptr(h_ctx), _lib.ptr(tensors["model_embeddings.weight"]),
                                       _lib.ptr(rows), _lib.ptr(scores, torch.float64), _lib.ptr(weights), _lib.ptr(ens),
                                       _lib.ptr(w_sess), stream))
        if any(ctx.needs_input_grad):
            ctx.model, ctx.batch, ctx.dims, ctx.params, ctx.ws_int = model, batch, dims, params, ws_int
            ctx.h_ctx, ctx.rows, ctx.w_sess = h_ctx, rows, w_sess
            ctx.save_for_backward(intents)          # an output: see _IntelFn.forward for why it is not a plain attribute
        else:
            _give_ws(model, ws_int)
        return weights, ens, intents

    @staticmethod
    def backward(ctx, d_weights, d_ens, d_intents):
        from .IntEL import _give_ws
        lib = _lib.load()
        model, cfg, batch, dims, params = ctx.model, ctx.model.cfg, ctx.batch, ctx.dims, ctx.params
        (intents,) = ctx.saved_tensors
        dev = params[0].device
        names = model._param_names
        tensors = dict(zip(names, params))
        scores = batch["scores"]
        B, L, K = scores.shape
        I, du, dint = cfg.intent_num, cfg.u_emb_size, cfg.intent_emb_size
        hid = du + dint
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 63) // 64 * 64
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        grads = dict(zip(names, (flat[o:o + p.numel()].view(p.shape) for o, p in zip(offs, params))))
        model._flat_grad = flat
        P, G = _lib.make_tensors(cfg, tensors), _lib.make_tensors(cfg, grads)
        bt = _lib.make_batch(batch, cfg)
        stream = _lib.stream_ptr(dev)
        d_int_head: Optional[torch.Tensor] = None
        if d_weights is not None or d_ens is not None:
            g_ctx = torch.zeros(B, hid, dtype=torch.float32, device=dev)
            dw = d_weights.contiguous() if d_weights is not None else None
            de = d_ens.contiguous() if d_ens is not None else None
            _lib.check(lib.intel_awelv_bwd(B, L, K, hid, _lib.ptr(ctx.h_ctx), _lib.ptr(tensors["model_embeddings.weight"]),
                                           _lib.ptr(ctx.rows), _lib.ptr(scores, torch.float64), _lib.ptr(ctx.w_sess),
                                           _lib.ptr(dw), _lib.ptr(de), _lib.ptr(g_ctx),
                                           _lib.ptr(grads["model_embeddings.weight"]), stream))
            _lib.check(lib.intel_scatter_add_bwd(B, du, _lib.ptr(g_ctx), hid, _lib.ptr(batch["u_id_c"], torch.int64),
                                                 _lib.ptr(grads["uid_embeddings.weight"]), stream))
            g_hint = g_ctx[:, du:].contiguous()
            d_int_head = torch.empty(B, I, dtype=torch.float32, device=dev)
            _lib.check(lib.intel_linear_dx(B, dint, I, _lib.ptr(g_hint), _lib.ptr(tensors["intent_embeddings.weight"]),
                                           _lib.ptr(d_int_head), None, stream))
            _lib.check(lib.intel_linear_dw(B, dint, I, _lib.ptr(g_hint), _lib.ptr(intents),
                                           _lib.ptr(grads["intent_embeddings.weight"]),
                                           _lib.ptr(grads["intent_embeddings.bias"]), stream))
        first = d_intents.contiguous() if d_intents is not None else d_int_head
        extra = d_int_head if d_intents is not None else None
        if first is not None:
            _lib.check(lib.intel_intent_bwd(dims, P, bt, _lib.ptr(intents), _lib.ptr(first), _lib.ptr(extra), G,
                                            _lib.ptr(ctx.ws_int), ctx.ws_int.numel(), stream))
        _give_ws(model, ctx.ws_int)
        ctx.ws_int = ctx.h_ctx = ctx.rows = ctx.w_sess = None
        # item_embeddings is declared but never read by the reference forward: its .grad stays None there (Adam skips it)
        return (None, None) + tuple(None if n == "item_embeddings.weight" else grads[n] for n in names)


class aWELv_Int(nn.Module):
    """aWELv with the predicted intent in the per-session context (models/supervise/aWELv_Int.py; script/baselines.sh:40):
    IntEL's intent predictor (same `predict_intent`, aWELv_Int.py:66-95 = IntEL.py:126-155) feeds
    softmax_m <[E_u[u] || intent_embeddings(intent)], E_model[m]>.  Same parameter names, flags and output dict as the
    reference module; trains with the Int*loss criteria like IntEL."""
    reader, runner = "SeqReader", "BaseRunner"
    extra_log_args = ['user_emb_size', 'intent_emb_size']

    @staticmethod
    def parse_model_args(parser):
        parser.add_argument('--context_emb_size', type=int, default=16, help='Embedding size for context.')
        parser.add_argument('--user_emb_size', type=int, default=16)
        parser.add_argument('--intent_emb_size', type=int, default=16)
        parser.add_argument('--encoder', type=str, default='BERT4Rec', help='A sequence encoder for intent prediction.')
        parser.add_argument('--i_emb_size', type=int, default=16, help='Embedding size for item id.')
        parser.add_argument('--im_emb_size', type=int, default=16, help='Embedding size for item metadata.')
        parser.add_argument('--history_max', type=int, default=20)
        parser.add_argument('--model_path', type=str, default='', help='Model save path.')
        parser.add_argument('--buffer', type=int, default=1, help='Whether to buffer feed dicts for dev/test')
        parser.add_argument('--model_num', type=int, default=2, help='Number of base models.')
        return parser

    def __init__(self, args, corpus=None, cfg=None):
        super().__init__()
        from .config import IntelConfig
        from .IntEL import _encoder_container
        if cfg is None:
            ns = argparse.Namespace(**vars(args))
            ns.u_emb_size = getattr(args, "user_emb_size", 16)      # the flag is spelled differently here (aWELv_Int.py:21)
            cfg = IntelConfig.from_args(ns, corpus)
        self.cfg = c = cfg
        self.device = getattr(args, "device", torch.device("cuda"))
        self.model_path, self.buffer = getattr(args, "model_path", ""), getattr(args, "buffer", 1)
        self.optimizer, self.scheduler = None, None
        self._ws_pool = {}
        self.intent_num, self.model_num = c.intent_num, c.model_num
        self.user_num, self.item_num, self.max_his = c.user_rows, c.item_rows, c.history_max
        self.hidden_size = c.u_emb_size + c.intent_emb_size
        # registered in the reference's order (aWELv_Int.py:34-64)
        self.uid_embeddings = nn.Embedding(c.user_rows, c.u_emb_size)
        self.intent_embeddings = nn.Linear(c.intent_num, c.intent_emb_size)
        self.model_embeddings = nn.Embedding(c.model_num, self.hidden_size)
        self.iid_embeddings = nn.Embedding(c.item_rows, c.i_emb_size)
        if c.class_rows > 0:
            self.item_embeddings = nn.Embedding(c.class_rows, c.im_emb_size)     # declared, never read (aWELv_Int.py:44)
        self.context_embeddings = nn.Embedding(c.ctx_rows, c.context_emb_size)
        self.encoder = _encoder_container(c, c.d_his)
        self.item_encoder = _encoder_container(c, c.d_his_item)
        self.pred_layer = nn.Linear(c.d_pred, c.intent_num)
        self._param_names = [n for n, _ in self.named_parameters()]

    def customize_parameters(self, define_dict=None) -> list:
        from .optim import customize_parameters
        return customize_parameters(self)

    def forward(self, data: Dict[str, object]) -> Dict[str, torch.Tensor]:
        params = [p for _, p in self.named_parameters()]
        w, ens, intents = _AWELvIntFn.apply(self, data, *params)
        return {"weights": w, "ens_score": ens, "intents": intents}


class _PoolHeadFn(torch.autograd.Function):
    """(per-slot head output [B,L,K] of the cross_attention = 0 path, scores) -> (weights [B,L,K], ens_score [B,L]) through
    intel_pool_head_fwd / intel_pool_head_bwd: mean over the list slots, softmax twice, fusion."""

    @staticmethod
    def forward(ctx, slot_weights, scores):
        lib = _lib.load()
        B, L, K = scores.shape
        dev = scores.device
        weights = torch.empty(B, L, K, dtype=torch.float32, device=dev)
        ens = torch.empty(B, L, dtype=torch.float32, device=dev)
        p_sess = torch.empty(B, K, dtype=torch.float32, device=dev)
        w_sess = torch.empty(B, K, dtype=torch.float32, device=dev)
        _lib.check(lib.intel_pool_head_fwd(B, L, K, _lib.ptr(slot_weights.contiguous(), torch.float32),
                                           _lib.ptr(scores, torch.float64), _lib.ptr(weights), _lib.ptr(ens), _lib.ptr(p_sess),
                                           _lib.ptr(w_sess), _lib.stream_ptr(dev)))
        ctx.save_for_backward(scores, p_sess, w_sess)
        return weights, ens

    @staticmethod
    def backward(ctx, d_weights, d_ens):
        lib = _lib.load()
        scores, p_sess, w_sess = ctx.saved_tensors
        B, L, K = scores.shape
        d_slot = torch.empty(B, L, K, dtype=torch.float32, device=scores.device)
        dw = d_weights.contiguous() if d_weights is not None else None
        de = d_ens.contiguous() if d_ens is not None else None
        _lib.check(lib.intel_pool_head_bwd(B, L, K, _lib.ptr(scores, torch.float64), _lib.ptr(p_sess), _lib.ptr(w_sess),
                                           _lib.ptr(dw), _lib.ptr(de), _lib.ptr(d_slot), _lib.stream_ptr(scores.device)))
        return d_slot, None


def _intel_base():
    from .IntEL import IntEL
    return IntEL


class aWELv_IntEL(_intel_base()):
    """aWELv with IntEL's networks (models/supervise/aWELv_IntEL.py; script/baselines.sh:47): the intent predictor, the two
    self-attention stacks and the gate form of the intent conditioning are IntEL's `cross_attention = 0` path unchanged
    (same parameters, same registration order); the weight head then sees the unmasked mean over the list slots and its
    output goes through softmax twice (aWELv_IntEL.py:190-201).  The head is affine, so the mean of its per-slot output is
    its output on the mean: `intel_ensemble_fwd` + `intel_pool_head_fwd`, and the reverse for the gradients."""
    extra_log_args = ['cross_attn_qsize', 'num_heads', 'num_layers', 'encoder', 'intent_emb_size']

    def __init__(self, args, corpus=None, cfg=None):
        import dataclasses
        from .config import IntelConfig
        cfg = cfg if cfg is not None else IntelConfig.from_args(args, corpus)
        super().__init__(args, corpus, cfg=dataclasses.replace(cfg, cross_attention=0))

    def forward(self, data: Dict[str, object]) -> Dict[str, torch.Tensor]:
        out = super().forward(data)
        w, ens = _PoolHeadFn.apply(out["weights"], data["scores"])
        return {"weights": w, "ens_score": ens, "intents": out["intents"]}
