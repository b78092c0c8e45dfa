"""Drop-in loss classes (reference IntEL/src/loss/*.py): ``criterion(out_dict, batch) -> (loss,
ensemble_loss, intent_loss)`` with ``loss`` differentiable.  Forward and gradient are ONE fused kernel
per loss (libintel_b200 ``intel_loss_*_fwd_bwd`` / ``intel_intent_loss_fwd_bwd``); autograd only chains
the upstream scalar into the stored gradients.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib


class _EnsembleLossFn(torch.autograd.Function):
    """(ens_score, weights) -> float32 scalar; kind in {"list","bpr","mse"}."""

    @staticmethod
    def forward(ctx, ens, weights, batch, kind, cal_div, alpha, noise, seed):
        lib = _lib.load()
        dev = ens.device
        B, L = ens.shape
        K = batch["scores"].shape[2]
        ens_c = ens.contiguous()
        w_c = weights.contiguous() if weights is not None else None
        out = torch.empty(1, dtype=torch.float64, device=dev)
        d_ens = torch.empty(B, L, dtype=torch.float32, device=dev)
        d_w = torch.empty(B, L, K, dtype=torch.float32, device=dev) if cal_div else None
        stream = _lib.stream_ptr(dev)
        args = (B, L, K, _lib.ptr(ens_c, torch.float32), _lib.ptr(w_c, torch.float32) if cal_div else None,
                _lib.ptr(batch["scores"], torch.float64), _lib.ptr(batch["ranking"], torch.int64),
                _lib.ptr(batch["session_len"], torch.int64))
        tail = (int(cal_div), float(alpha), _lib.ptr(out), _lib.ptr(d_ens), _lib.ptr(d_w), stream)
        if kind == "list":
            _lib.check(lib.intel_loss_pl_fwd_bwd(*args, *tail))
        elif kind == "bpr":
            nz = noise.contiguous().float() if noise is not None else None
            _lib.check(lib.intel_loss_bpr_fwd_bwd(*args, _lib.ptr(nz), int(seed), *tail))
        elif kind == "mse":
            _lib.check(lib.intel_loss_mse_fwd_bwd(*args, *tail))
        else:
            raise ValueError(kind)
        ctx.d_ens, ctx.d_w = d_ens, d_w
        ctx.has_w = weights is not None
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        dev = ctx.d_ens.device
        stream = _lib.stream_ptr(dev)
        g = g.contiguous().double()
        ge = torch.empty_like(ctx.d_ens)
        _lib.check(lib.intel_scale_by_device_scalar(ge.numel(), _lib.ptr(ctx.d_ens), _lib.ptr(g), 1.0, None, 0.0,
                                                    _lib.ptr(ge), stream))
        gw = None
        if ctx.d_w is not None and ctx.has_w:
            gw = torch.empty_like(ctx.d_w)
            _lib.check(lib.intel_scale_by_device_scalar(gw.numel(), _lib.ptr(ctx.d_w), _lib.ptr(g), 1.0, None, 0.0,
                                                        _lib.ptr(gw), stream))
        return ge, gw, None, None, None, None, None, None


class _IntentLossFn(torch.autograd.Function):
    """pred intents -> float64 [3] = (intent_loss, ce, kl*T^2)   (BaseIntloss.py:30-67)."""

    @staticmethod
    def forward(ctx, pred, true_intents, kl_weight, kl_temp):
        lib = _lib.load()
        dev = pred.device
        B, I = pred.shape
        pred_c = pred.contiguous()
        out = torch.empty(3, dtype=torch.float64, device=dev)
        d_pred = torch.empty(B, I, dtype=torch.float32, device=dev)
        scratch = torch.empty(4, dtype=torch.int32, device=dev)
        _lib.check(lib.intel_intent_loss_fwd_bwd(B, I, _lib.ptr(pred_c, torch.float32),
                                                 _lib.ptr(true_intents, torch.float64), float(kl_weight),
                                                 float(kl_temp), _lib.ptr(out), _lib.ptr(d_pred), _lib.ptr(scratch),
                                                 _lib.stream_ptr(dev)))
        ctx.d_pred = d_pred
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        dev = ctx.d_pred.device
        g0 = g[0:1].contiguous().double()     # only the combined intent loss carries gradient
        gp = torch.empty_like(ctx.d_pred)
        _lib.check(lib.intel_scale_by_device_scalar(gp.numel(), _lib.ptr(ctx.d_pred), _lib.ptr(g0), 1.0, None, 0.0,
                                                    _lib.ptr(gp), _lib.stream_ptr(dev)))
        return gp, None, None, None


class Baseloss(nn.Module):
    """Baseloss.py:6-20."""
    kind: Optional[str] = None

    @staticmethod
    def parse_loss_args(parser):
        parser.add_argument('--cal_diversity', type=int, default=0)
        parser.add_argument('--diversity_alpha', type=float, default=0.01)
        return parser

    def __init__(self, args):
        self.cal_diversity = args.cal_diversity
        self.diversity_alpha = args.diversity_alpha
        super().__init__()


class BaseIntloss(Baseloss):
    """BaseIntloss.py:10-71."""

    @staticmethod
    def parse_loss_args(parser):
        parser.add_argument('--intent_weight', type=float, default=0.1, help='Weight for intent loss.')
        parser.add_argument('--ensemble_weight', type=float, default=1, help='Weight for ensemble loss.')
        parser.add_argument('--kl_temp', type=float, default=2)
        parser.add_argument('--kl_weight', type=float, default=0.5)
        return Baseloss.parse_loss_args(parser)

    def __init__(self, args):
        self.intent_weight = args.intent_weight
        self.ensemble_weight = args.ensemble_weight
        self.kl_weight, self.T = args.kl_weight, args.kl_temp
        super().__init__(args)
        self.bpr_noise: Optional[torch.Tensor] = None   # test hook: replays torch.rand_like (BPRloss.py:26)
        self.bpr_seed = 0

    def ensemble_loss(self, out_dict: Dict[str, torch.Tensor], in_batch: Dict[str, object]) -> torch.Tensor:
        self.bpr_seed += 1
        out = _EnsembleLossFn.apply(out_dict['ens_score'], out_dict.get('weights') if self.cal_diversity else None,
                                    in_batch, self.kind, int(self.cal_diversity), float(self.diversity_alpha),
                                    self.bpr_noise, self.bpr_seed)
        return out[0].float()

    def get_intloss(self, out_dict, in_batch):
        out = _IntentLossFn.apply(out_dict['intents'], in_batch['intents'], self.kl_weight, self.T)
        return out[0], out[1], out[2]

    def _plain(self, out_dict, in_batch):
        loss = self.ensemble_loss(out_dict, in_batch)
        return loss, loss, loss

    def _with_intent(self, out_dict, in_batch):
        intent_loss, _, _ = self.get_intloss(out_dict, in_batch)
        ensemble_loss = self.ensemble_loss(out_dict, in_batch)
        loss = ensemble_loss * self.ensemble_weight + intent_loss * self.intent_weight
        return loss, ensemble_loss, intent_loss


class Listloss(BaseIntloss):       # Listloss.py
    kind = "list"
    forward = BaseIntloss._plain


class BPRloss(BaseIntloss):        # BPRloss.py
    kind = "bpr"
    forward = BaseIntloss._plain


class MSEloss(BaseIntloss):        # MSEloss.py
    kind = "mse"
    forward = BaseIntloss._plain


class IntListloss(Listloss):       # IntListloss.py:14-19
    forward = BaseIntloss._with_intent


class IntBPRloss(BPRloss):         # IntBPRloss.py:15-20
    forward = BaseIntloss._with_intent


class IntMSEloss(MSEloss):         # IntMSEloss.py:15-20
    forward = BaseIntloss._with_intent
