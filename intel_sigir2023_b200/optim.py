"""Fused Adam for the IntEL training loop.

Drop-in for ``torch.optim.Adam(model.customize_parameters(), lr=..., weight_decay=l2)`` as built by
``BaseRunner._build_optimizer`` (BaseRunner.py:182-188): same constructor arguments, same ``param_groups`` / ``state``
layout (``step``, ``exp_avg``, ``exp_avg_sq``), so ``state_dict()`` interchanges with torch's optimizer and
``torch.optim.lr_scheduler.StepLR`` works on it.  ``step()`` updates all parameters with one kernel launch per 48
tensors (``intel_adam_step``) instead of ~10 small kernels per tensor.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable

import torch

from . import _lib


def customize_parameters(model: torch.nn.Module) -> list:
    """The reference's two parameter groups (BaseModel.customize_parameters, BaseModel.py:53-62)."""
    weight_p, bias_p = [], []
    for name, p in model.named_parameters():
        if p.requires_grad:
            (bias_p if "bias" in name else weight_p).append(p)
    return [{"params": weight_p}, {"params": bias_p, "weight_decay": 0}]


class Adam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._cache = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            cache = self._cache.get(gi)
            if cache is None or cache["ids"] != [id(p) for p in ps] or any(not self.state[p] for p in ps):
                cache = self._build_cache(ps)
                self._cache[gi] = cache
            step_t = cache["step"]
            step_t += 1                                   # one shared tensor: every parameter's state["step"] aliases it
            n = len(ps)
            grads, params, ms, vs = cache["grads"], cache["params"], cache["m"], cache["v"]
            for i, p in enumerate(ps):
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device:
                    raise ValueError("intel_adam_step needs contiguous float32 gradients on the parameter's device")
                grads[i] = g.data_ptr()
                # the Parameter object can outlive its storage (model.to(), p.data = ..., a replaced state tensor):
                # re-read the pointers every step like torch.optim.Adam re-reads the tensors
                st = self.state[p]
                params[i], ms[i], vs[i] = p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            b1, b2 = group["betas"]
            wd = float(group["weight_decay"])
            if wd != cache["wd_value"]:
                cache["wd"] = (C.c_float * n)(*([wd] * n))
                cache["wd_value"] = wd
            _lib.check(lib.intel_adam_step(n, cache["params"], grads, cache["m"], cache["v"], cache["numel"], cache["wd"],
                                           float(group["lr"]), float(b1), float(b2), float(group["eps"]), int(step_t),
                                           _lib.stream_ptr(ps[0].device)))
        return loss

    def _build_cache(self, ps):
        """pointer tables of a group (parameters and moments keep their storage; gradients are refreshed every step)"""
        n = len(ps)
        step_t = None
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError("intel_adam_step needs contiguous float32 parameters")
            st = self.state[p]
            if not st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = torch.tensor(0.0)
            if step_t is None:
                step_t = st["step"] if torch.is_tensor(st["step"]) else torch.tensor(float(st["step"]))
            elif float(st["step"]) != float(step_t):
                raise ValueError("parameters of one group must share their step count")
        for p in ps:
            self.state[p]["step"] = step_t
        arr = lambda ts: (C.c_void_p * n)(*[_lib.ptr(t) for t in ts])
        return {"ids": [id(p) for p in ps], "step": step_t, "params": arr(ps),
                "m": arr([self.state[p]["exp_avg"] for p in ps]), "v": arr([self.state[p]["exp_avg_sq"] for p in ps]),
                "numel": (C.c_int64 * n)(*[p.numel() for p in ps]), "grads": (C.c_void_p * n)(), "wd": None, "wd_value": None}

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._cache = {}
