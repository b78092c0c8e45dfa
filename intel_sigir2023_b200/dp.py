"""Session-sharded data parallelism (absent in the reference, which is single process).

One process per GPU.  A global batch is cut into equal contiguous shards of sessions
(``synthetic.shard_batch`` keeps the global padded widths so pad keys take part in the unmasked
self-attention exactly as in a single-process run); tables and weights are replicated; the only
exchange step of the path is the gradient all-reduce (mean of the local batch-mean losses = global
batch mean), plus a sum of the additive evaluation terms.  NCCL over NVLink on the GPU box, gloo in
the CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.distributed as dist


class GradReducer:
    """Average the dense ``.grad`` buffers over the ranks after backward.

    The few large tensors (embedding tables) are reduced in place, one collective each, so no
    flatten/unflatten copies of ~100 MB are made; the many small weights travel in one flat bucket."""

    def __init__(self, model: torch.nn.Module, world: int, small_numel: int = 1 << 16):
        self.model = model
        self.params: List[torch.nn.Parameter] = [p for p in model.parameters() if p.requires_grad]
        self.world = world
        self.small_numel = small_numel
        backend = dist.get_backend() if dist.is_initialized() else "none"
        self.avg_op = dist.ReduceOp.AVG if backend == "nccl" else dist.ReduceOp.SUM
        self.post_scale = 1.0 if backend == "nccl" else 1.0 / world

    def allreduce(self) -> None:
        if self.world <= 1:
            return
        grads = [p.grad for p in self.params if p.grad is not None]
        # the backward pass writes all gradients into one zero-filled buffer (IntEL._IntelFn.backward): if every .grad
        # still aliases it (autograd adopted the views instead of copying them), one collective covers everything
        flat = getattr(self.model, "_flat_grad", None)
        if flat is not None and grads and len(grads) == len(self.params):
            base = flat.untyped_storage().data_ptr()
            if all(g.untyped_storage().data_ptr() == base for g in grads):
                dist.all_reduce(flat, op=self.avg_op)
                if self.post_scale != 1.0:
                    flat.mul_(self.post_scale)
                return
        big = [g for g in grads if g.numel() > self.small_numel]
        small = [g for g in grads if g.numel() <= self.small_numel]
        works = [dist.all_reduce(g, op=self.avg_op, async_op=True) for g in big]
        if small:
            flat = torch.cat([g.reshape(-1) for g in small])
            dist.all_reduce(flat, op=self.avg_op)
            if self.post_scale != 1.0:
                flat.mul_(self.post_scale)
            off = 0
            for g in small:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        for w in works:
            w.wait()
        if self.post_scale != 1.0:
            for g in big:
                g.mul_(self.post_scale)


def global_max_len(session_len: torch.Tensor) -> int:
    """max(session_len) over every rank: evaluate_method's max_len is a global quantity (BaseRunner.py:66)."""
    m = session_len.max().reshape(1).clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return int(m.item())


def evaluate_sharded(pred: torch.Tensor, ranking: torch.Tensor, pos_nums: Dict[str, torch.Tensor],
                     session_len: torch.Tensor, topk: Sequence[int], metrics: Sequence[str]) -> Dict[str, float]:
    """Every rank scores its shard of the eval sessions; the additive per-metric sums are all-reduced."""
    from . import evaluate
    max_len = max(global_max_len(session_len), max(topk))
    sums, counts = evaluate.ndcg_sums(pred, ranking, session_len, pos_nums['c_paynum_i'], pos_nums['c_favnum_i'],
                                      pos_nums['c_clicknum_i'], max_len, topk)
    if dist.is_initialized() and dist.get_world_size() > 1:
        both = torch.cat([sums, counts])
        dist.all_reduce(both, op=dist.ReduceOp.SUM)
        sums, counts = both[:sums.numel()], both[sums.numel():]
    return evaluate.metrics_from_sums(sums.cpu().numpy(), counts.cpu().numpy(), topk, metrics)
