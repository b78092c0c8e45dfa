"""Host -> device staging of collate_batch dicts for the B200 path.

The reference moves every batch to the GPU synchronously right before the step (utils.batch_to_gpu, utils.py:91-95,
called from BaseRunner.fit / BaseRunner.predict, BaseRunner.py:262-284 and 300-302).  At B = 4096 the float64
history tensors make one batch 1.5 GB, i.e. ~27 ms of PCIe time against a ~6 ms step, so here the copy of batch
i+1 runs on a side stream while batch i is being processed (same call order and same tensors; the consumer only
waits on the copy's event).  Host batches should live in pinned memory (DataLoader(pin_memory=True) or
`pin_batch`); pageable batches still work but copy synchronously.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator

import torch

from . import synthetic


def pin_batch(batch: Dict[str, object]) -> Dict[str, object]:
    """Pinned-memory copy of every tensor of a host batch."""
    return {k: (v.pin_memory() if torch.is_tensor(v) and not v.is_cuda else v) for k, v in batch.items()}


class DevicePrefetcher:
    """Iterates device-resident batches, staging `depth - 1` batches ahead on a copy stream.

    The device tensors live in `depth` persistent slots (re-allocated only when a shape changes), so the steady state
    never touches the allocator: a 1.5 GB batch per step would otherwise cost a cudaMalloc / cudaFree pair that
    serialises with the running step.  A slot is refilled only after the consumer has asked for the batch after it,
    i.e. after all work on the slot's previous batch has been enqueued (guarded by an event on the consumer stream).
    The yielded dict is only valid until the next `depth - 1` batches have been requested."""

    def __init__(self, batches: Iterable[Dict[str, object]], device, depth: int = 2):
        self.batches, self.device, self.depth = batches, torch.device(device), max(2, int(depth))
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher stages batches into GPU memory; got device %s" % (self.device,))
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [dict() for _ in range(self.depth)]
        self.free_ev = [None] * self.depth          # consumer-stream event: the slot's previous batch is no longer in use

    def _stage(self, host: Dict[str, object], slot: int):
        bufs = self.slots[slot]
        if self.free_ev[slot] is not None:
            self.stream.wait_event(self.free_ev[slot])
        out: Dict[str, object] = {}
        with torch.cuda.stream(self.stream):
            for k, v in host.items():
                if not torch.is_tensor(v):
                    out[k] = v
                    continue
                buf = bufs.get(k)
                if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                    buf = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                    bufs[k] = buf
                buf.copy_(v, non_blocking=True)
                out[k] = buf
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev, slot

    def __iter__(self) -> Iterator[Dict[str, object]]:
        queue = []
        n = 0
        for host in self.batches:
            queue.append(self._stage(host, n % self.depth))
            n += 1
            if len(queue) < self.depth:
                continue
            yield from self._hand_over(queue.pop(0))
        while queue:
            yield from self._hand_over(queue.pop(0))

    def _hand_over(self, staged):
        dev, ev, slot = staged
        torch.cuda.current_stream(self.device).wait_event(ev)
        yield dev
        # the consumer is back for the next batch: everything that reads this slot has been enqueued on its stream
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.device))
        self.free_ev[slot] = done
