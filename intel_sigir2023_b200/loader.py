"""Host -> device staging of collate_batch dicts for the B200 path.

The reference moves every batch to the GPU synchronously right before the step (utils.batch_to_gpu, utils.py:91-95,
called from BaseRunner.fit / BaseRunner.predict, BaseRunner.py:262-284 and 300-302).  At B = 4096 the float64
history tensors make one batch 1.5 GB, i.e. ~27 ms of PCIe time against a ~6 ms step, so here the copy of batch
i+1 runs on a side stream while batch i is being processed (same call order and same tensors; the consumer only
waits on the copy's event).  Host batches should live in pinned memory (DataLoader(pin_memory=True) or
`pin_batch`); pageable batches still work but copy synchronously.
"""
from __future__ import annotations

import os
import queue
import threading
from typing import Dict, Iterable, Iterator

import torch

from . import _lib, synthetic

# dense float64 [B,H,I] inputs of the reference's collate_batch that have a compact (index, value) device form
_PACKABLE = ("his_intents", "his_item_int")
# per-session number of real history rows of those tensors (collate_batch pads the rest with zeros and the encoders stop
# at these lengths, GeneralSeq.py:58-106)
_LENGTH_KEY = {"his_intents": "history_len", "his_item_int": "history_item_len"}


def pack_rows(dense: torch.Tensor, nz: int, idx: torch.Tensor, val: torch.Tensor, threads: int = 0,
              lengths: torch.Tensor = None) -> int:
    """Host-side packing of a dense float64 [..., I] CPU tensor into idx int32 / val float32 [..., nz] (pre-allocated,
    ideally pinned).  Returns the largest non-zero count of any row; > nz means the rows were truncated.
    lengths (int64 [B], for a [B,H,I] tensor): only the first lengths[b] rows of session b are real; the padding rows
    behind them are not scanned."""
    if dense.is_cuda or dense.dtype != torch.float64 or not dense.is_contiguous():
        raise ValueError("pack_rows expects a contiguous float64 CPU tensor")
    I = dense.shape[-1]
    rows = dense.numel() // I
    group, lens_ptr = 0, None
    if lengths is not None:
        if dense.dim() != 3 or lengths.is_cuda or lengths.dtype != torch.int64 or lengths.numel() != dense.shape[0]:
            raise ValueError("lengths must be an int64 CPU tensor with one entry per session of a [B,H,I] tensor")
        lengths = lengths.contiguous()
        group, lens_ptr = dense.shape[1], lengths.data_ptr()
    got = _lib.load().intel_host_pack_rows(rows, I, dense.data_ptr(), int(nz), idx.data_ptr(), val.data_ptr(), None, int(threads),
                                           group, lens_ptr)
    if got < 0:
        _lib.check(1)
    return int(got)


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

    def __init__(self, batches: Iterable[Dict[str, object]], device, depth: int = 2, pack_history: bool = False,
                 pack_nz: int = 16, threads: int = 0):
        """pack_history: scan the dense float64 history-intent tensors on the host (all cores, `intel_host_pack_rows`)
        and ship only their non-zeros: the yielded dicts then carry `his_intents_idx/_val` (and `his_item_int_idx/_val`)
        instead of the dense tensors - the model accepts either.  pack_nz is the initial capacity per row; it grows
        when a batch needs more, and a tensor whose rows are more than a quarter full stays dense."""
        self.batches, self.device, self.depth = batches, torch.device(device), max(2, int(depth))
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher stages batches into GPU memory; got device %s" % (self.device,))
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [dict() for _ in range(self.depth)]
        self.free_ev = [None] * self.depth          # consumer-stream event: the slot's previous batch is no longer in use
        if threads <= 0:        # share the host cores between the ranks of a node (torchrun exports LOCAL_WORLD_SIZE)
            try:
                cores = len(os.sched_getaffinity(0))          # the cores this process may run on (what `nproc` reports)
            except (AttributeError, OSError):
                cores = os.cpu_count() or 1
            threads = max(1, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
        self.pack_history, self.threads = bool(pack_history), int(threads)
        self.nz = {k: max(1, int(pack_nz)) for k in _PACKABLE}
        self.host = [dict() for _ in range(self.depth)]     # pinned staging of the packed tensors, per slot
        self.copied_ev = [None] * self.depth                # copy-stream event: the slot's pinned staging may be rewritten

    def _pack(self, key: str, v: torch.Tensor, slot: int, lengths=None):
        """dense [B,H,I] float64 on the host -> pinned (idx, val) [B,H,nz], or None to keep the tensor dense"""
        I = v.shape[-1]
        while True:
            nz = self.nz[key]
            if 4 * nz > I:
                return None
            shape = tuple(v.shape[:-1]) + (nz,)
            st = self.host[slot].get(key)
            if st is None or st[0].shape != shape:
                st = (torch.empty(shape, dtype=torch.int32).pin_memory(), torch.empty(shape, dtype=torch.float32).pin_memory())
                self.host[slot][key] = st
            got = pack_rows(v, nz, st[0], st[1], self.threads, lengths)
            if got <= nz:
                return st
            self.nz[key] = max(got, 2 * nz)

    def _stage(self, host: Dict[str, object], slot: int):
        bufs = self.slots[slot]
        items = dict(host)
        if self.pack_history:
            if self.copied_ev[slot] is not None:
                self.copied_ev[slot].synchronize()      # the previous copy out of this slot's pinned staging has finished
            for key in _PACKABLE:
                v = items.get(key)
                if torch.is_tensor(v) and not v.is_cuda and v.dtype == torch.float64 and v.dim() == 3:
                    lens = items.get(_LENGTH_KEY[key])
                    if not (torch.is_tensor(lens) and not lens.is_cuda and lens.dtype == torch.int64 and lens.numel() == v.shape[0]):
                        lens = None
                    packed = self._pack(key, v.contiguous(), slot, lens)
                    if packed is not None:
                        del items[key]
                        items[key + "_idx"], items[key + "_val"] = packed
        if self.free_ev[slot] is not None:
            self.stream.wait_event(self.free_ev[slot])
        out: Dict[str, object] = {}
        with torch.cuda.stream(self.stream):
            for k, v in items.items():
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
        self.copied_ev[slot] = ev
        return out, ev, slot

    def __iter__(self) -> Iterator[Dict[str, object]]:
        if self.pack_history:
            yield from self._iter_threaded()
            return
        staged = []
        n = 0
        for host in self.batches:
            staged.append(self._stage(host, n % self.depth))
            n += 1
            if len(staged) < self.depth:
                continue
            yield from self._hand_over(staged.pop(0))
        while staged:
            yield from self._hand_over(staged.pop(0))

    def _iter_threaded(self) -> Iterator[Dict[str, object]]:
        """Packing costs host time (a memory-bound scan of the dense tensors): a worker thread packs and stages batch
        i+1 while the consumer thread enqueues step i (the ctypes call releases the GIL)."""
        free_q: "queue.Queue" = queue.Queue()
        out_q: "queue.Queue" = queue.Queue()
        for slot in range(self.depth):
            free_q.put((slot, None))

        def worker():
            try:
                torch.cuda.set_device(self.device)
                for host in self.batches:
                    slot, ev = free_q.get()
                    if slot is None:
                        return
                    self.free_ev[slot] = ev
                    out_q.put(self._stage(host, slot))
                out_q.put(None)
            except BaseException as exc:        # surfaced in the consumer thread
                out_q.put(exc)

        th = threading.Thread(target=worker, name="intel-b200-prefetch", daemon=True)
        th.start()
        try:
            while True:
                item = out_q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                dev, ev, slot = item
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(ev)
                yield dev
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(self.device))
                free_q.put((slot, done))
        finally:
            free_q.put((None, None))
            th.join(timeout=30)

    def _hand_over(self, staged):
        dev, ev, slot = staged
        torch.cuda.current_stream(self.device).wait_event(ev)
        yield dev
        # the consumer is back for the next batch: everything that reads this slot has been enqueued on its stream
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.device))
        self.free_ev[slot] = done
