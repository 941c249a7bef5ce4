"""Host -> device feed of the training loop (the `.to(device)` of src/train.py:381-382), one batch ahead.

``DevicePrefetcher(batches, device)`` wraps any iterable of host batches (a tensor, or a tuple / list of tensors; pinned memory
makes the copies asynchronous) and yields them as device tensors.  The copy of batch i+1 is issued on a side stream as soon as
batch i has been handed out, so it runs on the copy engine while the kernels of step i execute; the consumer's stream waits on
the copy's event (no host synchronisation anywhere).  Every batch is still copied exactly once, inside the loop that uses it."""
from __future__ import annotations

from typing import Iterable, Iterator

import torch


class DevicePrefetcher:
    def __init__(self, batches: Iterable, device) -> None:
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("srcgan_b200.data.DevicePrefetcher: a CUDA device is required (there is no CPU fallback)")
        self._it: Iterator = iter(batches)
        self._stream = torch.cuda.Stream(device=self.device)
        self._next = None
        self._issue()

    def _copy(self, item):
        if torch.is_tensor(item):
            return item.to(self.device, non_blocking=True)
        if isinstance(item, (tuple, list)):
            return type(item)(self._copy(t) for t in item)
        return item

    def _issue(self) -> None:
        try:
            host = next(self._it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self._stream):
            dev = self._copy(host)
            ev = torch.cuda.Event()
            ev.record(self._stream)
        self._next = (dev, ev)

    @staticmethod
    def _record(item, stream) -> None:
        if torch.is_tensor(item):
            item.record_stream(stream)
        elif isinstance(item, (tuple, list)):
            for t in item:
                DevicePrefetcher._record(t, stream)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        dev, ev = self._next
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)                      # stream-side wait: the host does not block
        self._record(dev, cur)                  # the caching allocator must not hand the memory back to the side stream early
        self._issue()                           # next batch: copy engine, concurrent with this step's kernels
        return dev
