"""Host→device input pipelining for ``Network.run`` windows.

The reference's callers (``examples/mnist/batch_eth_mnist.py:262-277``, the ``DataLoader``
pipelines) hand ``Network.run`` a freshly encoded ``[time, batch, ...]`` uint8 spike tensor per
batch.  At ~2 ms of GPU time per window the 25 MB host→device copy of the next batch is worth
hiding: ``WindowPrefetcher`` copies batch k+1 on a side stream (pinned host memory, double-buffered
device staging) while the window kernel of batch k runs, and hands ``run`` a device tensor.

    pre = WindowPrefetcher(device, iter_of_host_tensors)
    for x_dev in pre:                 # x_dev is ready on the compute stream
        net.run({"X": x_dev}, time=T)
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional

import torch


class WindowPrefetcher:
    """Iterates over host spike tensors, yielding device copies; the copy of item k+1 overlaps
    whatever the caller launches on the current stream for item k."""

    def __init__(self, device, source: Iterable[torch.Tensor], depth: int = 2):
        self.device = torch.device(device)
        self.source: Iterator[torch.Tensor] = iter(source)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.depth = max(2, depth)
        self._bufs = [None] * self.depth           # device staging buffers
        self._ready = [None] * self.depth          # event: H2D copy into buffer finished
        self._free = [None] * self.depth           # event: consumer finished with buffer
        self._k = 0
        self._pending: Optional[int] = None
        self._issue()

    def _issue(self) -> None:
        try:
            x = next(self.source)
        except StopIteration:
            self._pending = None
            return
        slot = self._k % self.depth
        self._k += 1
        if not x.is_pinned():
            x = x.pin_memory()
        with torch.cuda.stream(self.copy_stream):
            if self._free[slot] is not None:
                self.copy_stream.wait_event(self._free[slot])  # previous consumer of this buffer done
            buf = self._bufs[slot]
            if buf is None or buf.shape != x.shape or buf.dtype != x.dtype:
                buf = torch.empty(x.shape, dtype=x.dtype, device=self.device)
                self._bufs[slot] = buf
            buf.copy_(x, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            self._ready[slot] = ev
        self._pending = slot

    def __iter__(self):
        return self

    def __next__(self) -> torch.Tensor:
        if self._pending is None:
            raise StopIteration
        slot = self._pending
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ready[slot])          # the consumer's stream waits for the copy
        out = self._bufs[slot]
        self._issue()                              # start copying the next item right away
        # mark the buffer free once everything the caller enqueues before its NEXT __next__ is done:
        # recorded lazily at the next call on the same slot via this event
        ev = torch.cuda.Event()
        self._free[slot] = ev
        self._last = (slot, ev)
        return out

    def release(self) -> None:
        """Record that the consumer is done with the tensor returned last (call after ``run``)."""
        slot, ev = self._last
        ev.record(torch.cuda.current_stream(self.device))
