"""Host→device input pipelining for ``Network.run`` windows.

The reference's callers (``examples/mnist/batch_eth_mnist.py:262-277``, the ``DataLoader``
pipelines) hand ``Network.run`` a freshly encoded ``[time, batch, ...]`` uint8 spike tensor per
batch.  At ~2 ms of GPU time per window the 25 MB host→device copy of the next batch is worth
hiding: ``WindowPrefetcher`` copies batch k+1 on a side stream (pinned host memory, double-buffered
device staging) while the window kernel of batch k runs, and hands ``run`` a device tensor.

    pre = WindowPrefetcher(device, iter_of_host_tensors)
    for x_dev in pre:                 # x_dev is ready on the compute stream
        net.run({"X": x_dev}, time=T)  # the buffer is handed back when the loop asks for the next item

``AsyncReadback`` is the other direction: the per-window result (spike counts for label assignment,
``examples/mnist/batch_eth_mnist.py:300-318``) is copied to pinned host memory without stalling the
launch of the next window; the host consumes window k-1's result while window k runs.
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
        self.release()                             # whatever was enqueued since the last item used that buffer
        if self._pending is None:
            raise StopIteration
        slot = self._pending
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ready[slot])          # the consumer's stream waits for the copy
        out = self._bufs[slot]
        out.record_stream(cur)                     # allocated on the copy stream, consumed on this one
        self._last = slot
        self._issue()                              # start copying the next item right away
        return out

    def release(self) -> None:
        """Hand the tensor returned last back to the prefetcher: the copy that reuses its buffer waits for
        everything enqueued on the current stream so far.  Called automatically when the next item is
        requested; call it earlier (right after ``run``) to let that copy start sooner."""
        slot = getattr(self, "_last", None)
        if slot is None:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[slot] = ev
        self._last = None


class AsyncReadback:
    """Device→host result pipeline: ``push`` snapshots a device tensor (device-to-device, on the
    caller's stream), then copies the snapshot into a ring of pinned host buffers on a side stream,
    so that neither the copy nor the host's read of it delays the next window's launch; ``pop``
    returns the oldest copy once it has landed (waiting only on that copy's event).

        rb = AsyncReadback(depth=2)
        for x_dev in pre:
            net.run({"X": x_dev}, time=T)
            rb.push(counter.get("s"))
            if len(rb) == rb.depth:      # window k-1's counts, while window k runs
                use(rb.pop())
        while len(rb): use(rb.pop())
    """

    def __init__(self, depth: int = 2):
        self.depth = max(1, depth)
        self._host = [None] * self.depth
        self._snap = [None] * self.depth
        self._done = [None] * self.depth   # event: D2H out of the snapshot finished
        self._queue = []                   # slots in submission order
        self._k = 0
        self._stream = None

    def __len__(self) -> int:
        return len(self._queue)

    def push(self, t: torch.Tensor) -> None:
        if len(self._queue) >= self.depth:
            raise RuntimeError("AsyncReadback ring is full: pop() before pushing again")
        slot = self._k % self.depth
        self._k += 1
        h = self._host[slot]
        if h is None or h.shape != t.shape or h.dtype != t.dtype:
            h = torch.empty(t.shape, dtype=t.dtype)
            if t.is_cuda:
                h = h.pin_memory()
            self._host[slot] = h
        if not t.is_cuda:  # host tensor (CPU tests): nothing to overlap
            h.copy_(t)
            self._done[slot] = None
            self._queue.append(slot)
            return
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=t.device)
        cur = torch.cuda.current_stream(t.device)
        snap = self._snap[slot]
        if snap is None or snap.shape != t.shape or snap.dtype != t.dtype:
            snap = torch.empty_like(t)
            self._snap[slot] = snap
        if self._done[slot] is not None:
            cur.wait_event(self._done[slot])   # the previous copy out of this snapshot (long finished)
        snap.copy_(t)                          # stream-ordered after the producer of ``t``
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ready)
            h.copy_(snap, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._stream)
        self._done[slot] = done
        self._queue.append(slot)

    def pop(self) -> torch.Tensor:
        """Oldest result as a pinned host tensor (valid until ``depth`` further pushes)."""
        slot = self._queue.pop(0)
        ev = self._done[slot]
        if ev is not None:
            # poll instead of cudaEventSynchronize: a blocking wait can return milliseconds late
            # (measured on virtualised hosts), which would stall the launch of the next window
            while not ev.query():
                pass
        return self._host[slot]
