"""State recording — host-side mirror of ``bindsnet/network/monitors.py`` (``Monitor``
:30-124).  During a fused window the kernels write ``s``/``v`` rasters straight into a
``[T, B, n]`` device buffer (no per-step copies, no Python list churn); ``get`` serves the
reference's ``[time, batch, *shape]`` view from that buffer."""
from __future__ import annotations

from abc import ABC
from typing import Iterable, Optional

import torch


class AbstractMonitor(ABC):
    """Reference: monitors.py:23-27."""


class Monitor(AbstractMonitor):
    """Records state variables of one object over time (reference: monitors.py:30-124)."""

    #: state variables the window kernels can record in-kernel; anything else makes
    #: ``Network.run`` fall back to one-step windows with a host-side snapshot per step.
    FUSED_VARS = ("s", "v")

    def __init__(
        self,
        obj,
        state_vars: Iterable[str],
        time: Optional[int] = None,
        batch_size: int = 1,
        device: str = "cpu",
        sparse: Optional[bool] = False,
    ):
        super().__init__()
        self.obj = obj
        self.state_vars = list(state_vars)
        self.time = time
        self.batch_size = batch_size
        self.device = device
        self.sparse = sparse
        if self.time is None:
            self.device = "cpu"  # monitors.py:68-70
        self.recording = {}
        self.reset_state_variables()

    def get(self, var: str) -> torch.Tensor:
        """``[time, batch, *shape]`` recording (reference: monitors.py:75-92).  With
        ``time=None`` the log is drained by the call, as in the reference."""
        chunks = self.recording[var]
        if self.clean or not chunks:
            return torch.empty(0, device=self.device)
        out = torch.cat(chunks, 0) if len(chunks) > 1 else chunks[0]
        if self.time is None:
            self.recording[var] = []
        else:
            self.recording[var] = [out]
        if self.sparse:
            out = out.to_sparse()
        return out

    def record(self) -> None:
        """Append the object's current value (reference: monitors.py:94-111) — used by the
        one-step fallback and by user code; fused windows call ``_push_window``."""
        for v in self.state_vars:
            data = getattr(self.obj, v)
            if not isinstance(data, torch.Tensor):
                data = torch.as_tensor(data)
            self._push_window(v, data.detach().unsqueeze(0).clone())

    def _push_window(self, var: str, block: torch.Tensor) -> None:
        """Append a ``[t, B, *shape]`` block; keeps only the last ``time`` steps when a
        horizon was given (monitors.py:109-111)."""
        self.clean = False
        block = block.to(self.device)
        chunks = self.recording[var]
        chunks.append(block)
        if self.time is not None:
            total = sum(c.shape[0] for c in chunks)
            while total - chunks[0].shape[0] >= self.time:
                total -= chunks[0].shape[0]
                chunks.pop(0)
            if total > self.time:
                chunks[0] = chunks[0][total - self.time:]

    def reset_state_variables(self) -> None:
        """monitors.py:113-124."""
        self.recording = {v: [] for v in self.state_vars}
        self.clean = True


class SpikeCounter(AbstractMonitor):
    """Per-neuron spike counts of the last ``run`` window, ``[batch, *shape]`` int32 — what the
    reference's callers reduce the full raster to (``spikes.sum(time)``, e.g.
    examples/mnist/batch_eth_mnist.py:280-284; ``evaluation.assign_labels`` consumes exactly
    this, evaluation/evaluation.py:8-61).  The window kernels count in registers, so no
    ``[T, B, n]`` raster is ever written.  Extension: not in the reference.
    """

    def __init__(self, obj, device: str = None):
        super().__init__()
        self.obj = obj
        self.device = device
        self.counts = None

    def get(self, var: str = "s") -> torch.Tensor:
        assert var == "s", "SpikeCounter records spikes only"
        if self.counts is None:
            return torch.empty(0, dtype=torch.int32)
        out = self.counts.view(self.counts.shape[0], *self.obj.shape)
        return out if self.device is None else out.to(self.device)

    def record(self) -> None:  # step-wise fallback path
        s = self.obj.s
        if self.counts is None or self.counts.shape[0] != s.shape[0] or self.counts.device != s.device:
            self.counts = torch.zeros(s.shape[0], self.obj.n, dtype=torch.int32, device=s.device)
        self.counts += s.reshape(s.shape[0], -1).to(torch.int32)

    def _begin_window(self, B: int, device) -> torch.Tensor:
        if self.counts is None or self.counts.shape[0] != B or self.counts.device != device:
            self.counts = torch.zeros(B, self.obj.n, dtype=torch.int32, device=device)
        else:
            self.counts.zero_()
        return self.counts

    def reset_state_variables(self) -> None:
        if self.counts is not None:
            self.counts.zero_()


class NetworkMonitor(AbstractMonitor):
    """Reference: monitors.py:127-329 — whole-network snapshots every step; not on the
    accelerated path."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("NetworkMonitor is outside the hot path bindsnet_b200 implements")
