"""Spike-train encoders — own restatement of the two encoders the hot path's callers use
(reference: bindsnet/encoding/encodings.py ``bernoulli`` :50-96, ``poisson`` :99-156).  They
produce the ``[time, *shape]`` uint8 tensors ``Network.run`` consumes; bench.py uses
``poisson`` to synthesise the metric's 28x28 input.  Host-side (an on-device encoder is the
first "next" row of SURVEY.md §8f)."""
from __future__ import annotations

from typing import Optional

import torch


def bernoulli(datum: torch.Tensor, time: Optional[int] = None, dt: float = 1.0, device="cpu", **kwargs) -> torch.Tensor:
    """Independent Bernoulli trials per step with probability proportional to intensity
    (reference semantics: encodings.py:50-96)."""
    max_prob = kwargs.get("max_prob", 1.0)
    assert 0 <= max_prob <= 1, "Maximum firing probability must be in range [0, 1]"
    assert (datum >= 0).all(), "Inputs must be non-negative"
    shape = datum.shape
    p = datum.flatten().to(device).float()
    if p.max() > 1.0:
        p = p / p.max()
    p = max_prob * p
    if time is None:
        return torch.bernoulli(p).view(*shape).byte()
    steps = int(time / dt)
    return torch.bernoulli(p.expand(steps, -1)).view(steps, *shape).byte()


def poisson(datum: torch.Tensor, time: int, dt: float = 1.0, device="cpu", **kwargs) -> torch.Tensor:
    """Spike trains whose inter-spike intervals are Poisson(1000 / (rate * dt)) distributed
    steps, zero intervals bumped to one (reference semantics: encodings.py:99-156; ``datum``
    is the firing rate in Hz)."""
    assert (datum >= 0).all(), "Inputs must be non-negative"
    shape, size = datum.shape, datum.numel()
    rate_hz = datum.flatten().to(device).float()
    steps = int(time / dt)
    active = rate_hz != 0
    mean_isi = torch.zeros(size, device=device)
    mean_isi[active] = (1000.0 / dt) / rate_hz[active]
    isi = torch.poisson(mean_isi.expand(steps + 1, -1))
    isi[:, active] += (isi[:, active] == 0).float()
    when = torch.cumsum(isi, dim=0).long()
    when[when >= steps + 1] = 0
    spikes = torch.zeros(steps + 1, size, device=device, dtype=torch.uint8)
    spikes[when, torch.arange(size, device=device)] = 1
    return spikes[1:].view(steps, *shape)
