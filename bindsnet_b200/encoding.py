"""Spike-train encoders — own restatement of the two encoders the hot path's callers use
(reference: bindsnet/encoding/encodings.py ``bernoulli`` :50-96, ``poisson`` :99-156).  They
produce the ``[time, *shape]`` uint8 tensors ``Network.run`` consumes; bench.py uses
``poisson`` to synthesise the metric's 28x28 input.

CUDA tensors are encoded ON THE DEVICE by the kernels of ``csrc/snn_encode.cu`` (``snn_b200_encode_poisson`` /
``snn_b200_encode_bernoulli``): only the rate image crosses PCIe, the ``[time, batch, ...]``
spike tensor is written where ``Network.run`` reads it.  Same distributions as the reference, a counter-based
random stream of their own (``seed=`` keyword; default: drawn from torch's CPU generator).  CPU tensors take the
torch restatement below (what the golden fixtures and the CPU tests use)."""
from __future__ import annotations

from typing import Optional

import torch


def bernoulli(datum: torch.Tensor, time: Optional[int] = None, dt: float = 1.0, device="cpu", **kwargs) -> torch.Tensor:
    """Independent Bernoulli trials per step with probability proportional to intensity
    (reference semantics: encodings.py:50-96)."""
    max_prob = kwargs.get("max_prob", 1.0)
    assert 0 <= max_prob <= 1, "Maximum firing probability must be in range [0, 1]"
    if datum.is_cuda and time is not None:
        return _bernoulli_cuda(datum, int(time / dt), max_prob, kwargs.get("seed"))
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
    if datum.is_cuda:
        return _poisson_cuda(datum, int(time / dt), dt, kwargs.get("seed"))
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


def _seed(seed) -> int:
    return int(torch.randint(0, 2**62, (1,)).item()) if seed is None else int(seed)


def _poisson_cuda(datum: torch.Tensor, steps: int, dt: float, seed=None) -> torch.Tensor:
    """``poisson`` on the device (no host synchronisation: negative rates are clamped to silence by the kernel's
    ``rate > 0`` test instead of asserted)."""
    from . import _backend

    rate = datum.detach().float().contiguous()
    out = torch.empty((steps,) + tuple(datum.shape), dtype=torch.uint8, device=datum.device)
    _backend.encode_poisson(rate, steps, dt, _seed(seed), out)
    return out


def _bernoulli_cuda(datum: torch.Tensor, steps: int, max_prob: float, seed=None) -> torch.Tensor:
    from . import _backend

    p = datum.detach().float()
    p = torch.where(p.max() > 1.0, p / p.max(), p)   # stays on the device (encodings.py:80-81)
    p = (max_prob * p).contiguous()
    out = torch.empty((steps,) + tuple(datum.shape), dtype=torch.uint8, device=datum.device)
    _backend.encode_bernoulli(p, steps, _seed(seed), out)
    return out
