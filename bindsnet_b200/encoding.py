"""Spike-train encoders — own restatement of the encoders the hot path's callers use
(reference: bindsnet/encoding/encodings.py ``bernoulli`` :50-96, ``poisson`` :99-156 — the two with kernels of their
own — and the deterministic ``single`` :6-33, ``repeat`` :36-47, ``rank_order`` :159-191 as device-resident tensor
code; ``encoders.py``'s callable wrappers).  They
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
    if kwargs.get("approx", False):
        # encodings.py:130-137: |N(0,1)| ** ((rate * 0.11 + 5) / 50) < 0.6, torch's generator on the datum's device
        dev = datum.device if datum.is_cuda else torch.device(device)
        steps = int(time / dt)
        x = torch.randn((steps, datum.numel()), device=dev).abs()
        x = torch.pow(x, (datum.flatten().to(dev) * 0.11 + 5) / 50)
        return (x < 0.6).view(steps, *datum.shape).byte()
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


def single(datum: torch.Tensor, time: int, dt: float = 1.0, sparsity: float = 0.5, device="cpu", **kwargs) -> torch.Tensor:
    """One spike at the first step for every feature above the ``1 - sparsity`` quantile, silence afterwards
    (reference semantics: encodings.py:6-33).  Deterministic; computed where ``datum`` lives (CUDA tensors stay on
    the device), else on ``device``."""
    steps = int(time / dt)
    dev = datum.device if datum.is_cuda else torch.device(device)
    d = datum.detach().to(dev)
    s = torch.zeros((steps, *d.shape), dtype=torch.uint8, device=dev)
    s[0] = (d > torch.quantile(d, 1 - sparsity)).to(torch.uint8)
    return s


def repeat(datum: torch.Tensor, time: int, dt: float = 1.0, **kwargs) -> torch.Tensor:
    """``datum`` at every one of the ``int(time / dt)`` steps (encodings.py:36-47)."""
    steps = int(time / dt)
    return datum.repeat([steps, *([1] * datum.dim())])


def rank_order(datum: torch.Tensor, time: int, dt: float = 1.0, device="cpu", **kwargs) -> torch.Tensor:
    """At most one spike per feature, the stronger the earlier: feature ``i`` fires at step
    ``ceil(steps * (max / x_i) / max_j(max / x_j)) - 1`` when that lies inside the window; zero-valued features and those
    landing on the last step stay silent (reference semantics: encodings.py:159-191, same float32 operations in the
    same order, the per-feature loop replaced by one scatter).  Unlike the reference the caller's tensor is not
    normalised in place."""
    assert (datum >= 0).all(), "Inputs must be non-negative"
    shape, size = datum.shape, datum.numel()
    dev = datum.device if datum.is_cuda else torch.device(device)
    d = datum.detach().flatten().to(dev).clone()
    steps = int(time / dt)
    d /= d.max()
    times = torch.zeros(size, device=dev)
    live = d != 0
    times[live] = 1 / d[live]
    times *= steps / times.max()
    times = torch.ceil(times).long()
    fires = (times > 0) & (times < steps)
    spikes = torch.zeros(steps, size, dtype=torch.uint8, device=dev)
    spikes[times[fires] - 1, torch.arange(size, device=dev)[fires]] = 1
    return spikes.reshape(steps, *shape)


class Encoder:
    """Callable that applies one of the encodings above with fixed arguments (reference: encoders.py:4-18)."""

    def __init__(self, *args, **kwargs) -> None:
        self.enc_args = args
        self.enc_kwargs = kwargs

    def __call__(self, img):
        return self.enc(img, *self.enc_args, **self.enc_kwargs)


class NullEncoder(Encoder):
    """Hands the datum through unchanged (encoders.py:21-34)."""

    def __init__(self):
        super().__init__()

    def __call__(self, img):
        return img


def _encoder(name: str, fn, where: str, **defaults):
    def __init__(self, time: int, dt: float = 1.0, **kwargs):
        Encoder.__init__(self, time, dt=dt, **{**defaults, **kwargs})
        self.enc = fn

    return type(name, (Encoder,), {"__init__": __init__, "__doc__": f"``{fn.__name__}`` with fixed arguments (reference: {where})."})


SingleEncoder = _encoder("SingleEncoder", single, "encoders.py:37-51", sparsity=0.5)
RepeatEncoder = _encoder("RepeatEncoder", repeat, "encoders.py:54-66")
BernoulliEncoder = _encoder("BernoulliEncoder", bernoulli, "encoders.py:69-85")
PoissonEncoder = _encoder("PoissonEncoder", poisson, "encoders.py:88-102", approx=False)
RankOrderEncoder = _encoder("RankOrderEncoder", rank_order, "encoders.py:105-117")


def _loader(fn, where: str, pass_max_prob: bool = False):
    def loader(data, time=None, dt: float = 1.0, **kwargs):
        for i in range(len(data)):
            if pass_max_prob:   # loaders.py:31: the reference reads ``max_prob`` from the ``dt`` keyword (always 1.0)
                yield fn(datum=data[i], time=time, dt=dt, max_prob=kwargs.get("dt", 1.0))
            else:
                yield fn(datum=data[i], time=time, dt=dt)

    loader.__name__ = f"{fn.__name__}_loader"
    loader.__doc__ = f"Lazily encodes ``data[i]`` with ``{fn.__name__}``, one item per ``next`` (reference: {where})."
    return loader


bernoulli_loader = _loader(bernoulli, "loaders.py:8-33", pass_max_prob=True)
poisson_loader = _loader(poisson, "loaders.py:36-55")
rank_order_loader = _loader(rank_order, "loaders.py:58-77")
