"""Synapse matrices — host-side mirror of ``bindsnet/network/topology.py``.

``Connection`` (reference: topology.py:265-399) and ``MulticompartmentConnection`` with a
``Weight`` feature (topology.py:402-537) keep the reference's constructor signatures and
attributes (``w``, ``b``, ``wmin``, ``wmax``, ``norm``, ``source``, ``target``,
``update_rule`` / ``pipeline``).  ``compute``/``update``/``normalize`` of a whole network run
inside the CUDA window kernels; the standalone methods below submit single-connection
work to the same kernels.
"""
from __future__ import annotations

import warnings
from abc import ABC
from typing import Optional, Sequence, Union

import numpy as np
import torch
from torch.nn import Module, Parameter

from .. import _abi
from .nodes import Nodes, _scalar


class AbstractConnection(ABC, Module):
    """Reference: topology.py:17-156."""

    def __init__(
        self,
        source: Nodes,
        target: Nodes,
        nu: Optional[Union[float, Sequence[float], Sequence[torch.Tensor]]] = None,
        reduction: Optional[callable] = None,
        weight_decay: float = 0.0,
        **kwargs,
    ) -> None:
        super().__init__()
        assert isinstance(source, Nodes), "Source is not a Nodes object"
        assert isinstance(target, Nodes), "Target is not a Nodes object"
        self.source = source
        self.target = target
        self.weight_decay = weight_decay
        self.reduction = reduction

        from ..learning import NoOp

        self.wmin = Parameter(torch.as_tensor(kwargs.get("wmin", -np.inf), dtype=torch.float32), requires_grad=False)
        self.wmax = Parameter(torch.as_tensor(kwargs.get("wmax", np.inf), dtype=torch.float32), requires_grad=False)
        self.norm = kwargs.get("norm", None)
        self.decay = kwargs.get("decay", None)
        if kwargs.get("Dales_rule", None) is not None:
            raise NotImplementedError("Dales_rule is not implemented by the CUDA core (DESIGN.md 'Out of scope')")
        self.Dales_rule = None

        rule_cls = kwargs.get("update_rule", None) or NoOp
        self.update_rule = rule_cls(connection=self, nu=nu, reduction=reduction, weight_decay=weight_decay, **kwargs)

    def compute(self, s: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def update(self, **kwargs) -> None:
        """topology.py:112-139: apply the learning rule to the current layer state."""
        if kwargs.get("learning", True):
            self.update_rule.update(**kwargs)
        mask = kwargs.get("mask", None)
        if mask is not None:                                              # topology.py:127-131
            self.w.masked_fill_(mask.to(self.w.device).bool(), 0)

    def reset_state_variables(self) -> None:
        pass

    @staticmethod
    def cast_dtype_if_needed(w, w_dtype):
        if w.dtype != w_dtype:
            warnings.warn(f"Provided w has data type {w.dtype} but parameter w_dtype is {w_dtype}")
            return w.to(dtype=w_dtype)
        return w


class Connection(AbstractConnection):
    """Dense all-to-all synapses (reference: topology.py:265-399)."""

    def __init__(
        self,
        source: Nodes,
        target: Nodes,
        nu: Optional[Union[float, Sequence[float], Sequence[torch.Tensor]]] = None,
        reduction: Optional[callable] = None,
        weight_decay: float = 0.0,
        w_dtype: torch.dtype = torch.float32,
        **kwargs,
    ) -> None:
        if w_dtype != torch.float32:
            raise NotImplementedError("bindsnet_b200 computes in float32 only (SURVEY.md §8b)")
        super().__init__(source, target, nu, reduction, weight_decay, **kwargs)
        w = kwargs.get("w", None)
        if w is None:
            # topology.py:308-313
            if (self.wmin == -np.inf).any() or (self.wmax == np.inf).any():
                w = torch.clamp(torch.rand(source.n, target.n), self.wmin, self.wmax)
            else:
                w = self.wmin + torch.rand(source.n, target.n) * (self.wmax - self.wmin)
            w = w.to(dtype=w_dtype)
        else:
            # topology.py:314-317
            if (self.wmin != -np.inf).any() or (self.wmax != np.inf).any():
                w = torch.clamp(torch.as_tensor(w), self.wmin, self.wmax)
            w = self.cast_dtype_if_needed(torch.as_tensor(w), w_dtype)
        self.w = Parameter(w.detach().clone().contiguous(), requires_grad=False)
        b = kwargs.get("b", None)
        self.b = Parameter(torch.as_tensor(b, dtype=torch.float32), requires_grad=False) if b is not None else None

    def compute(self, s: torch.Tensor) -> torch.Tensor:
        """``s.float() @ w (+ b)`` through the CUDA spike-gather (reference: topology.py:332-346)."""
        from . import _plan

        return _plan.compute_single_connection(self, s)

    def normalize(self) -> None:
        """topology.py:383-392."""
        if self.norm is not None:
            from . import _plan

            _plan.normalize_single_connection(self)

    # -- plan export -----------------------------------------------------------------------
    def _fill_desc(self, d: "_abi.SnnConn", dt: float, rule: bool = True) -> None:
        d.kind = _abi.SNN_CONN_DENSE
        if self.wmin.numel() != 1 or self.wmax.numel() != 1:
            raise NotImplementedError("per-synapse wmin/wmax tensors are not supported by the CUDA core yet")
        d.wmin = _scalar(self.wmin, "wmin")
        d.wmax = _scalar(self.wmax, "wmax")
        d.has_norm = int(self.norm is not None)
        d.norm_abs = 1
        d.norm = float(self.norm) if self.norm is not None else 0.0
        d.dt_scale = 1.0
        if rule:
            self.update_rule._fill_desc(d)


def _pair(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class Conv2dConnection(AbstractConnection):
    """2-D convolutional synapses (reference: topology.py:686-844).  ``w`` is
    ``[out_channels, in_channels, kh, kw]``, ``b`` ``[out_channels]`` (zeros by default, :795-797);
    source / target populations are ``[C, H, W]`` shaped.  Inside ``Network.run`` the convolution is a
    spike-gather over each target neuron's receptive field; ``MSTDP`` is the learning rule the CUDA
    core fuses for it (learning.py:1942-2015)."""

    def __init__(
        self,
        source: Nodes,
        target: Nodes,
        kernel_size,
        stride=1,
        padding=0,
        dilation=1,
        nu: Optional[Union[float, Sequence[float], Sequence[torch.Tensor]]] = None,
        reduction: Optional[callable] = None,
        weight_decay: float = 0.0,
        w_dtype: torch.dtype = torch.float32,
        **kwargs,
    ) -> None:
        if w_dtype != torch.float32:
            raise NotImplementedError("bindsnet_b200 computes in float32 only (SURVEY.md §8b)")
        # geometry first: the learning rule built by the base constructor looks at it
        self_kernel, self_stride = _pair(kernel_size), _pair(stride)
        self_padding, self_dilation = _pair(padding), _pair(dilation)
        assert len(source.shape) == 3 and len(target.shape) == 3, "Conv2dConnection needs [C, H, W] populations"
        in_channels, input_height, input_width = source.shape
        out_channels = target.shape[0]
        # topology.py:752-772 (the reference swaps the names width / height; the values are these)
        out_h = int((input_height - self_kernel[0] + 2 * self_padding[0]) / self_stride[0] + 1)
        out_w = int((input_width - self_kernel[1] + 2 * self_padding[1]) / self_stride[1] + 1)
        assert target.shape[1] == out_h and target.shape[2] == out_w, (
            "Target dimensionality must be (out_channels, ?,"
            "(input_height - filter_height + 2 * padding_height) / stride_height + 1,"
            "(input_width - filter_width + 2 * padding_width) / stride_width + 1"
        )
        object.__setattr__(self, "_geometry", (self_kernel, self_stride, self_padding, self_dilation))
        super().__init__(source, target, nu, reduction, weight_decay, **kwargs)
        self.kernel_size, self.stride, self.padding, self.dilation = self_kernel, self_stride, self_padding, self_dilation
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        w = kwargs.get("w", None)
        shape = (self.out_channels, self.in_channels, *self.kernel_size)
        if w is None:
            # topology.py:775-786
            if (self.wmin == -np.inf).any() or (self.wmax == np.inf).any():
                w = torch.clamp(torch.rand(*shape), self.wmin, self.wmax)
            else:
                w = (self.wmax - self.wmin) * torch.rand(*shape)
                w = w + self.wmin
        else:
            # topology.py:787-790
            w = torch.as_tensor(w)
            if (self.wmin == -np.inf).any() or (self.wmax == np.inf).any():
                w = torch.clamp(w, self.wmin, self.wmax)
            w = self.cast_dtype_if_needed(w, w_dtype)
        assert tuple(w.shape) == shape, f"w must have shape {shape}"
        self.w = Parameter(w.detach().clone().float().contiguous(), requires_grad=False)
        self.b = Parameter(torch.as_tensor(kwargs.get("b", torch.zeros(self.out_channels)), dtype=torch.float32).clone(),
                           requires_grad=False)

    def compute(self, s: torch.Tensor) -> torch.Tensor:
        """``F.conv2d(s.float(), w, b, stride, padding, dilation)`` for {0,1} spikes (topology.py:799-815), as the
        spike-gather the window kernels use (``snn_b200_conn_compute``)."""
        from . import _plan

        return _plan.compute_single_connection(self, s)

    def normalize(self) -> None:
        """Every (out, in) filter scaled to sum ``norm`` (topology.py:824-837); also runs at the end of every
        ``Network.run`` window."""
        if self.norm is not None:
            from . import _plan

            _plan.normalize_single_connection(self)

    def _fill_desc(self, d: "_abi.SnnConn", dt: float, rule: bool = True) -> None:
        d.kind = _abi.SNN_CONN_CONV2D
        if self.wmin.numel() != 1 or self.wmax.numel() != 1:
            raise NotImplementedError("per-synapse wmin/wmax tensors are not supported by the CUDA core yet")
        d.wmin = _scalar(self.wmin, "wmin")
        d.wmax = _scalar(self.wmax, "wmax")
        d.has_norm = int(self.norm is not None)
        d.norm_abs = 0
        d.norm = float(self.norm) if self.norm is not None else 0.0
        d.dt_scale = 1.0
        d.cin, d.hin, d.win = (int(v) for v in self.source.shape)
        d.cout, d.hout, d.wout = (int(v) for v in self.target.shape)
        d.kh, d.kw = self.kernel_size
        d.sh, d.sw = self.stride
        d.ph, d.pw = self.padding
        d.dh, d.dw = self.dilation
        if rule:
            self.update_rule._fill_desc(d)


class AbstractMulticompartmentConnection(ABC, Module):
    """Reference: topology.py:159-262."""

    def __init__(self, source: Nodes, target: Nodes, device, pipeline: list = None, **kwargs) -> None:
        super().__init__()
        assert isinstance(source, Nodes), "Source is not a Nodes object"
        assert isinstance(target, Nodes), "Target is not a Nodes object"
        self.source = source
        self.target = target
        self.device = device
        self.pipeline = [] if pipeline is None else pipeline
        self.feature_index = {}
        for feature in self.pipeline:
            self.feature_index[feature.name] = feature
            feature.prime_feature(connection=self, device=self.device, **kwargs)

    def append_pipeline(self, feature) -> None:
        self.pipeline.append(feature)
        feature.prime_feature(connection=self, device=self.device)
        self.feature_index[feature.name] = feature


class MulticompartmentConnection(AbstractMulticompartmentConnection):
    """Feature-pipeline connection (reference: topology.py:402-537).  The CUDA core executes
    pipelines consisting of exactly one dense ``Weight`` feature — what every model in
    ``bindsnet.models`` builds (models.py:185-236)."""

    def __init__(
        self,
        source: Nodes,
        target: Nodes,
        device="cpu",
        pipeline: list = None,
        manual_update: bool = False,
        traces: bool = False,
        **kwargs,
    ) -> None:
        super().__init__(source, target, device, pipeline if pipeline is not None else [], **kwargs)
        self.traces = traces
        self.manual_update = manual_update
        if self.traces:
            raise NotImplementedError("MulticompartmentConnection(traces=True) is not implemented by the CUDA core")

    def _weight(self):
        from .topology_features import Weight

        if len(self.pipeline) != 1 or not isinstance(self.pipeline[0], Weight):
            raise NotImplementedError(
                "the CUDA core executes MulticompartmentConnection pipelines made of a single Weight feature"
            )
        return self.pipeline[0]

    @property
    def w(self) -> torch.Tensor:
        return self._weight().value

    def compute(self, s: torch.Tensor) -> torch.Tensor:
        """Reference: topology.py:437-479 + Weight.compute topology_features.py:633-645."""
        from . import _plan

        return _plan.compute_single_connection(self, s)

    def update(self, **kwargs) -> None:
        """topology.py:509-518."""
        if kwargs.get("learning", False) and not self.manual_update:
            for f in self.pipeline:
                f.update(**kwargs)

    def normalize(self) -> None:
        """topology.py:520-527."""
        for f in self.pipeline:
            f.normalize()

    def reset_state_variables(self) -> None:
        for f in self.pipeline:
            f.reset_state_variables()

    def _apply(self, fn, *args, **kwargs):
        # Features are not nn.Modules in the reference (they take an explicit device,
        # topology.py:169,192); here Network.to(device) carries their value along.
        out = super()._apply(fn, *args, **kwargs)
        for f in self.pipeline:
            f._apply(fn)
        return out

    def _fill_desc(self, d: "_abi.SnnConn", dt: float, rule: bool = True) -> None:
        d.kind = _abi.SNN_CONN_MCC
        self._weight()._fill_desc(d, dt, self.manual_update)


class LocalConnection(Connection):
    """Locally connected synapses (reference: topology.py:1304-1484): a dense ``[source.n, target.n]`` matrix that is
    non-zero only inside each target neuron's receptive field.  The reference keeps it dense too — ``compute`` is the
    plain matrix product (:1441-1455) — and holds the structure with a mask of the initially-zero weights that
    ``update`` passes on when the caller gives none (:1457-1469), so here it IS a dense ``Connection`` whose plan always
    carries that mask; ``normalize`` divides by the plain column sum (:1471-1479) and ``norm`` is scaled by the kernel
    size (:1437-1438).  Target neuron ``f * conv_prod + c`` is filter ``f`` at receptive field ``c``."""

    def __init__(
        self,
        source: Nodes,
        target: Nodes,
        kernel_size,
        stride,
        n_filters: int,
        nu: Optional[Union[float, Sequence[float], Sequence[torch.Tensor]]] = None,
        reduction: Optional[callable] = None,
        weight_decay: float = 0.0,
        w_dtype: torch.dtype = torch.float32,
        **kwargs,
    ) -> None:
        kernel_size, stride = _pair(kernel_size), _pair(stride)
        shape = kwargs.get("input_shape", None)
        if shape is None:
            shape = _pair(int(np.sqrt(source.n)))                                  # topology.py:1370-1373
        if tuple(kernel_size) == tuple(shape):
            conv_size = (1, 1)
        else:
            conv_size = (int((shape[0] - kernel_size[0]) / stride[0]) + 1, int((shape[1] - kernel_size[1]) / stride[1]) + 1)
        conv_prod, kernel_prod = int(np.prod(conv_size)), int(np.prod(kernel_size))
        assert target.n == n_filters * conv_prod, f"Total neurons in target layer must be {n_filters * conv_prod}. Got {target.n}."
        # topology.py:1393-1407 — the index arithmetic is the reference's (including `k1 * shape[0]`)
        c1, c2, k1, k2 = np.meshgrid(np.arange(conv_size[0]), np.arange(conv_size[1]), np.arange(kernel_size[0]),
                                     np.arange(kernel_size[1]), indexing="ij")
        loc = c1 * stride[0] * shape[1] + c2 * stride[1] + k1 * shape[0] + k2      # [c1, c2, k1, k2]
        locations = torch.from_numpy(loc.transpose(2, 3, 0, 1).reshape(kernel_prod, conv_prod).astype(np.int64))
        w = kwargs.get("w", None)
        if w is None:
            # topology.py:1410-1423: random weights inside the receptive fields only
            lo, hi = kwargs.get("wmin", -np.inf), kwargs.get("wmax", np.inf)
            w = torch.zeros(source.n, target.n)
            cols = (torch.arange(n_filters).view(-1, 1, 1) * conv_prod + torch.arange(conv_prod).view(1, 1, -1)).expand(n_filters, kernel_prod, conv_prod)
            rows = locations.view(1, kernel_prod, conv_prod).expand(n_filters, kernel_prod, conv_prod)
            w[rows.reshape(-1), cols.reshape(-1)] = torch.rand(n_filters * kernel_prod * conv_prod)
            if np.isinf(lo) or np.isinf(hi):
                w = torch.clamp(w, lo, hi)
            else:
                w = lo + w * (hi - lo)
            kwargs = dict(kwargs, w=w)
        kwargs.setdefault("b", torch.zeros(target.n))                              # topology.py:1433
        super().__init__(source, target, nu=nu, reduction=reduction, weight_decay=weight_decay, w_dtype=w_dtype, **kwargs)
        self.kernel_size, self.stride, self.n_filters, self.conv_size = kernel_size, stride, n_filters, conv_size
        self.register_buffer("locations", locations)
        self.register_buffer("mask", self.w == 0)                                  # topology.py:1431
        if self.norm is not None:
            self.norm = self.norm * kernel_prod                                    # topology.py:1437-1438

    def update(self, **kwargs) -> None:
        """topology.py:1457-1469."""
        if kwargs.get("mask", None) is None:
            kwargs["mask"] = self.mask
        super().update(**kwargs)

    def _fill_desc(self, d: "_abi.SnnConn", dt: float, rule: bool = True) -> None:
        super()._fill_desc(d, dt, rule)
        d.norm_abs = 0   # w *= norm / w.sum(0)  (topology.py:1476-1479)


def _unsupported(name: str, where: str):
    class _Unsupported:
        __doc__ = f"``{name}`` (reference: {where}) — not on the accelerated path (SURVEY.md §8f)."

        def __init__(self, *args, **kwargs):
            raise NotImplementedError(
                f"{name} is outside the hot path bindsnet_b200 implements (Connection, "
                "MulticompartmentConnection[Weight]); see DESIGN.md 'Out of scope'"
            )

    _Unsupported.__name__ = name
    return _Unsupported


Conv1dConnection = _unsupported("Conv1dConnection", "topology.py:540-683")
Conv3dConnection = _unsupported("Conv3dConnection", "topology.py:847-1025")
MaxPool1dConnection = _unsupported("MaxPool1dConnection", "topology.py:1028-1121")
MaxPool2dConnection = _unsupported("MaxPool2dConnection", "topology.py:1124-1211")
MaxPoo3dConnection = _unsupported("MaxPoo3dConnection", "topology.py:1214-1301")
LocalConnection1D = _unsupported("LocalConnection1D", "topology.py:1487-1620")
LocalConnection2D = _unsupported("LocalConnection2D", "topology.py:1623-1767")
LocalConnection3D = _unsupported("LocalConnection3D", "topology.py:1770-1917")
MeanFieldConnection = _unsupported("MeanFieldConnection", "topology.py:1920-2006")
SparseConnection = _unsupported("SparseConnection", "topology.py:2009-2017")
