"""Connection features — host-side mirror of ``bindsnet/network/topology_features.py``.

Only ``Weight`` (reference: topology_features.py:575-671, base class :15-362) is on the
accelerated path: it is the single feature every ``bindsnet.models`` network puts in a
``MulticompartmentConnection`` pipeline (models.py:185-236).
"""
from __future__ import annotations

import warnings
from abc import ABC
from typing import Optional, Sequence, Union

import torch
from torch.nn import Parameter

from .. import _abi


class AbstractFeature(ABC):
    """Reference: topology_features.py:15-362."""

    def __init__(
        self,
        name: str,
        value: Union[torch.Tensor, float, int] = None,
        value_dtype: torch.dtype = torch.float32,
        range: Optional[Union[list, tuple]] = None,
        clamp_frequency: Optional[int] = 1,
        norm: Optional[Union[torch.Tensor, float, int]] = None,
        learning_rule=None,
        nu: Optional[Union[list, tuple, int, float]] = None,
        reduction: Optional[callable] = None,
        enforce_polarity: Optional[bool] = False,
        decay: float = 0.0,
        parent_feature=None,
        sparse: Optional[bool] = False,
        batch_size: int = 1,
        **kwargs,
    ) -> None:
        from ..learning.MCC_learning import MSTDP, MSTDPET, NoOp, PostPre

        assert isinstance(name, str), f"Feature {name}'s name should be of type str"
        assert value is None or isinstance(value, (torch.Tensor, float, int)), (
            f"Feature {name} should be of type float, int, or torch.Tensor, not {type(value)}"
        )
        assert norm is None or isinstance(norm, (torch.Tensor, float, int)), (
            f"Feature {name}'s norm should be of type float, int, or torch.Tensor, not {type(norm)}"
        )
        assert learning_rule is None or learning_rule in (NoOp, PostPre, MSTDP, MSTDPET), (
            f"Feature {name}'s learning_rule should be an MCC learning rule, not {learning_rule}"
        )
        assert nu is None or isinstance(nu, (list, tuple)), (
            f"Feature {name}'s nu should be of type list or tuple, not {type(nu)}"
        )
        assert decay is None or isinstance(decay, float), f"Feature {name}'s decay should be of type float"
        if sparse:
            raise NotImplementedError("sparse feature values are not implemented by the CUDA core")
        if parent_feature is not None:
            raise NotImplementedError("feature linking (parent_feature) is not implemented by the CUDA core")
        if value_dtype != torch.float32:
            raise NotImplementedError("bindsnet_b200 computes in float32 only (SURVEY.md §8b)")

        self.name = name
        self.value = value
        self.range = [-1.0, 1.0] if range is None else range
        self.clamp_frequency = clamp_frequency
        self.norm = norm
        self.learning_rule = learning_rule
        self.nu = nu
        self.reduction = reduction
        self.decay = decay
        self.parent_feature = parent_feature
        self.sparse = sparse
        self.batch_size = batch_size
        self.kwargs = kwargs
        self.is_primed = False

        # topology_features.py:310-328
        r = self.range
        assert isinstance(r, (list, tuple)) and len(r) == 2, f"Invalid range for feature {name}"
        assert r[0] < r[1], f"Invalid range for feature {name}: the min value is larger than the max value"
        if value is None:
            return
        if isinstance(value, torch.Tensor):
            # topology_features.py:330-351
            assert (value >= r[0]).all() and (value <= r[1]).all(), (
                f"Feature out of range for {name}: Features values not in [{r[0]}, {r[1]}]"
            )
            if value.dtype != value_dtype:
                warnings.warn(f"Provided value has data type {value.dtype} but parameter w_dtype is {value_dtype}")
                self.value = value.to(dtype=value_dtype)

    def initialize_value(self):
        raise NotImplementedError

    def prime_feature(self, connection, device, **kwargs) -> None:
        """topology_features.py:173-240: wrap the value, move it to ``device`` and
        instantiate the learning rule."""
        from ..learning.MCC_learning import NoOp

        if self.is_primed:
            return
        self.is_primed = True
        if isinstance(self.value, torch.Tensor):
            assert tuple(self.value.shape) == (connection.source.n, connection.target.n)
        if self.norm is not None and isinstance(self.norm, torch.Tensor):
            assert self.norm.shape[0] == connection.target.n
        if self.value is None:
            self.value = self.initialize_value()
        if isinstance(self.value, (int, float)):
            self.value = torch.Tensor([self.value])
        self.value = Parameter(self.value.detach().clone().contiguous(), requires_grad=False).to(device)
        rule_cls = self.learning_rule or NoOp
        self.learning_rule = rule_cls(
            connection=connection, feature_value=self.value, range=self.range, nu=self.nu,
            reduction=self.reduction, decay=self.decay, **kwargs,
        )
        del self.nu, self.reduction, self.decay, self.range

    def update(self, **kwargs) -> None:
        """topology_features.py:242-248."""
        self.learning_rule.update(**kwargs)

    def normalize(self) -> None:
        """topology_features.py:250-266 (plain, not absolute, column sums)."""
        if self.norm is not None:
            from . import _plan

            _plan.normalize_feature(self)

    def reset_state_variables(self) -> None:
        pass

    def _apply(self, fn) -> None:
        if isinstance(self.value, torch.Tensor):
            moved = fn(self.value)
            if moved is not self.value:
                self.value = moved
                if getattr(self.learning_rule, "feature_value", None) is not None:
                    self.learning_rule.feature_value = moved


class Weight(AbstractFeature):
    """Per-synapse multiplicative weight (reference: topology_features.py:575-671)."""

    def __init__(
        self,
        name: str,
        value: Union[torch.Tensor, float, int] = None,
        value_dtype: torch.dtype = torch.float32,
        range: Optional[Sequence[float]] = None,
        norm: Optional[Union[torch.Tensor, float, int]] = None,
        norm_frequency: Optional[str] = "sample",
        learning_rule=None,
        nu: Optional[Union[list, tuple]] = None,
        reduction: Optional[callable] = None,
        enforce_polarity: Optional[bool] = False,
        decay: float = 0.0,
        sparse: Optional[bool] = False,
        batch_size: int = 1,
    ) -> None:
        if norm_frequency != "sample":
            raise NotImplementedError("Weight(norm_frequency='time step') is not implemented by the CUDA core")
        if enforce_polarity:
            raise NotImplementedError("Weight(enforce_polarity=True) is not implemented by the CUDA core")
        self.norm_frequency = norm_frequency
        self.enforce_polarity = enforce_polarity
        super().__init__(
            name=name, value=value, value_dtype=value_dtype,
            range=[-torch.inf, +torch.inf] if range is None else range,
            norm=norm, learning_rule=learning_rule, nu=nu, reduction=reduction,
            decay=decay, sparse=sparse, batch_size=batch_size,
        )

    def prime_feature(self, connection, device, **kwargs) -> None:
        """topology_features.py:647-661."""
        if self.value is None:
            self.initialize_value = lambda: torch.rand(connection.source.n, connection.target.n)
        super().prime_feature(connection, device, enforce_polarity=self.enforce_polarity, **kwargs)

    def _fill_desc(self, d: "_abi.SnnConn", dt: float, manual_update: bool) -> None:
        if self.value.dim() != 2:
            raise NotImplementedError("scalar Weight values are not supported by the CUDA core")
        if isinstance(self.norm, torch.Tensor):
            raise NotImplementedError("per-target tensor norms are not supported by the CUDA core yet")
        d.has_norm = int(self.norm is not None)
        d.norm_abs = 0
        d.norm = float(self.norm) if self.norm is not None else 0.0
        d.dt_scale = float(dt)
        self.learning_rule._fill_desc(d)
        if manual_update:
            d.rule = _abi.SNN_RULE_NONE


def _unsupported(name: str, where: str):
    class _Unsupported:
        __doc__ = f"``{name}`` (reference: {where}) — not on the accelerated path."

        def __init__(self, *args, **kwargs):
            raise NotImplementedError(
                f"the {name} feature is outside the hot path bindsnet_b200 implements (Weight only)"
            )

    _Unsupported.__name__ = name
    return _Unsupported


Probability = _unsupported("Probability", "topology_features.py:365-464")
Mask = _unsupported("Mask", "topology_features.py:467-549")
MeanField = _unsupported("MeanField", "topology_features.py:552-572")
Bias = _unsupported("Bias", "topology_features.py:674-721")
Intensity = _unsupported("Intensity", "topology_features.py:724-769")
Degradation = _unsupported("Degradation", "topology_features.py:772-813")
