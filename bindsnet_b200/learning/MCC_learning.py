"""Learning rules for ``MulticompartmentConnection`` features — host-side mirror of
``bindsnet/learning/MCC_learning.py`` (``MCC_LearningRule`` :16-118, ``NoOp`` :121-146,
``PostPre`` :149-302).  The rule objects hold hyper-parameters; the update itself is fused
into the CUDA window kernels."""
from __future__ import annotations

import warnings
from abc import ABC
from typing import Optional, Sequence, Union

import torch

from .. import _abi


def _reduction_code(reduction, batch_size) -> int:
    if reduction is None:
        # learning.py:76-80 / MCC_learning.py:73-79: squeeze when batch_size == 1 at rule
        # construction, else sum.  Over a batch of one they coincide; over a larger batch
        # the reference's squeeze raises — we check that at plan time.
        return _abi.SNN_REDUCE_SUM
    if reduction in (torch.sum,):
        return _abi.SNN_REDUCE_SUM
    if reduction in (torch.mean,):
        return _abi.SNN_REDUCE_MEAN
    if reduction is torch.squeeze:
        return _abi.SNN_REDUCE_SUM
    raise NotImplementedError(
        f"reduction {reduction!r} is not supported by the CUDA core (torch.sum, torch.mean, torch.squeeze)"
    )


class MCC_LearningRule(ABC):
    """Reference: MCC_learning.py:16-118."""

    rule_code = _abi.SNN_RULE_NONE

    def __init__(
        self,
        connection,
        feature_value: Union[float, int, torch.Tensor],
        range: Optional[Union[list, tuple]] = None,
        nu: Optional[Union[float, Sequence[float]]] = None,
        reduction: Optional[callable] = None,
        decay: float = 0.0,
        enforce_polarity: bool = False,
        **kwargs,
    ) -> None:
        self.connection = connection
        self.source = connection.source
        self.target = connection.target
        self.feature_value = feature_value
        self.enforce_polarity = enforce_polarity
        self.min, self.max = range
        if nu is None:
            nu = [0.2, 0.1]
        elif isinstance(nu, (float, int)):
            nu = [nu, nu]
        self.nu = torch.zeros(2, dtype=torch.float)
        self.nu[0] = nu[0]
        self.nu[1] = nu[1]
        if (self.nu == torch.zeros(2)).all() and not isinstance(self, NoOp):
            warnings.warn(
                f"nu is set to [0., 0.] for {type(self).__name__} learning rule. "
                "It will disable the learning process."
            )
        self._squeeze = reduction is None and self.source.batch_size == 1 or reduction is torch.squeeze
        self.reduction = reduction if reduction is not None else (
            torch.squeeze if self.source.batch_size == 1 else torch.sum
        )
        self._reduction_code = _reduction_code(reduction, self.source.batch_size)
        self.decay = 1.0 - decay if decay else 1.0

    def update(self, **kwargs) -> None:
        from ..network import _plan

        _plan.update_single_connection(self.connection)

    def reset_state_variables(self) -> None:
        pass

    def _fill_desc(self, d: "_abi.SnnConn") -> None:
        d.rule = self.rule_code
        d.reduction = self._reduction_code
        d.nu0 = float(self.nu[0])
        d.nu1 = float(self.nu[1])
        d.weight_decay = float(self.decay)
        d.wmin = float(self.min) if self.min is not None else float("-inf")
        d.wmax = float(self.max) if self.max is not None else float("inf")
        d.has_clamp = int(self.min is not None or self.max is not None)


class NoOp(MCC_LearningRule):
    """Reference: MCC_learning.py:121-146 — really does nothing (no decay, no clamp)."""

    rule_code = _abi.SNN_RULE_NONE

    def __init__(self, **args) -> None:
        pass

    def update(self, **kwargs) -> None:
        pass

    def _fill_desc(self, d) -> None:
        d.rule = _abi.SNN_RULE_NONE
        d.reduction = _abi.SNN_REDUCE_SUM
        d.weight_decay = 1.0
        d.wmin, d.wmax = float("-inf"), float("inf")
        d.has_clamp = 0


class PostPre(MCC_LearningRule):
    """Pair-based STDP on a ``Weight`` feature (reference: MCC_learning.py:149-302): both
    terms are scaled by ``connection.dt`` (:262,:298), then decay and clamp (:86-110)."""

    rule_code = _abi.SNN_RULE_MCC_POSTPRE

    def __init__(
        self,
        connection,
        feature_value: Union[torch.Tensor, float, int],
        range: Optional[Sequence[float]] = None,
        nu: Optional[Union[float, Sequence[float]]] = None,
        reduction: Optional[callable] = None,
        decay: float = 0.0,
        enforce_polarity: bool = False,
        **kwargs,
    ) -> None:
        super().__init__(
            connection=connection, feature_value=feature_value,
            range=[-1, +1] if range is None else range, nu=nu, reduction=reduction,
            decay=decay, enforce_polarity=enforce_polarity, **kwargs,
        )
        assert self.source.traces and self.target.traces, (
            "Both pre- and post-synaptic nodes must record spike traces "
            "(use traces='True' on source/target layers)"
        )
        from ..network.topology import MulticompartmentConnection

        if not isinstance(connection, MulticompartmentConnection):
            raise NotImplementedError("This learning rule is not supported for this Connection type.")
        if enforce_polarity:
            raise NotImplementedError("enforce_polarity is not implemented by the CUDA core")
        if kwargs.get("average_update", 0):
            raise NotImplementedError("PostPre(average_update>0) is not implemented by the CUDA core")


def _unsupported(name: str, where: str):
    class _Unsupported(MCC_LearningRule):
        __doc__ = f"``{name}`` (reference: {where}) — not on the accelerated path (SURVEY.md §8f)."

        def __init__(self, *args, **kwargs):
            raise NotImplementedError(f"MCC_learning.{name} is outside the hot path bindsnet_b200 implements")

    _Unsupported.__name__ = name
    return _Unsupported


MSTDP = _unsupported("MSTDP", "MCC_learning.py:392-551")
MSTDPET = _unsupported("MSTDPET", "MCC_learning.py:554-738")
