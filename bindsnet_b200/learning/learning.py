"""Learning rules for ``Connection`` — host-side mirror of ``bindsnet/learning/learning.py``
(``LearningRule`` :25-104, ``NoOp`` :107-146, ``PostPre`` :149-420,
``WeightDependentPostPre`` :562-653).  The rule objects hold hyper-parameters; the update
itself is fused into the CUDA window kernels (``Network.run``) or submitted for one step by
``rule.update()``."""
from __future__ import annotations

import warnings
from abc import ABC
from typing import Optional, Sequence, Union

import numpy as np
import torch

from .. import _abi
from .MCC_learning import _reduction_code


class LearningRule(ABC):
    """Reference: learning.py:25-104."""

    rule_code = None

    def __init__(
        self,
        connection,
        nu: Optional[Union[float, Sequence[float], Sequence[torch.Tensor]]] = None,
        reduction: Optional[callable] = None,
        weight_decay: float = 0.0,
        **kwargs,
    ) -> None:
        self.connection = connection
        self.source = connection.source
        self.target = connection.target
        self.wmin = connection.wmin
        self.wmax = connection.wmax
        # learning.py:58-66
        if nu is None:
            self.nu = torch.tensor([0.0, 0.0], dtype=torch.float)
        elif isinstance(nu, (float, int)):
            self.nu = torch.tensor([nu, nu], dtype=torch.float)
        elif all(isinstance(e, (float, int)) for e in nu):
            self.nu = torch.tensor(nu, dtype=torch.float)
        else:
            raise NotImplementedError("per-synapse learning-rate tensors are not supported by the CUDA core yet")
        if not self.nu.any() and not isinstance(self, NoOp):
            warnings.warn(
                f"nu is set to zeros for {type(self).__name__} learning rule. "
                "It will disable the learning process."
            )
        self.reduction = reduction if reduction is not None else (
            torch.squeeze if self.source.batch_size == 1 else torch.sum
        )
        self._squeeze = self.reduction is torch.squeeze
        self._reduction_code = _reduction_code(reduction, self.source.batch_size)
        self.weight_decay = 1.0 - weight_decay if weight_decay else 1.0

    def update(self, **kwargs) -> None:
        """Apply this rule once to the connection from the layers' current ``s``/``x``
        (reference: the rule-specific ``_connection_update`` + learning.py:87-104)."""
        from ..network import _plan

        _plan.update_single_connection(self.connection)

    def _fill_desc(self, d: "_abi.SnnConn") -> None:
        if self.rule_code is None:
            raise NotImplementedError(
                f"user-defined learning rule {type(self).__name__} cannot be fused into the CUDA window; "
                "supported: NoOp, PostPre, WeightDependentPostPre"
            )
        d.rule = self.rule_code
        d.reduction = self._reduction_code
        d.nu0 = float(self.nu[0])
        d.nu1 = float(self.nu[1])
        d.weight_decay = float(self.weight_decay)
        # learning.py:97-104: clamp iff a bound is finite and the rule is not NoOp
        from ..network.nodes import _scalar

        finite = _scalar(self.connection.wmin, "wmin") != -np.inf or _scalar(self.connection.wmax, "wmax") != np.inf
        d.has_clamp = int(finite and not isinstance(self, NoOp))


class NoOp(LearningRule):
    """Reference: learning.py:107-146 — weight decay only."""

    rule_code = _abi.SNN_RULE_NOOP


class PostPre(LearningRule):
    """Pair-based STDP (reference: learning.py:149-420; dense update :390-420)."""

    rule_code = _abi.SNN_RULE_POSTPRE

    def __init__(self, connection, nu=None, reduction=None, weight_decay: float = 0.0, **kwargs) -> None:
        super().__init__(connection=connection, nu=nu, reduction=reduction, weight_decay=weight_decay, **kwargs)
        assert self.source.traces and self.target.traces, (
            "Both pre- and post-synaptic nodes must record spike traces."
        )
        from ..network.topology import Connection

        if not isinstance(connection, Connection):
            raise NotImplementedError("This learning rule is not supported for this Connection type.")


class WeightDependentPostPre(LearningRule):
    """Weight-dependent STDP (reference: learning.py:562-653; dense update :626-653)."""

    rule_code = _abi.SNN_RULE_WDEP_POSTPRE

    def __init__(self, connection, nu=None, reduction=None, weight_decay: float = 0.0, **kwargs) -> None:
        super().__init__(connection=connection, nu=nu, reduction=reduction, weight_decay=weight_decay, **kwargs)
        assert self.source.traces, "Pre-synaptic nodes must record spike traces."
        assert self.target.traces, "Post-synaptic nodes must record spike traces."
        assert (connection.wmin != -np.inf).any() and (connection.wmax != np.inf).any(), (
            "Connection must define finite wmin and wmax."
        )
        from ..network.topology import Connection

        if not isinstance(connection, Connection):
            raise NotImplementedError("This learning rule is not supported for this Connection type.")


def _unsupported(name: str, where: str):
    class _Unsupported(LearningRule):
        __doc__ = f"``{name}`` (reference: {where}) — not on the accelerated path (SURVEY.md §8f)."

        def __init__(self, *args, **kwargs):
            raise NotImplementedError(f"learning.{name} is outside the hot path bindsnet_b200 implements")

    _Unsupported.__name__ = name
    return _Unsupported


Hebbian = _unsupported("Hebbian", "learning.py:1052-1438")
MSTDP = _unsupported("MSTDP", "learning.py:1441-2121")
MSTDPET = _unsupported("MSTDPET", "learning.py:2124-2855")
Rmax = _unsupported("Rmax", "learning.py:2858-2960")
