"""Learning rules for ``Connection`` — host-side mirror of ``bindsnet/learning/learning.py``
(``LearningRule`` :25-104, ``NoOp`` :107-146, ``PostPre`` :149-420 / :457-497 (conv2d),
``WeightDependentPostPre`` :562-653 / :920-975, ``Hebbian`` :1052-1136 / :1348-1380, ``MSTDP``).  The rule objects hold hyper-parameters; the update
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
        (reference: the rule-specific ``_connection_update`` + learning.py:87-104).

        A USER-DEFINED rule (a subclass that changes ``self.connection.w`` itself with torch ops and then calls
        ``super().update()``, learning.py:31-104) gets the reference's base behaviour here: weight decay and the
        clamp to ``[wmin, wmax]``; ``Network.run`` drives networks with such rules step by step (the scripted tier)."""
        if self.rule_code is None:
            w = self.connection.w
            if self.weight_decay:                                         # learning.py:93-94
                w *= self.weight_decay
            wmin, wmax = self.connection.wmin, self.connection.wmax       # learning.py:97-104
            if bool((wmin != -np.inf).any()) or bool((wmax != np.inf).any()):
                w.clamp_(float(wmin), float(wmax))
            return
        from ..network import _plan

        _plan.update_single_connection(self.connection)

    def _fill_desc(self, d: "_abi.SnnConn") -> None:
        if self.rule_code is None:
            raise NotImplementedError(
                f"user-defined learning rule {type(self).__name__} cannot be fused into the CUDA window; "
                "supported: NoOp, PostPre, WeightDependentPostPre, Hebbian, MSTDP"
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


def _check_connection(rule, connection) -> None:
    """The STDP-family rules exist for ``Connection`` (``_connection_update``) and ``Conv2dConnection``
    (``_conv2d_connection_update``); the reference's im2col ignores dilation, so a dilated filter is refused."""
    from ..network.topology import Connection, Conv2dConnection

    if not isinstance(connection, (Connection, Conv2dConnection)):
        raise NotImplementedError("This learning rule is not supported for this Connection type.")
    if isinstance(connection, Conv2dConnection) and connection._geometry[3] != (1, 1):
        raise NotImplementedError(f"{type(rule).__name__} on a dilated Conv2dConnection is undefined in the reference (im2col ignores dilation)")


class NoOp(LearningRule):
    """Reference: learning.py:107-146 — weight decay only."""

    rule_code = _abi.SNN_RULE_NOOP


class PostPre(LearningRule):
    """Pair-based STDP (reference: learning.py:149-420; dense update :390-420, conv2d :457-497)."""

    rule_code = _abi.SNN_RULE_POSTPRE

    def __init__(self, connection, nu=None, reduction=None, weight_decay: float = 0.0, **kwargs) -> None:
        super().__init__(connection=connection, nu=nu, reduction=reduction, weight_decay=weight_decay, **kwargs)
        assert self.source.traces and self.target.traces, (
            "Both pre- and post-synaptic nodes must record spike traces."
        )
        _check_connection(self, connection)


class WeightDependentPostPre(LearningRule):
    """Weight-dependent STDP (reference: learning.py:562-653; dense update :626-653, conv2d :920-975)."""

    rule_code = _abi.SNN_RULE_WDEP_POSTPRE

    def __init__(self, connection, nu=None, reduction=None, weight_decay: float = 0.0, **kwargs) -> None:
        super().__init__(connection=connection, nu=nu, reduction=reduction, weight_decay=weight_decay, **kwargs)
        assert self.source.traces, "Pre-synaptic nodes must record spike traces."
        assert self.target.traces, "Post-synaptic nodes must record spike traces."
        assert (connection.wmin != -np.inf).any() and (connection.wmax != np.inf).any(), (
            "Connection must define finite wmin and wmax."
        )
        _check_connection(self, connection)


class Hebbian(LearningRule):
    """Hebbian rule: both terms positive, learning rates applied after the batch reduction
    (reference: learning.py:1052-1438; dense update :1110-1136, conv2d :1348-1380)."""

    rule_code = _abi.SNN_RULE_HEBBIAN

    def __init__(self, connection, nu=None, reduction=None, weight_decay: float = 0.0, **kwargs) -> None:
        super().__init__(connection=connection, nu=nu, reduction=reduction, weight_decay=weight_decay, **kwargs)
        assert self.source.traces and self.target.traces, (
            "Both pre- and post-synaptic nodes must record spike traces."
        )
        _check_connection(self, connection)


class MSTDP(LearningRule):
    """Reward-modulated STDP (reference: learning.py:1440-2121): dense ``Connection``
    (``_connection_update`` :1504-1574) and ``Conv2dConnection`` (``_conv2d_connection_update``
    :1942-2015).  Rule state lives here like in the reference: ``p_plus``, ``p_minus`` and — for the
    convolutional form — ``eligibility`` ``[B, *w.shape]``.  The dense form never materialises the
    ``[B, n_src, n_tgt]`` eligibility: it is ``p_plus (x) s_post + s_pre (x) p_minus`` of the previous
    step, so the spikes the rule saw last are kept instead (``eligibility`` rebuilds it on request).
    ``Network.run(..., reward=r)`` is mandatory, ``a_plus`` / ``a_minus`` optional (:1540-1556).  For
    ``Conv2dConnection`` the eligibility is per sample for any batch size (the reference's final
    ``.view(w.size())`` at :2013 only works for batch size 1)."""

    rule_code = _abi.SNN_RULE_MSTDP

    def __init__(self, connection, nu=None, reduction=None, weight_decay: float = 0.0, **kwargs) -> None:
        super().__init__(connection=connection, nu=nu, reduction=reduction, weight_decay=weight_decay, **kwargs)
        from ..network.topology import Connection, Conv2dConnection

        if not isinstance(connection, (Connection, Conv2dConnection)):
            raise NotImplementedError("This learning rule is not supported for this Connection type.")
        self._conv = isinstance(connection, Conv2dConnection)
        if self._conv and connection._geometry[3] != (1, 1):
            raise NotImplementedError("MSTDP on a dilated Conv2dConnection is undefined in the reference (im2col ignores dilation)")
        self.tc_plus = torch.tensor(kwargs.get("tc_plus", 20.0))
        self.tc_minus = torch.tensor(kwargs.get("tc_minus", 20.0))
        self._run_kwargs = {}

    def update(self, **kwargs) -> None:
        raise NotImplementedError("MSTDP.update is fused into Network.run (it needs the run's reward); the standalone call is not exposed")

    def _prepare(self, B: int, dev: torch.device, run_kwargs: dict) -> None:
        """Allocate / validate the rule state for a window (learning.py:1519-1535, 1958-1961, 1979-1991)."""
        if run_kwargs.get("reward", None) is None:
            raise KeyError("reward")  # learning.py:1541: kwargs["reward"]
        for key in ("reward", "a_plus", "a_minus"):
            v = run_kwargs.get(key, None)
            if isinstance(v, dict) or (isinstance(v, torch.Tensor) and v.numel() != 1):
                raise NotImplementedError(f"run(..., {key}=...) must be a scalar for the CUDA core")
        self._run_kwargs = run_kwargs
        src, tgt = self.source, self.target

        def ensure(name, shape, dtype=torch.float32):
            t = getattr(self, name, None)
            if not isinstance(t, torch.Tensor) or tuple(t.shape) != tuple(shape) or t.device != dev or t.dtype != dtype:
                setattr(self, name, torch.zeros(*shape, dtype=dtype, device=dev))

        if self._conv:
            ensure("p_plus", (B, *src.shape))
            ensure("p_minus", (B, tgt.shape[0], tgt.shape[1] * tgt.shape[2]))
            ensure("_elig", (B, *self.connection.w.shape))
        else:
            ensure("p_plus", (B, src.n))
            ensure("p_minus", (B, tgt.n))
            ensure("_spre", (B, src.n), torch.uint8)
            ensure("_spost", (B, tgt.n), torch.uint8)

    @property
    def eligibility(self) -> torch.Tensor:
        """``[B, *w.shape]`` eligibility that the next update will apply (learning.py:1568-1572, 2005-2010)."""
        if self._conv:
            return self._elig
        return torch.bmm(self.p_plus.unsqueeze(2), self._spost.float().unsqueeze(1)) + torch.bmm(
            self._spre.float().unsqueeze(2), self.p_minus.unsqueeze(1))

    def _fill_desc(self, d: "_abi.SnnConn") -> None:
        super()._fill_desc(d)
        rk = self._run_kwargs
        dt = float(self.connection.dt)
        d.reward = float(rk["reward"])
        d.a_plus = float(rk["a_plus"]) if rk.get("a_plus", None) is not None else 1.0
        d.a_minus = float(rk["a_minus"]) if rk.get("a_minus", None) is not None else -1.0
        d.p_plus_decay = float(torch.exp(-dt / self.tc_plus))    # learning.py:1565 (fp32 tensor arithmetic)
        d.p_minus_decay = float(torch.exp(-dt / self.tc_minus))  # learning.py:1567
        d.p_plus, d.p_minus = self.p_plus.data_ptr(), self.p_minus.data_ptr()
        if self._conv:
            d.elig = self._elig.data_ptr()
        else:
            d.mst_spre, d.mst_spost = self._spre.data_ptr(), self._spost.data_ptr()


class MSTDPET(MSTDP):
    """Reward-modulated STDP with an eligibility trace on a dense ``Connection`` (reference: learning.py:2124-2249;
    ``_connection_update`` :2187-2249).  Batch size 1 only: the reference flattens the spikes of the whole batch into its
    ``[n]`` traces (:2214-2215), which only has a meaning for one sample.  Rule state like the reference's: ``p_plus [n_src]``,
    ``p_minus [n_tgt]``, ``eligibility_trace [n_src, n_tgt]``; ``eligibility`` is rebuilt on request from the traces and the
    spikes the rule saw last.  The convolutional / local forms (:2251-2855) are outside the implemented path."""

    rule_code = _abi.SNN_RULE_MSTDPET

    def __init__(self, connection, nu=None, reduction=None, weight_decay: float = 0.0, **kwargs) -> None:
        super().__init__(connection=connection, nu=nu, reduction=reduction, weight_decay=weight_decay, **kwargs)
        if self._conv:
            raise NotImplementedError("MSTDPET on a Conv2dConnection is outside the implemented path (dense Connection only)")
        self.tc_e_trace = torch.tensor(kwargs.get("tc_e_trace", 25.0))

    def _prepare(self, B: int, dev: torch.device, run_kwargs: dict) -> None:
        if B != 1:
            raise NotImplementedError("MSTDPET is defined for batch size 1 only (learning.py:2214-2215 flattens the batch)")
        super()._prepare(B, dev, run_kwargs)
        t = getattr(self, "eligibility_trace", None)
        shape = tuple(self.connection.w.shape)
        if not isinstance(t, torch.Tensor) or tuple(t.shape) != shape or t.device != dev:
            self.eligibility_trace = torch.zeros(*shape, device=dev)

    @property
    def eligibility(self) -> torch.Tensor:
        """``[n_src, n_tgt]`` (learning.py:2245-2247)."""
        return torch.outer(self.p_plus.view(-1), self._spost.float().view(-1)) + torch.outer(self._spre.float().view(-1), self.p_minus.view(-1))

    def _fill_desc(self, d: "_abi.SnnConn") -> None:
        super()._fill_desc(d)
        dt = float(self.connection.dt)
        d.e_trace = self.eligibility_trace.data_ptr()
        d.e_trace_decay = float(torch.exp(-dt / self.tc_e_trace))          # learning.py:2229
        d.tc_e_trace = float(self.tc_e_trace)
        # update = nu[0] * dt * reward * eligibility_trace (learning.py:2232): the scalar product in fp32, left to right
        d.et_coef = float(self.nu[0].float() * dt * float(self._run_kwargs["reward"]))


def _unsupported(name: str, where: str):
    class _Unsupported(LearningRule):
        __doc__ = f"``{name}`` (reference: {where}) — not on the accelerated path (SURVEY.md §8f)."

        def __init__(self, *args, **kwargs):
            raise NotImplementedError(f"learning.{name} is outside the hot path bindsnet_b200 implements")

    _Unsupported.__name__ = name
    return _Unsupported


class Rmax(LearningRule):
    """Reward-maximisation rule for stochastic ``SRM0Nodes`` targets (reference: learning.py:2858-2960; update
    :2921-2960): per synapse an eligibility trace that decays by ``1 - dt / tc_e_trace`` and gains
    ``(s_post - p / (1 + tc_c / dt * p)) * x_pre`` each step (``p`` the target's spike probability of the step), and
    ``w += nu[0] * reward * eligibility``.  Its target draws from torch's generator, so — like ``SRM0Nodes`` — the rule
    has no kernel form: it is host torch code on the connection's device and the network runs on the scripted tier.
    As in the reference the flattened views make it a batch-size-1 rule."""

    rule_code = None

    def __init__(self, connection, nu=None, reduction=None, weight_decay: float = 0.0, **kwargs) -> None:
        super().__init__(connection=connection, nu=nu, reduction=reduction, weight_decay=weight_decay, **kwargs)
        from ..network.nodes import SRM0Nodes
        from ..network.topology import Connection

        assert self.source.traces and self.source.traces_additive, "Pre-synaptic nodes must use additive spike traces."
        assert isinstance(self.target, SRM0Nodes), "R-max needs stochastically firing neurons, use SRM0Nodes."
        if not isinstance(connection, Connection):        # Connection and its subclass LocalConnection (learning.py:2905-2910)
            raise NotImplementedError("This learning rule is not supported for this Connection type.")
        self.tc_c = torch.tensor(kwargs.get("tc_c", 5.0))                 # 0: naive Hebbian ... inf: policy gradient
        self.tc_e_trace = torch.tensor(kwargs.get("tc_e_trace", 25.0))

    def update(self, **kwargs) -> None:
        w, dt = self.connection.w, self.connection.dt
        if not hasattr(self, "eligibility_trace"):
            self.eligibility_trace = torch.zeros(*w.shape, device=w.device)
        fired = self.target.s.view(-1).float()
        p = self.target.s_prob.view(-1)
        self.eligibility_trace *= 1 - dt / self.tc_e_trace
        self.eligibility_trace += (fired - p / (1.0 + self.tc_c / dt * p)) * self.source.x.view(-1)[:, None]
        with torch.no_grad():
            w += self.nu[0] * kwargs["reward"] * self.eligibility_trace
        super().update()
