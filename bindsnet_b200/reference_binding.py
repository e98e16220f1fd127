"""The binding a BindsNET maintainer would add to the REFERENCE: fill the C ABI of ``include/snn_b200.h`` straight
from live ``bindsnet`` objects (no ``bindsnet_b200`` host classes involved) and run one ``Network.run`` window through
a library that implements it — ``libsnn_b200.so`` on CUDA tensors, or the oracle library on CPU tensors (which is how
``tests/test_reference_binding.py`` proves, without a GPU, that the ABI can be driven from the reference's own
``DiehlAndCook2015`` and reproduces the reference's own ``run``).

    from bindsnet_b200 import reference_binding as rb
    rb.run_window(reference_network, {"X": spikes}, time=250)        # drop-in for network.run(...)

Covered: what ``bindsnet.models`` builds for the hot path — ``Input`` / ``LIFNodes`` / ``DiehlAndCookNodes`` layers,
``MulticompartmentConnection`` with one ``Weight`` feature (``MCC_learning.NoOp`` / ``PostPre``) and the classic
``Connection`` with ``learning.NoOp`` / ``PostPre`` / ``WeightDependentPostPre`` / ``Hebbian``.  Every attribute is read where the
reference keeps it (file:line in the comments); state tensors are handed over by pointer and updated in place.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import torch

from . import _abi

_KIND = {"Input": _abi.SNN_NODE_INPUT, "LIFNodes": _abi.SNN_NODE_LIF, "DiehlAndCookNodes": _abi.SNN_NODE_DC,
         "AdaptiveLIFNodes": _abi.SNN_NODE_DC,            # DiehlAndCookNodes.forward without the one-spike arbitration (nodes.py:921-946)
         "IFNodes": _abi.SNN_NODE_IF, "CurrentLIFNodes": _abi.SNN_NODE_CURRENT_LIF, "BoostedLIFNodes": _abi.SNN_NODE_BOOSTED_LIF,
         "McCullochPitts": _abi.SNN_NODE_MCP}


def _f(x) -> float:
    return float(x.item() if isinstance(x, torch.Tensor) else x)


def _u8(t: torch.Tensor) -> torch.Tensor:
    return t.view(torch.uint8) if t.dtype == torch.bool else t


def _reduction_code(fn) -> int:
    if fn in (torch.sum, torch.squeeze):       # learning.py:76-80 / MCC_learning.py:68-74
        return _abi.SNN_REDUCE_SUM
    if fn is torch.mean:
        return _abi.SNN_REDUCE_MEAN
    raise NotImplementedError(f"reduction {fn} is not one the core implements (sum, mean)")


def fill_layer(d: "_abi.SnnLayer", layer, B: int, keep: List[torch.Tensor]) -> None:
    """bindsnet.network.nodes: Nodes.__init__ nodes.py:15-86, LIFNodes :425-498, DiehlAndCookNodes :988-1066."""
    kind = _KIND.get(type(layer).__name__)
    if kind is None:
        raise NotImplementedError(f"{type(layer).__name__} is outside the accelerated path")
    d.kind, d.n = kind, int(layer.n)
    d.traces, d.traces_additive = int(layer.traces), int(layer.traces_additive)
    d.sum_input, d.learning = int(layer.sum_input), int(layer.learning)
    d.dt = _f(layer.dt)
    if layer.traces:
        d.trace_decay, d.trace_scale = _f(layer.trace_decay), _f(layer.trace_scale)      # nodes.py:129-131
        d.x = layer.x.data_ptr()
    if layer.sum_input:
        d.summed = layer.summed.data_ptr()
    # Input.forward aliases the caller's input into s (nodes.py:219): give the core a private bool tensor
    if layer.s.dtype not in (torch.bool, torch.uint8) or tuple(layer.s.shape) != (B, *layer.shape) or not layer.s.is_contiguous():
        layer.s = torch.zeros(B, *layer.shape, dtype=torch.bool, device=layer.s.device)
    d.s = _u8(layer.s).data_ptr()
    if kind != _abi.SNN_NODE_INPUT:
        d.v = layer.v.data_ptr()
        d.thresh = _f(layer.thresh)
        if kind != _abi.SNN_NODE_MCP:                                                     # McCullochPitts: v = x, no other state
            d.refrac_count, d.refrac = layer.refrac_count.data_ptr(), _f(layer.refrac)
        if hasattr(layer, "decay") and kind not in (_abi.SNN_NODE_IF, _abi.SNN_NODE_MCP):
            d.decay = _f(layer.decay)                                                     # nodes.py:546-548, 1128-1130
        if hasattr(layer, "rest"):
            d.rest = _f(layer.rest)
        if hasattr(layer, "reset"):
            d.reset = _f(layer.reset)
        lb = getattr(layer, "lbound", None)
        d.has_lbound, d.lbound = int(lb is not None), (_f(lb) if lb is not None else 0.0)
    if kind == _abi.SNN_NODE_CURRENT_LIF:
        d.i, d.i_decay = layer.i.data_ptr(), _f(layer.i_decay)                            # nodes.py:771, 818-820
    if kind == _abi.SNN_NODE_DC:
        d.theta = layer.theta.data_ptr()
        d.theta_plus, d.theta_decay = _f(layer.theta_plus), _f(layer.theta_decay)         # nodes.py:1131-1133
        d.one_spike = int(getattr(layer, "one_spike", False))


def fill_connection(d: "_abi.SnnConn", conn, src: int, tgt: int, dt: float, keep: Optional[List[torch.Tensor]] = None) -> None:
    keep = keep if keep is not None else []
    d.src, d.tgt = src, tgt
    d.weight_decay, d.dt_scale = 1.0, 1.0
    if hasattr(conn, "pipeline"):
        # MulticompartmentConnection (topology.py:402-537) with one Weight feature (topology_features.py:575-671)
        if len(conn.pipeline) != 1 or type(conn.pipeline[0]).__name__ != "Weight":
            raise NotImplementedError("only MulticompartmentConnection pipelines of exactly one Weight feature")
        feat = conn.pipeline[0]
        rule = feat.learning_rule                                                         # an MCC_learning instance after priming
        d.kind = _abi.SNN_CONN_MCC
        w = feat.value
        d.has_norm = int(feat.norm is not None)                                           # topology_features.py:250-266: plain sum
        d.norm, d.norm_abs = (_f(feat.norm) if feat.norm is not None else 0.0), 0
        name = type(rule).__name__
        if name == "NoOp" or conn.manual_update:                                          # MCC_learning.py:120-146; topology.py:509-518
            d.rule = _abi.SNN_RULE_NONE
        elif name == "PostPre":
            d.rule = _abi.SNN_RULE_MCC_POSTPRE                                            # MCC_learning.py:224-302
            d.nu0, d.nu1 = _f(rule.nu[0]), _f(rule.nu[1])
            d.reduction = _reduction_code(rule.reduction)
            d.weight_decay = _f(rule.decay)                                               # MCC_learning.py:84: 1 - decay (1.0 = off)
            d.dt_scale = _f(conn.dt if getattr(conn, "dt", None) is not None else dt)     # MCC_learning.py:262,298
            lo, hi = rule.min, rule.max
            d.wmin = _f(lo) if lo is not None else -math.inf
            d.wmax = _f(hi) if hi is not None else math.inf
            d.has_clamp = int(lo is not None or hi is not None)                           # MCC_learning.py:101-110
        else:
            raise NotImplementedError(f"MCC learning rule {name}")
    else:
        # Connection (topology.py:265-399) + learning.LearningRule (learning.py:31-104)
        if type(conn).__name__ not in ("Connection", "LocalConnection"):
            raise NotImplementedError(f"{type(conn).__name__} is outside the accelerated path")
        rule = conn.update_rule
        d.kind = _abi.SNN_CONN_DENSE
        w = conn.w
        d.has_norm = int(conn.norm is not None)                                           # topology.py:383-392: sum of |w|
        d.norm, d.norm_abs = (_f(conn.norm) if conn.norm is not None else 0.0), 1
        if type(conn).__name__ == "LocalConnection":
            # a dense matrix confined to its receptive fields by the connection's own mask (topology.py:1431, 1457-1469),
            # plain column sums in normalize (:1471-1479; norm is already scaled by the kernel size, :1437-1438)
            d.norm_abs = 0
            m = _u8(conn.mask.to(w.device)).contiguous()
            keep.append(m)
            d.mask = m.data_ptr()
        d.wmin, d.wmax = _f(conn.wmin), _f(conn.wmax)
        name = type(rule).__name__
        d.rule = {"NoOp": _abi.SNN_RULE_NOOP, "PostPre": _abi.SNN_RULE_POSTPRE, "Hebbian": _abi.SNN_RULE_HEBBIAN,
                  "WeightDependentPostPre": _abi.SNN_RULE_WDEP_POSTPRE}.get(name, -1)
        if d.rule < 0:
            raise NotImplementedError(f"learning rule {name}")
        d.nu0, d.nu1 = _f(rule.nu[0]), _f(rule.nu[1])
        d.reduction = _reduction_code(rule.reduction)
        d.weight_decay = _f(rule.weight_decay)                                            # learning.py:85
        finite = math.isfinite(d.wmin) or math.isfinite(d.wmax)
        d.has_clamp = int(finite and name != "NoOp")                                      # learning.py:97-104
        if conn.b is not None:
            d.b = conn.b.data_ptr()
    if w.dtype != torch.float32 or not w.is_contiguous():
        raise TypeError("weights must be contiguous float32")
    d.w = w.data_ptr()
    # plan-time structure hints for static square matrices (DiehlAndCook2015's exc / inh, models.py:204,217-220)
    static = d.rule == _abi.SNN_RULE_NONE or (d.rule == _abi.SNN_RULE_NOOP and d.weight_decay in (0.0, 1.0))
    if static and not d.has_norm and w.dim() == 2 and w.shape[0] == w.shape[1] and w.shape[0] > 1:
        with torch.no_grad():
            diag, eye = torch.diagonal(w), torch.eye(w.shape[0], dtype=torch.bool, device=w.device)
            if bool(((w == 0) | eye).all() & (diag == diag[0]).all()):
                d.structure, d.structure_val = _abi.SNN_W_DIAG, _f(diag[0])
            elif bool(((w == w[0, 1]) | eye).all() & (diag == 0).all()):
                d.structure, d.structure_val = _abi.SNN_W_OFFDIAG, _f(w[0, 1])


def build_net(network, inputs: Dict[str, torch.Tensor], T: int, B: int):
    """The window plan of a live reference ``Network`` (insertion orders: network.py:225, 386)."""
    net = _abi.SnnNet()
    net.abi_version = _abi.SNN_ABI_VERSION
    net.n_layers, net.n_conns = len(network.layers), len(network.connections)
    net.learning = int(bool(network.learning))
    keep: List[torch.Tensor] = []
    names = list(network.layers)
    for i, name in enumerate(names):
        layer = network.layers[name]
        fill_layer(net.layers[i], layer, B, keep)
        if name in inputs:                                                                # network.py:388-392
            x = inputs[name][:T]
            if x.dtype not in (torch.bool, torch.uint8, torch.float32):
                x = x.float()
            x = _u8(x.to(layer.s.device).reshape(T, B, layer.n).contiguous())
            net.layers[i].ext = x.data_ptr()
            net.layers[i].ext_dtype = _abi.SNN_EXT_F32 if x.dtype == torch.float32 else _abi.SNN_EXT_U8
            keep.append(x)
    for i, ((s, t), conn) in enumerate(network.connections.items()):
        fill_connection(net.conns[i], conn, names.index(s), names.index(t), float(network.dt), keep)
    return net, keep


def run_window(network, inputs: Dict[str, torch.Tensor], time: int, seed: Optional[int] = None, library: Optional[C.CDLL] = None) -> int:
    """Drop-in for ``network.run(inputs, time)`` of a reference ``Network`` (the body of network.py:329-465).
    ``library``: a CDLL exporting ``snn_oracle_run_window`` (CPU tensors) — default: the CUDA core on CUDA tensors."""
    inputs = dict(inputs)
    for k in inputs:                                                                       # network.py:329-340
        if inputs[k].dim() == 1:
            inputs[k] = inputs[k].unsqueeze(0).unsqueeze(0)
        elif inputs[k].dim() == 2:
            inputs[k] = inputs[k].unsqueeze(1)
    for k in inputs:                                                                       # network.py:342-353
        if inputs[k].size(1) != network.batch_size:
            network.batch_size = inputs[k].size(1)
            for layer in network.layers.values():
                layer.set_batch_size(network.batch_size)
        break
    T, B = int(time / network.dt), int(network.batch_size)
    net, keep = build_net(network, inputs, T, B)
    opts = _abi.SnnRunOpts()
    opts.T, opts.B, opts.normalize = T, B, 1
    opts.seed = (int(torch.randint(0, 2**31 - 1, (1,)).item()) if seed is None else seed) & 0xFFFFFFFF
    if library is not None:
        err = C.c_int32(0)
        opts.err_flag = C.cast(C.pointer(err), C.c_void_p).value
        library.snn_oracle_run_window.restype = C.c_int
        library.snn_oracle_run_window.argtypes = [C.POINTER(_abi.SnnNet), C.POINTER(_abi.SnnRunOpts), C.c_int, C.c_int]
        rc = library.snn_oracle_run_window(C.byref(net), C.byref(opts), 0, 0)
        del keep
        return rc | int(err.value)
    from . import _backend

    dev = next(iter(network.layers.values())).s.device
    _backend.run_window(net, opts, dev)
    del keep
    return 0
