"""Plan building: turns the user's objects into the POD structs of ``include/snn_b200.h``.

``Network.run`` calls ``build_net`` once per window; the standalone object methods
(``Nodes.forward``, ``Connection.compute`` / ``update`` / ``normalize``) go through the
single-operator entry points.  Everything here is host bookkeeping — no arithmetic.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from .. import _abi, _backend


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _as_u8(t: torch.Tensor) -> torch.Tensor:
    """Reinterpret a bool tensor as uint8 without copying (same storage)."""
    return t.view(torch.uint8) if t.dtype == torch.bool else t


def _state(t: torch.Tensor, name: str, owner: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{owner}.{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{owner}.{name} must be contiguous (it is updated in place by the CUDA core)")
    return t


def fill_layer(d: "_abi.SnnLayer", layer, name: str, B: int) -> None:
    layer._fill_desc(d)
    if layer.s.dtype not in (torch.bool, torch.uint8) or tuple(layer.s.shape) != (B, *layer.shape):
        # Input.forward aliases user input into s in the reference (nodes.py:219); we keep a
        # private bool tensor of the canonical shape instead.
        layer.s = torch.zeros(B, *layer.shape, dtype=torch.bool, device=layer.s.device)
    if not layer.s.is_contiguous():
        layer.s = layer.s.contiguous()
    d.s = _ptr(_as_u8(layer.s))
    if d.kind != _abi.SNN_NODE_INPUT:
        d.v = _ptr(_state(layer.v, "v", name))
        if d.kind != _abi.SNN_NODE_MCP:
            d.refrac_count = _ptr(_state(layer.refrac_count, "refrac_count", name))
    if d.kind == _abi.SNN_NODE_DC:
        d.theta = _ptr(_state(layer.theta, "theta", name))
    if d.kind == _abi.SNN_NODE_CURRENT_LIF:
        d.i = _ptr(_state(layer.i, "i", name))
    if layer.traces:
        d.x = _ptr(_state(layer.x, "x", name))
    if layer.sum_input:
        d.summed = _ptr(_state(layer.summed, "summed", name))


def weight_structure(conn, w: torch.Tensor):
    """Plan-time structure detection for STATIC square weight matrices (DiehlAndCook2015's
    ``exc * I`` and ``-inh * (1 - I)``, models.py:204,217-220): lets the fused kernel replace an
    n x n matrix by one constant.  Verified on the actual tensor (one small reduction + host
    read) and cached until the tensor is modified in place (``Tensor._version``) or replaced."""
    key = (w.data_ptr(), w._version, tuple(w.shape), str(w.device))
    cached = getattr(conn, "_b200_structure", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    result = (_abi.SNN_W_DENSE, 0.0)
    if w.dim() == 2 and w.shape[0] == w.shape[1] and w.shape[0] > 1:
        with torch.no_grad():
            diag = torch.diagonal(w)
            d0, o0 = diag[0], w[0, 1]
            eye = torch.eye(w.shape[0], dtype=torch.bool, device=w.device)
            is_diag = bool(((w == 0) | eye).all() & (diag == d0).all())
            is_off = bool(((w == o0) | eye).all() & (diag == 0).all())
        if is_diag:
            result = (_abi.SNN_W_DIAG, float(d0))
        elif is_off:
            result = (_abi.SNN_W_OFFDIAG, float(o0))
    conn._b200_structure = (key, result)
    return result


def fill_conn(d: "_abi.SnnConn", conn, src_idx: int, tgt_idx: int, dt: float, B: int, rule_kwargs=None, rule: bool = True) -> None:
    """``rule=False``: geometry, weights and normalisation only (the single-operator compute / normalize calls
    do not involve the learning rule, whose plan entry may need a run's keyword arguments)."""
    d.src, d.tgt = src_idx, tgt_idx
    rule0 = getattr(conn, "update_rule", None)
    if rule and hasattr(rule0, "_prepare"):  # rules with state of their own (MSTDP): allocate for this batch size / device
        rule0._prepare(B, conn.w.device, rule_kwargs or {})
    conn._fill_desc(d, dt, rule)
    w = conn.w
    if w.dtype != torch.float32 or not w.is_contiguous():
        raise TypeError("connection weights must be contiguous float32")
    if d.kind != _abi.SNN_CONN_CONV2D and tuple(w.shape) != (conn.source.n, conn.target.n):
        raise ValueError(f"weight shape {tuple(w.shape)} != ({conn.source.n}, {conn.target.n})")
    d.w = _ptr(w)
    static = d.rule == _abi.SNN_RULE_NONE or (d.rule == _abi.SNN_RULE_NOOP and d.weight_decay in (0.0, 1.0))
    if static and not d.has_norm:
        d.structure, d.structure_val = weight_structure(conn, w)
    b = getattr(conn, "b", None)
    d.b = _ptr(b) if b is not None else None
    rule = getattr(conn, "update_rule", None)
    if rule is None and hasattr(conn, "pipeline"):
        rule = conn.pipeline[0].learning_rule
    if d.rule >= _abi.SNN_RULE_POSTPRE and getattr(rule, "_squeeze", False) and B != 1:
        # The reference would fail inside torch with a broadcast error.
        raise RuntimeError(
            "learning rule was built with reduction=torch.squeeze (source.batch_size == 1 at construction) "
            f"but the run uses batch size {B}; pass reduction=torch.sum like the reference requires"
        )


def network_device(network) -> torch.device:
    for layer in network.layers.values():
        return layer.s.device
    return torch.device("cpu")


def build_net(
    network,
    B: int,
    ext: Dict[str, torch.Tensor],
    clamps: Dict[str, torch.Tensor],
    unclamps: Dict[str, torch.Tensor],
    injects: Dict[str, torch.Tensor],
    rec: Dict[str, Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]],
) -> Tuple["_abi.SnnNet", List[torch.Tensor]]:
    """Describe ``network`` for one window.  ``ext`` maps layer name to a contiguous
    ``[T, B, n]`` device tensor (uint8/bool or float32)."""
    if len(network.layers) > _abi.SNN_MAX_LAYERS or len(network.connections) > _abi.SNN_MAX_CONNS:
        raise NotImplementedError(
            f"networks with more than {_abi.SNN_MAX_LAYERS} layers / {_abi.SNN_MAX_CONNS} connections "
            "are not supported by this build"
        )
    net = _abi.SnnNet()
    net.abi_version = _abi.SNN_ABI_VERSION
    net.n_layers = len(network.layers)
    net.n_conns = len(network.connections)
    net.learning = int(bool(network.learning))
    keep: List[torch.Tensor] = []
    index = {}
    for i, (name, layer) in enumerate(network.layers.items()):
        index[name] = i
        d = net.layers[i]
        fill_layer(d, layer, name, B)
        e = ext.get(name)
        if e is not None:
            d.ext = _ptr(_as_u8(e))
            d.ext_dtype = _abi.SNN_EXT_F32 if e.dtype == torch.float32 else _abi.SNN_EXT_U8
            keep.append(e)
        for key, table, field, flag in (
            ("clamp", clamps, "clamp", "clamp_per_step"),
            ("unclamp", unclamps, "unclamp", "unclamp_per_step"),
            ("inject_v", injects, "inject_v", "inject_per_step"),
        ):
            m = table.get(name)
            if m is not None:
                setattr(d, field, _ptr(_as_u8(m)))
                setattr(d, flag, int(m.dim() == 2))
                keep.append(m)
        r = rec.get(name)
        if r is not None:
            if r[0] is not None:
                d.rec_s = _ptr(_as_u8(r[0])); keep.append(r[0])
            if r[1] is not None:
                d.rec_v = _ptr(r[1]); keep.append(r[1])
            if len(r) > 2 and r[2] is not None:
                d.rec_count = _ptr(r[2]); keep.append(r[2])
    masks = getattr(network, "_conn_masks", None) or {}
    for i, ((src, tgt), conn) in enumerate(network.connections.items()):
        fill_conn(net.conns[i], conn, index[src], index[tgt], float(network.dt), B, network._rule_kwargs_of((src, tgt)))
        m = masks.get((src, tgt))
        if m is not None:
            net.conns[i].mask = _ptr(m)
            keep.append(m)
    return net, keep


# ---- single-operator helpers ---------------------------------------------------------------

def _conn_desc(conn, B: int, dt: float = 1.0, rule: bool = True) -> "_abi.SnnConn":
    d = _abi.SnnConn()
    fill_conn(d, conn, 0, 1, dt, B, rule=rule)
    return d


def compute_single_connection(conn, s: torch.Tensor) -> torch.Tensor:
    """``conn.compute(s)``: ``[B, *target.shape]`` currents for spikes ``s``."""
    B = s.shape[0]
    _backend.require_cuda(conn.w, "connection weights")
    su8 = _as_u8(s if s.dtype in (torch.bool, torch.uint8) else (s != 0)).reshape(B, -1).contiguous()
    su8 = su8.to(conn.w.device)
    out = torch.empty(B, conn.target.n, dtype=torch.float32, device=conn.w.device)
    d = _conn_desc(conn, B, rule=False)
    _backend.conn_compute(d, conn.source.n, conn.target.n, B, su8, out)
    return out.view(B, *conn.target.shape)


def _pair_net(conn, B: int) -> "_abi.SnnNet":
    net = _abi.SnnNet()
    net.abi_version = _abi.SNN_ABI_VERSION
    net.n_layers, net.n_conns, net.learning = 2, 1, 1
    dt = getattr(conn, "dt", None) or 1.0
    for k, layer in enumerate((conn.source, conn.target)):
        if layer.dt is None:
            layer.compute_decays(dt)
        if layer.kind is None:
            _fill_endpoint(net.layers[k], layer, f"layer{k}", B)
        else:
            fill_layer(net.layers[k], layer, f"layer{k}", B)
    fill_conn(net.conns[0], conn, 0, 1, float(dt), B)
    return net


def _fill_endpoint(d: "_abi.SnnLayer", layer, name: str, B: int) -> None:
    """A USER-DEFINED population (no ``kind``: its ``forward`` is torch code, scripted tier) as the source or target of a
    built-in connection's single-operator update: the rule reads the population's current spikes and traces and nothing
    else (snn_b200_conn_update), so that is all the descriptor carries."""
    if layer.s.dtype not in (torch.bool, torch.uint8) or layer.s.numel() != B * layer.n:
        raise TypeError(f"{name}.s must be a bool / uint8 tensor of {B} x {layer.n} spikes, got {layer.s.dtype} {tuple(layer.s.shape)}")
    if not layer.s.is_contiguous():
        layer.s = layer.s.contiguous()
    d.kind, d.n = _abi.SNN_NODE_INPUT, layer.n
    d.traces = int(bool(layer.traces))
    d.traces_additive = int(bool(layer.traces_additive))
    d.learning = int(bool(getattr(layer, "learning", True)))
    d.dt = float(layer.dt) if layer.dt is not None else 1.0
    d.s = _ptr(_as_u8(layer.s))
    if layer.traces:
        d.x = _ptr(_state(layer.x, "x", name))


def update_single_connection(conn) -> None:
    """``conn.update(learning=True)`` from the layers' current ``s`` / ``x``."""
    B = conn.source.s.shape[0]
    _backend.require_cuda(conn.w, "connection weights")
    net = _pair_net(conn, B)
    _backend.conn_update(net, 0, B, conn.w.device)


def normalize_single_connection(conn) -> None:
    _backend.require_cuda(conn.w, "connection weights")
    d = _conn_desc(conn, 1, rule=False)
    _backend.conn_normalize(d, conn.source.n, conn.target.n, conn.w.device)


def normalize_feature(feature) -> None:
    _backend.require_cuda(feature.value, "feature value")
    d = _abi.SnnConn()
    d.w = feature.value.data_ptr()
    d.has_norm, d.norm_abs, d.norm = 1, 0, float(feature.norm)
    _backend.conn_normalize(d, feature.value.shape[0], feature.value.shape[1], feature.value.device)


def step_single_layer(layer, x: torch.Tensor) -> None:
    """``layer.forward(x)``: one step of one population, submitted as a one-layer window."""
    from .network import Network

    if layer.dt is None:
        raise RuntimeError("add the layer to a Network (or call compute_decays/set_batch_size) before forward()")
    host = Network(dt=float(layer.dt), batch_size=layer.batch_size or x.shape[0], learning=layer.learning)
    host.layers["L"] = layer  # bypass add_layer: keep the layer's state and batch size
    host._run_window({"L": x.unsqueeze(0)}, 1, normalize=False)
