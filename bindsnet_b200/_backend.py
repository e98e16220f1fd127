"""ctypes binding of ``libsnn_b200.so`` (the CUDA core behind ``include/snn_b200.h``).

There is deliberately no CPU path: if the extension is missing or a tensor is not on a
CUDA device, the calls raise.  PyTorch is used for device memory (workspace, recordings) and
for the current CUDA stream only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsnn_b200.so")

_lib = None


class BackendError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the CUDA core.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BackendError(
            f"{LIB_PATH} not found: build the CUDA core first "
            "(python -c 'import __graft_entry__ as g; g.build()').  bindsnet_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, i32, f32, sz = C.c_void_p, C.c_int32, C.c_float, C.c_size_t
    L.snn_b200_abi_version.restype = C.c_int
    L.snn_b200_build_info.restype = C.c_char_p
    L.snn_b200_workspace_bytes.restype = sz
    L.snn_b200_workspace_bytes.argtypes = [C.POINTER(_abi.SnnNet), C.POINTER(_abi.SnnRunOpts)]
    L.snn_b200_run_window.restype = C.c_int
    L.snn_b200_run_window.argtypes = [C.POINTER(_abi.SnnNet), C.POINTER(_abi.SnnRunOpts), vp, sz, vp]
    L.snn_b200_select_tier.restype = C.c_int
    L.snn_b200_select_tier.argtypes = [C.POINTER(_abi.SnnNet), C.POINTER(_abi.SnnRunOpts)]
    L.snn_b200_last_launch_count.restype = C.c_int
    L.snn_b200_delta_prepare.restype = C.c_int
    L.snn_b200_delta_prepare.argtypes = [vp, vp, vp, sz, vp]
    L.snn_b200_delta_apply.restype = C.c_int
    L.snn_b200_delta_apply.argtypes = [vp, vp, vp, i32, i32, i32, f32, f32, i32, i32, f32, vp]
    L.snn_b200_delta_apply_fused.restype = C.c_int
    L.snn_b200_delta_apply_fused.argtypes = [vp, vp, i32, i32, i32, f32, f32, i32, i32, f32, vp, vp, i32, vp]
    L.snn_b200_conn_compute.restype = C.c_int
    L.snn_b200_conn_compute.argtypes = [C.POINTER(_abi.SnnConn), i32, i32, i32, vp, vp, vp]
    L.snn_b200_conn_update.restype = C.c_int
    L.snn_b200_conn_update.argtypes = [C.POINTER(_abi.SnnNet), i32, i32, vp, sz, vp]
    L.snn_b200_conn_normalize.restype = C.c_int
    L.snn_b200_conn_normalize.argtypes = [C.POINTER(_abi.SnnConn), i32, i32, vp]
    L.snn_b200_encode_poisson.restype = C.c_int
    L.snn_b200_encode_poisson.argtypes = [vp, i32, i32, f32, C.c_uint64, vp, vp]
    L.snn_b200_encode_bernoulli.restype = C.c_int
    L.snn_b200_encode_bernoulli.argtypes = [vp, i32, i32, C.c_uint64, vp, vp]
    L.snn_b200_assign_labels.restype = C.c_int
    L.snn_b200_assign_labels.argtypes = [vp, vp, i32, i32, i32, f32, vp, vp, vp, vp]
    L.snn_b200_predict.restype = C.c_int
    L.snn_b200_predict.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp]
    if L.snn_b200_abi_version() != _abi.SNN_ABI_VERSION:
        raise BackendError("libsnn_b200.so ABI version does not match bindsnet_b200/_abi.py — rebuild")
    _lib = L
    return L


def is_built() -> bool:
    return os.path.exists(LIB_PATH)


# ---- per-device scratch ---------------------------------------------------------------------
_workspaces: Dict[int, torch.Tensor] = {}
_err_dev: Dict[int, torch.Tensor] = {}
_err_host: Dict[int, torch.Tensor] = {}


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise BackendError(
            f"{what} lives on {t.device}; bindsnet_b200 executes on CUDA devices only "
            "(move the network with network.to('cuda'))"
        )


def _index(device: torch.device) -> int:
    return device.index if device.index is not None else torch.cuda.current_device()


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    idx = _index(device)
    ws = _workspaces.get(idx)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[idx] = ws
    return ws


def err_flag(device: torch.device) -> torch.Tensor:
    idx = _index(device)
    if idx not in _err_dev:
        _err_dev[idx] = torch.zeros(1, dtype=torch.int32, device=device)
        _err_host[idx] = torch.zeros(1, dtype=torch.int32).pin_memory()
    return _err_dev[idx]


def poll_errors(device: torch.device, sync: bool = False) -> None:
    """Raise if a previous window reported a device-side error.  Without ``sync`` this reads
    the host mirror filled by the asynchronous copy that follows every window (so an error
    surfaces at the latest on the next call); with ``sync`` it waits for the device."""
    idx = _index(device)
    if idx not in _err_dev:
        return
    if sync:
        code = int(_err_dev[idx].item())
    else:
        code = int(_err_host[idx][0])
    if code:
        _err_dev[idx].zero_()
        _err_host[idx].zero_()
        raise BackendError("CUDA window failed: " + _abi.describe_error(code))


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _check(rc: int, what: str) -> None:
    if rc != _abi.SNN_OK:
        raise BackendError(f"{what}: {_abi.describe_error(rc)}")


#: when set to a list, every window appends a (start, end) pair of CUDA events recorded on the
#: launching stream immediately around the kernel launch (bench.py's roofline measurement)
kernel_events: Optional[list] = None
#: number of window kernels launched through this module since import
launches_total = 0
#: kernel tier of the last window (1 generic, 2 fused DiehlAndCook2015 v1, 3 fused DiehlAndCook2015 v2)
last_tier = 0


def run_window(net: _abi.SnnNet, opts: _abi.SnnRunOpts, device: torch.device) -> None:
    """One ``Network.run`` window on ``device`` (asynchronous)."""
    global launches_total, last_tier
    L = lib()
    poll_errors(device)
    flag = err_flag(device)
    opts.err_flag = flag.data_ptr()
    with torch.cuda.device(device):
        nbytes = int(L.snn_b200_workspace_bytes(C.byref(net), C.byref(opts)))
        ws = workspace(device, nbytes)
        if kernel_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        rc = L.snn_b200_run_window(C.byref(net), C.byref(opts), ws.data_ptr(), ws.numel(), _stream_ptr(device))
        if kernel_events is not None:
            ev[1].record()
            kernel_events.append(ev)
        _check(rc, "snn_b200_run_window")
        last_tier = int(L.snn_b200_select_tier(C.byref(net), C.byref(opts)))
        launches_total += int(L.snn_b200_last_launch_count())
        _err_host[_index(device)].copy_(flag, non_blocking=True)


def select_tier(net: _abi.SnnNet, opts: _abi.SnnRunOpts) -> int:
    return int(lib().snn_b200_select_tier(C.byref(net), C.byref(opts)))


def last_launch_count() -> int:
    return int(lib().snn_b200_last_launch_count())


def conn_compute(conn: _abi.SnnConn, n_src: int, n_tgt: int, B: int, s: torch.Tensor, out: torch.Tensor) -> None:
    require_cuda(s, "spikes"); require_cuda(out, "output")
    with torch.cuda.device(s.device):
        _check(lib().snn_b200_conn_compute(C.byref(conn), n_src, n_tgt, B, s.data_ptr(), out.data_ptr(),
                                           _stream_ptr(s.device)), "snn_b200_conn_compute")


def conn_update(net: _abi.SnnNet, conn_index: int, B: int, device: torch.device) -> None:
    with torch.cuda.device(device):
        ws = workspace(device, 1 << 20)
        _check(lib().snn_b200_conn_update(C.byref(net), conn_index, B, ws.data_ptr(), ws.numel(),
                                          _stream_ptr(device)), "snn_b200_conn_update")


def conn_normalize(conn: _abi.SnnConn, n_src: int, n_tgt: int, device: torch.device) -> None:
    with torch.cuda.device(device):
        _check(lib().snn_b200_conn_normalize(C.byref(conn), n_src, n_tgt, _stream_ptr(device)),
               "snn_b200_conn_normalize")


def delta_prepare(w: torch.Tensor, w0: torch.Tensor, dw: torch.Tensor) -> None:
    global launches_total
    require_cuda(w, "w")
    launches_total += 1
    with torch.cuda.device(w.device):
        _check(lib().snn_b200_delta_prepare(w.data_ptr(), w0.data_ptr(), dw.data_ptr(), w.numel(),
                                            _stream_ptr(w.device)), "snn_b200_delta_prepare")


def delta_apply(w, w0, dw_sum, has_clamp, wmin, wmax, has_norm, norm_abs, norm) -> None:
    global launches_total
    require_cuda(w, "w")
    launches_total += 1
    with torch.cuda.device(w.device):
        _check(lib().snn_b200_delta_apply(w.data_ptr(), w0.data_ptr(), dw_sum.data_ptr(), w.shape[0], w.shape[1],
                                          int(has_clamp), float(wmin), float(wmax), int(has_norm), int(norm_abs),
                                          float(norm), _stream_ptr(w.device)), "snn_b200_delta_apply")


def delta_apply_fused(w, dw_sum, has_clamp, wmin, wmax, has_norm, norm_abs, norm, theta=None, dtheta_sum=None) -> None:
    """In-place combine after a delta window: ``w = clamp(w + dw_sum)``, normalize, ``theta += dtheta_sum`` — one launch."""
    global launches_total
    require_cuda(w, "w")
    launches_total += 1
    with torch.cuda.device(w.device):
        _check(lib().snn_b200_delta_apply_fused(w.data_ptr(), dw_sum.data_ptr(), w.shape[0], w.shape[1], int(has_clamp), float(wmin), float(wmax),
                                                int(has_norm), int(norm_abs), float(norm), theta.data_ptr() if theta is not None else None,
                                                dtheta_sum.data_ptr() if theta is not None else None, theta.numel() if theta is not None else 0,
                                                _stream_ptr(w.device)), "snn_b200_delta_apply_fused")


def encode_poisson(rate_hz: torch.Tensor, T: int, dt: float, seed: int, out: torch.Tensor) -> None:
    """``out[T, n]`` uint8 Poisson spike trains for ``rate_hz[n]`` (Hz) on the tensor's CUDA device."""
    global launches_total
    require_cuda(rate_hz, "rate image"); require_cuda(out, "spike tensor")
    launches_total += 1
    with torch.cuda.device(rate_hz.device):
        _check(lib().snn_b200_encode_poisson(rate_hz.data_ptr(), rate_hz.numel(), T, float(dt), seed & (2**64 - 1), out.data_ptr(),
                                             _stream_ptr(rate_hz.device)), "snn_b200_encode_poisson")


def encode_bernoulli(prob: torch.Tensor, T: int, seed: int, out: torch.Tensor) -> None:
    global launches_total
    require_cuda(prob, "probability image"); require_cuda(out, "spike tensor")
    launches_total += 1
    with torch.cuda.device(prob.device):
        _check(lib().snn_b200_encode_bernoulli(prob.data_ptr(), prob.numel(), T, seed & (2**64 - 1), out.data_ptr(),
                                               _stream_ptr(prob.device)), "snn_b200_encode_bernoulli")


def assign_labels(counts, labels, n_labels: int, alpha: float, rates, proportions, assignments) -> None:
    global launches_total
    require_cuda(counts, "spike counts")
    launches_total += 1
    with torch.cuda.device(counts.device):
        _check(lib().snn_b200_assign_labels(counts.data_ptr(), labels.data_ptr(), counts.shape[0], counts.shape[1], n_labels, float(alpha),
                                            rates.data_ptr(), proportions.data_ptr(), assignments.data_ptr(), _stream_ptr(counts.device)),
               "snn_b200_assign_labels")


def predict(counts, assignments, proportions, n_labels: int, predictions) -> None:
    global launches_total
    require_cuda(counts, "spike counts")
    launches_total += 1
    with torch.cuda.device(counts.device):
        _check(lib().snn_b200_predict(counts.data_ptr(), assignments.data_ptr(), proportions.data_ptr() if proportions is not None else None,
                                      counts.shape[0], counts.shape[1], n_labels, predictions.data_ptr(), _stream_ptr(counts.device)),
               "snn_b200_predict")
