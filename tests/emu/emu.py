"""ctypes loader of the EMULATED generic kernel (tests/emu/emu_lib.cpp: the product's CUDA sources compiled for the host
on a small CUDA-model emulation).  TEST INFRASTRUCTURE: imported by tests/ only; ``bindsnet_b200`` never loads it.

``EmuBackend`` routes the host API (on CPU tensors) to the emulated ``snn_b200_run_window`` the way
``oracle.oracle.OracleBackend`` routes it to the oracle, so the same golden cases can be replayed through the kernel's
real source without a GPU and compared with the oracle bit for bit."""
from __future__ import annotations

import ctypes as C
import mmap
import os
import subprocess

import numpy as np

from bindsnet_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
LIB = os.path.join(HERE, "libsnn_emu.so")
_SOURCES = [os.path.join(HERE, "emu_lib.cpp"), os.path.join(HERE, "emu_dc2.cpp"), os.path.join(HERE, "cuda_emu.h")] + [
    os.path.join(ROOT, "bindsnet_b200", "csrc", f) for f in ("snn_generic.cu", "snn_phases.cuh", "snn_common.cuh", "snn_api.cu", "snn_combine.cuh", "snn_fused_dc.cu", "snn_fused_dc2.cu", "snn_ops.cu", "snn_encode.cu", "snn_readout.cu")
] + [os.path.join(ROOT, "include", "snn_b200.h")]
_lib = None


def build(force: bool = False) -> str:
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(f) > os.path.getmtime(LIB) for f in _SOURCES)
    if force or stale:
        cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-DSNN_EMU", "-ffp-contract=off", "-Wno-unknown-pragmas",
               "-I" + HERE, "-o", LIB, os.path.join(HERE, "emu_lib.cpp"), os.path.join(HERE, "emu_dc2.cpp"), "-lpthread"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("building the emulated kernel failed:\n" + res.stderr[-4000:])
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.snn_b200_workspace_bytes.restype = C.c_size_t
        L.snn_b200_workspace_bytes.argtypes = [C.POINTER(_abi.SnnNet), C.POINTER(_abi.SnnRunOpts)]
        L.snn_b200_run_window.restype = C.c_int
        L.snn_b200_run_window.argtypes = [C.POINTER(_abi.SnnNet), C.POINTER(_abi.SnnRunOpts), C.c_void_p, C.c_size_t, C.c_void_p]
        L.snn_b200_select_tier.restype = C.c_int
        L.snn_b200_select_tier.argtypes = [C.POINTER(_abi.SnnNet), C.POINTER(_abi.SnnRunOpts)]
        vp, i32, f32, sz = C.c_void_p, C.c_int32, C.c_float, C.c_size_t
        L.snn_b200_delta_prepare.restype = C.c_int
        L.snn_b200_delta_prepare.argtypes = [vp, vp, vp, sz, vp]
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
        L.snn_b200_delta_apply.restype = C.c_int
        L.snn_b200_delta_apply.argtypes = [vp, vp, vp, i32, i32, i32, f32, f32, i32, i32, f32, vp]
        L.snn_b200_delta_apply_fused.restype = C.c_int
        L.snn_b200_delta_apply_fused.argtypes = [vp, vp, i32, i32, i32, f32, f32, i32, i32, f32, vp, vp, i32, vp]
        L.snn_b200_abi_version.restype = C.c_int
        assert L.snn_b200_abi_version() == _abi.SNN_ABI_VERSION
        _lib = L
    return _lib


last_tier = 0


def run_window(net: _abi.SnnNet, opts: _abi.SnnRunOpts) -> int:
    """One window on HOST tensors through the emulated generic kernel.  Returns the device-style error flags."""
    global last_tier
    L = lib()
    err = C.c_int32(0)
    opts.err_flag = C.addressof(err)
    last_tier = int(L.snn_b200_select_tier(C.byref(net), C.byref(opts)))   # the kernel this window goes to (tests assert on it)
    nbytes = int(L.snn_b200_workspace_bytes(C.byref(net), C.byref(opts)))
    # the workspace ends at a PROT_NONE guard page: a kernel that writes past snn_b200_workspace_bytes() faults
    page = mmap.PAGESIZE
    al = (max(nbytes, 8) + 255) & ~255
    total = ((al + page - 1) // page + 1) * page
    buf = mmap.mmap(-1, total)
    addr = C.addressof(C.c_char.from_buffer(buf))
    libc = C.CDLL(None, use_errno=True)
    libc.mprotect.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    assert libc.mprotect(addr + total - page, page, 0) == 0
    base = addr + total - page - al
    C.memset(base, 0xA5, al)   # the real workspace is uninitialised device memory: no kernel may count on zeros
    try:
        rc = L.snn_b200_run_window(C.byref(net), C.byref(opts), base, nbytes, None)
    finally:
        libc.mprotect(addr + total - page, page, 3)
        del addr
    opts.err_flag = None
    if rc != _abi.SNN_OK:
        raise RuntimeError("emulated kernel: " + _abi.describe_error(rc))
    return int(err.value)


class EmuBackend:
    """Context manager: ``Network.run`` on CPU tensors executes the generic kernel's source under emulation."""

    def __init__(self):
        self.err = 0

    def __enter__(self):
        from bindsnet_b200 import _backend
        from bindsnet_b200.network.network import Network

        self._saved = Network._launch
        outer = self

        def _launch(net_self, net, opts, dev):
            outer.err |= run_window(net, opts)

        Network._launch = _launch
        self._req = _backend.require_cuda
        _backend.require_cuda = lambda t, what: None
        # the single-operator entry points, the encoders and the read-out on host tensors
        L = lib()
        names = ("conn_compute", "conn_update", "conn_normalize", "delta_prepare", "delta_apply", "delta_apply_fused", "encode_poisson",
                 "encode_bernoulli", "assign_labels", "predict")
        self._ops = {n: getattr(_backend, n) for n in names}

        def ok(rc, what):
            if rc != _abi.SNN_OK:
                raise _backend.BackendError(what + ": " + _abi.describe_error(rc))

        ws_keep = []

        def conn_compute(conn, n_src, n_tgt, B, s_, out):
            ok(L.snn_b200_conn_compute(C.byref(conn), n_src, n_tgt, B, s_.data_ptr(), out.data_ptr(), None), "conn_compute")

        def conn_update(net, ci, B, device):
            ws = np.zeros(1 << 20, dtype=np.uint8)
            ws_keep.append(ws)
            ok(L.snn_b200_conn_update(C.byref(net), ci, B, (ws.ctypes.data + 255) & ~255, ws.size - 256, None), "conn_update")

        def conn_normalize(conn, n_src, n_tgt, device):
            ok(L.snn_b200_conn_normalize(C.byref(conn), n_src, n_tgt, None), "conn_normalize")

        def delta_prepare(w, w0, dw):
            ok(L.snn_b200_delta_prepare(w.data_ptr(), w0.data_ptr(), dw.data_ptr(), w.numel(), None), "delta_prepare")

        def delta_apply(w, w0, dw_sum, has_clamp, wmin, wmax, has_norm, norm_abs, norm):
            ok(L.snn_b200_delta_apply(w.data_ptr(), w0.data_ptr(), dw_sum.data_ptr(), w.shape[0], w.shape[1], int(has_clamp), float(wmin),
                                      float(wmax), int(has_norm), int(norm_abs), float(norm), None), "delta_apply")

        def delta_apply_fused(w, dw_sum, has_clamp, wmin, wmax, has_norm, norm_abs, norm, theta=None, dtheta_sum=None):
            ok(L.snn_b200_delta_apply_fused(w.data_ptr(), dw_sum.data_ptr(), w.shape[0], w.shape[1], int(has_clamp), float(wmin), float(wmax),
                                            int(has_norm), int(norm_abs), float(norm), theta.data_ptr() if theta is not None else None,
                                            dtheta_sum.data_ptr() if theta is not None else None, theta.numel() if theta is not None else 0,
                                            None), "delta_apply_fused")

        def encode_poisson(rate_hz, T, dt, seed, out):
            ok(L.snn_b200_encode_poisson(rate_hz.data_ptr(), rate_hz.numel(), T, float(dt), seed & (2**64 - 1), out.data_ptr(), None), "encode_poisson")

        def encode_bernoulli(prob, T, seed, out):
            ok(L.snn_b200_encode_bernoulli(prob.data_ptr(), prob.numel(), T, seed & (2**64 - 1), out.data_ptr(), None), "encode_bernoulli")

        def assign_labels(counts, labels, n_labels, alpha, rates, proportions, assignments):
            ok(L.snn_b200_assign_labels(counts.data_ptr(), labels.data_ptr(), counts.shape[0], counts.shape[1], n_labels, float(alpha),
                                        rates.data_ptr(), proportions.data_ptr(), assignments.data_ptr(), None), "assign_labels")

        def predict(counts, assignments, proportions, n_labels, predictions):
            ok(L.snn_b200_predict(counts.data_ptr(), assignments.data_ptr(), proportions.data_ptr() if proportions is not None else None,
                                  counts.shape[0], counts.shape[1], n_labels, predictions.data_ptr(), None), "predict")

        for n, f in (("conn_compute", conn_compute), ("conn_update", conn_update), ("conn_normalize", conn_normalize),
                     ("delta_prepare", delta_prepare), ("delta_apply", delta_apply), ("delta_apply_fused", delta_apply_fused),
                     ("encode_poisson", encode_poisson), ("encode_bernoulli", encode_bernoulli), ("assign_labels", assign_labels),
                     ("predict", predict)):
            setattr(_backend, n, f)
        return self

    def __exit__(self, *exc):
        from bindsnet_b200 import _backend
        from bindsnet_b200.network.network import Network

        Network._launch = self._saved
        _backend.require_cuda = self._req
        for n, f in self._ops.items():
            setattr(_backend, n, f)
        return False
