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
_SOURCES = [os.path.join(HERE, "emu_lib.cpp"), os.path.join(HERE, "cuda_emu.h")] + [
    os.path.join(ROOT, "bindsnet_b200", "csrc", f) for f in ("snn_generic.cu", "snn_phases.cuh", "snn_common.cuh", "snn_api.cu", "snn_combine.cuh", "snn_fused_dc.cu")
] + [os.path.join(ROOT, "include", "snn_b200.h")]
_lib = None


def build(force: bool = False) -> str:
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(f) > os.path.getmtime(LIB) for f in _SOURCES)
    if force or stale:
        cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-DSNN_EMU", "-ffp-contract=off", "-Wno-unknown-pragmas",
               "-I" + HERE, "-o", LIB, os.path.join(HERE, "emu_lib.cpp"), "-lpthread"]
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
        vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
        L.snn_b200_delta_apply.restype = C.c_int
        L.snn_b200_delta_apply.argtypes = [vp, vp, vp, i32, i32, i32, f32, f32, i32, i32, f32, vp]
        L.snn_b200_delta_apply_fused.restype = C.c_int
        L.snn_b200_delta_apply_fused.argtypes = [vp, vp, i32, i32, i32, f32, f32, i32, i32, f32, vp, vp, i32, vp]
        L.snn_b200_abi_version.restype = C.c_int
        assert L.snn_b200_abi_version() == _abi.SNN_ABI_VERSION
        _lib = L
    return _lib


def run_window(net: _abi.SnnNet, opts: _abi.SnnRunOpts) -> int:
    """One window on HOST tensors through the emulated generic kernel.  Returns the device-style error flags."""
    L = lib()
    err = C.c_int32(0)
    opts.err_flag = C.addressof(err)
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
        return self

    def __exit__(self, *exc):
        from bindsnet_b200 import _backend
        from bindsnet_b200.network.network import Network

        Network._launch = self._saved
        _backend.require_cuda = self._req
        return False
