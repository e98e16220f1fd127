"""ctypes loader of the CPU oracle (``oracle/snn_oracle.c``).

TEST INFRASTRUCTURE: imported only by ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` leg.  ``bindsnet_b200`` never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from bindsnet_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsnn_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off, OpenMP)."""
    src = os.path.join(HERE, "snn_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "snn_b200.h")
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(f) > os.path.getmtime(LIB) for f in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
        L.snn_oracle_run_window.restype = C.c_int
        L.snn_oracle_run_window.argtypes = [C.POINTER(_abi.SnnNet), C.POINTER(_abi.SnnRunOpts), C.c_int, C.c_int]
        L.snn_oracle_delta_apply.restype = C.c_int
        L.snn_oracle_delta_apply.argtypes = [vp, vp, vp, i32, i32, i32, f32, f32, i32, i32, f32]
        L.snn_oracle_conn_compute.restype = C.c_int
        L.snn_oracle_conn_compute.argtypes = [C.POINTER(_abi.SnnConn), i32, i32, i32, vp, vp]
        L.snn_oracle_conn_update.restype = C.c_int
        L.snn_oracle_conn_update.argtypes = [C.POINTER(_abi.SnnNet), i32, i32]
        L.snn_oracle_conn_normalize.restype = C.c_int
        L.snn_oracle_conn_normalize.argtypes = [C.POINTER(_abi.SnnConn), i32, i32]
        L.snn_oracle_abi_version.restype = C.c_int
        assert L.snn_oracle_abi_version() == _abi.SNN_ABI_VERSION, "oracle ABI mismatch: rebuild (make -C oracle)"
        _lib = L
    return _lib


def run_window(net: _abi.SnnNet, opts: _abi.SnnRunOpts, dense: int = 0, threads: int = 0) -> int:
    """Run one window on HOST tensors.  Returns the device-style error flags."""
    err = C.c_int32(0)
    opts.err_flag = C.addressof(err)
    rc = lib().snn_oracle_run_window(C.byref(net), C.byref(opts), dense, threads)
    opts.err_flag = None
    if rc != _abi.SNN_OK:
        raise RuntimeError("oracle: " + _abi.describe_error(rc))
    return int(err.value)


class OracleBackend:
    """Context manager that routes ``bindsnet_b200`` host objects (on CPU tensors) to the
    oracle — so the host-side logic (plan building, batch inference, monitors, kwargs) can be
    tested without a GPU, and so GPU tests have a like-for-like CPU twin to compare with."""

    def __init__(self, dense: int = 0, threads: int = 0):
        self.dense, self.threads = dense, threads
        self.err = 0

    def __enter__(self):
        from bindsnet_b200 import _backend
        from bindsnet_b200.network.network import Network

        self._saved = (Network._launch, _backend.conn_compute, _backend.conn_update, _backend.conn_normalize)
        outer = self

        def _launch(net_self, net, opts, dev):
            outer.err |= run_window(net, opts, outer.dense, outer.threads)

        def _compute(conn, n_src, n_tgt, B, s, out):
            assert lib().snn_oracle_conn_compute(C.byref(conn), n_src, n_tgt, B, s.data_ptr(), out.data_ptr()) == 0

        def _update(net, ci, B, device):
            assert lib().snn_oracle_conn_update(C.byref(net), ci, B) == 0

        def _normalize(conn, n_src, n_tgt, device):
            assert lib().snn_oracle_conn_normalize(C.byref(conn), n_src, n_tgt) == 0

        Network._launch = _launch
        _backend.conn_compute, _backend.conn_update, _backend.conn_normalize = _compute, _update, _normalize
        self._req = _backend.require_cuda
        _backend.require_cuda = lambda t, what: None
        return self

    def __exit__(self, *exc):
        from bindsnet_b200 import _backend
        from bindsnet_b200.network.network import Network

        Network._launch, _backend.conn_compute, _backend.conn_update, _backend.conn_normalize = self._saved
        _backend.require_cuda = self._req
        return False
