"""Build libsnn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python bindsnet_b200/csrc/build.py [--force] [--verbose]

--fmad=false keeps every fp32 multiply and add separately rounded, which is what makes the
kernels bit-compatible with the reference's op-by-op ATen arithmetic (and with the oracle).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["snn_api.cu", "snn_generic.cu", "snn_fused_dc.cu", "snn_fused_dc2.cu", "snn_ops.cu", "snn_encode.cu", "snn_readout.cu"]
HEADERS = ["snn_common.cuh", "snn_phases.cuh", "snn_combine.cuh", os.path.join("..", "..", "include", "snn_b200.h")]
OUT = os.path.join(HERE, "libsnn_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++", "-cudart", "static",
]


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS + ["build.py"])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(HERE, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libsnn_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
