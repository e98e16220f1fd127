"""Shared test plumbing: fixture loading, state snapshots, comparisons."""
from __future__ import annotations

import json
import os
from typing import Dict

import numpy as np
import torch

import cases  # tests/golden/cases.py (path added by conftest)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Fixture:
    def __init__(self, name: str):
        self.name = name
        z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
        self.z = z
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.T = self.meta["T"]

    def inputs(self) -> Dict[str, torch.Tensor]:
        out = {}
        for k, shape in self.meta["inputs"].items():
            n = int(np.prod(shape))
            if f"inf/{k}" in self.z:  # analog input current
                out[k] = torch.from_numpy(self.z[f"inf/{k}"].copy())
                continue
            out[k] = torch.from_numpy(np.unpackbits(self.z[f"in/{k}"])[:n].reshape(shape).copy())
        return out

    def build(self, device: str = "cpu"):
        """Our network for this case, on ``device``, plus staged inputs and kwargs."""
        ns = cases.namespace("b200")
        torch.manual_seed(1234)
        net, inputs, kw, T = cases.CASES[self.name](ns, inputs=self.inputs())
        assert T == self.T
        if device != "cpu":
            net.to(device)
        return net, inputs, kw, T


def conn_weight(conn) -> torch.Tensor:
    return conn.w


def snapshot(net) -> Dict[str, np.ndarray]:
    """All state the fixtures record, as numpy arrays keyed like the fixture."""
    out = {}
    for lname, layer in net.layers.items():
        B = layer.s.shape[0]
        out[f"L/{lname}/s"] = layer.s.reshape(B, -1).to(torch.uint8).cpu().numpy()
        for var in ("v", "refrac_count", "x", "summed", "i"):
            val = getattr(layer, var, None)
            if isinstance(val, torch.Tensor) and val.numel() > 0:
                out[f"L/{lname}/{var}"] = val.detach().reshape(B, -1).float().cpu().numpy()
        th = getattr(layer, "theta", None)
        if isinstance(th, torch.Tensor):
            out[f"L/{lname}/theta"] = th.detach().float().reshape(-1).cpu().numpy()
    for (s, t), conn in net.connections.items():
        out[f"C/{s}->{t}/w"] = conn.w.detach().float().cpu().numpy()
    return out


def add_spike_monitors(net, T: int, device: str = "cpu"):
    from bindsnet_b200.network.monitors import Monitor

    for lname, layer in net.layers.items():
        net.add_monitor(Monitor(layer, ["s"], time=T, device=device), f"mon_{lname}")


def spike_counts(net, T: int) -> Dict[str, np.ndarray]:
    out = {}
    for lname, layer in net.layers.items():
        B = layer.s.shape[0]
        r = net.monitors[f"mon_{lname}"].get("s").reshape(T, B, -1)
        out[f"L/{lname}/count"] = r.sum(dim=(0, 1)).to(torch.int32).cpu().numpy()
        out[f"L/{lname}/count_b"] = r.sum(dim=(0, 2)).to(torch.int32).cpu().numpy()
    return out


def run_case_oracle(name: str, dense: int = 0):
    """Run a golden case through the host API on the CPU oracle; returns (fixture, state, counts)."""
    from oracle.oracle import OracleBackend

    fx = Fixture(name)
    net, inputs, kw, T = fx.build("cpu")
    add_spike_monitors(net, T)
    with OracleBackend(dense=dense) as ob:
        net.run(inputs=inputs, time=T, one_spike_seed=cases.ONE_SPIKE_SEED, **kw)
        assert ob.err == 0
    return fx, snapshot(net), spike_counts(net, T)


def assert_close_to_golden(fx: Fixture, state, counts, rtol_w=1e-4, atol_state=1e-4, rtol_state=1e-5, count_slack=0, atol_w=2e-6):
    """The north_star's parity statement: final weights within 1e-4 relative, voltages/traces
    within an fp32 tolerance, spike rasters compared by per-neuron count."""
    z = fx.z
    for key, val in state.items():
        if key.startswith("C/"):
            if key in z:
                ref = z[key]
                err = np.abs(val - ref).max() / max(np.abs(ref).max(), 1e-12)
                assert err <= rtol_w, f"{fx.name} {key}: max rel err {err:.3e}"
                # element-wise as well: |d| <= rtol * |ref| + atol, atol = a few fp32 ulps of the largest weight
                bad = np.abs(val - ref) > rtol_w * np.abs(ref) + atol_w * max(np.abs(ref).max(), 1e-12)
                assert not bad.any(), f"{fx.name} {key}: {bad.sum()} entries beyond the element-wise tolerance, max |d| {np.abs(val - ref).max():.3e}"
            else:  # large case: subsampled rows + column sums
                ref = z[key + "_rows8"]
                err = np.abs(val[::8] - ref).max() / max(np.abs(ref).max(), 1e-12)
                assert err <= rtol_w, f"{fx.name} {key} rows: max rel err {err:.3e}"
                bad = np.abs(val[::8] - ref) > rtol_w * np.abs(ref) + atol_w * max(np.abs(ref).max(), 1e-12)
                assert not bad.any(), f"{fx.name} {key} rows: {bad.sum()} entries beyond the element-wise tolerance"
                cs = z[key + "_colsum"]
                errc = np.abs(val.astype(np.float64).sum(0) - cs).max() / np.abs(cs).max()
                assert errc <= rtol_w, f"{fx.name} {key} colsum: {errc:.3e}"
        elif key.endswith("/s"):
            assert (val != z[key]).sum() <= count_slack, f"{fx.name} {key}: final spikes differ"
        else:
            ref = z[key]
            bad = np.abs(val - ref) > atol_state + rtol_state * np.abs(ref)
            assert not bad.any(), f"{fx.name} {key}: max |d| {np.abs(val - ref).max():.3e} ({bad.sum()} entries)"
    for key, val in counts.items():
        assert np.abs(val.astype(np.int64) - z[key]).sum() <= count_slack, (
            f"{fx.name} {key}: spike counts differ by {np.abs(val.astype(np.int64) - z[key]).sum()}"
        )


def assert_bit_identical(a: Dict[str, np.ndarray], b: Dict[str, np.ndarray], what: str):
    assert a.keys() == b.keys()
    for k in a:
        same = np.array_equal(a[k].view(np.uint32) if a[k].dtype == np.float32 else a[k],
                              b[k].view(np.uint32) if b[k].dtype == np.float32 else b[k])
        if not same:
            diff = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))
            raise AssertionError(f"{what}: {k} differs in {(diff > 0).sum()} entries, max |d| {diff.max():.3e}")
