"""Parity tests proper (need a B200): the CUDA path, called through the C ABI exactly as
Network.run calls it, against (a) the oracle on the same seeded inputs — bit-exact, spike
rasters included — and (b) the golden fixtures produced by the live reference — within the
north_star's tolerances (weights 1e-4 relative, fp32 state tolerance, spike counts)."""
import numpy as np
import pytest
import torch

import cases
import helpers

pytestmark = pytest.mark.gpu

ALL = list(cases.CASES)
# cases whose CUDA path is exercised by test_gpu_mstdp_conv.py (generic tier only)
NEW_ROWS = ("mstdp_dense", "conv_mstdp", "conv_stride_norm", "conv_mstdp_c4", "conv_mstdp_c4_b128", "mstdp_mean_decay", "conv_bias_stride")
ALL = [c for c in ALL if c not in NEW_ROWS]
SMALL = [c for c in ALL if c not in ("dc2015_c2", "dc2015_metric_t40", "dc2015_metric_t250")]


def run_case_gpu(name, tier=1):
    fx = helpers.Fixture(name)
    net, inputs, kw, T = fx.build("cuda")
    net.force_tier = tier
    helpers.add_spike_monitors(net, T, device="cuda")
    kw = {k: ({l: v.cuda() for l, v in d.items()} if isinstance(d, dict) else d) for k, d in kw.items()}
    net.run(inputs={k: v.cuda() for k, v in inputs.items()}, time=T, one_spike_seed=cases.ONE_SPIKE_SEED, **kw)
    net.check_errors()
    return fx, helpers.snapshot(net), helpers.spike_counts(net, T)


@pytest.mark.parametrize("name", ALL)
def test_generic_kernel_bit_exact_vs_oracle(name):
    _, s_gpu, c_gpu = run_case_gpu(name, tier=1)
    _, s_cpu, c_cpu = helpers.run_case_oracle(name)
    helpers.assert_bit_identical(s_gpu, s_cpu, f"{name} state")
    helpers.assert_bit_identical(c_gpu, c_cpu, f"{name} spike counts")


DC = [c for c in ALL if c.startswith("dc2015") and c != "dc2015v2"]


@pytest.mark.parametrize("tier", [0, 3])
@pytest.mark.parametrize("name", DC)
def test_fused_kernel_selected_and_bit_exact_vs_oracle(name, tier):
    """tier=0 (auto) must pick the grid-barrier fused DiehlAndCook2015 kernel (tier 2, the faster one) for these
    graphs, tier=3 the column-group kernel; both must agree with the oracle bit for bit."""
    from bindsnet_b200 import _backend

    _, s_gpu, c_gpu = run_case_gpu(name, tier=tier)
    _, s_cpu, c_cpu = helpers.run_case_oracle(name)
    helpers.assert_bit_identical(s_gpu, s_cpu, f"{name} state (fused, tier {tier})")
    helpers.assert_bit_identical(c_gpu, c_cpu, f"{name} spike counts (fused, tier {tier})")
    assert _backend.last_tier == (2 if tier == 0 else 3), "tier selection did not pick the expected fused kernel"


@pytest.mark.parametrize("name", DC)
def test_fused_kernel_vs_reference_golden(name):
    fx, state, counts = run_case_gpu(name, tier=0)
    helpers.assert_close_to_golden(fx, state, counts)


@pytest.mark.parametrize("name", ALL)
def test_generic_kernel_vs_reference_golden(name):
    fx, state, counts = run_case_gpu(name, tier=1)
    helpers.assert_close_to_golden(fx, state, counts)


def test_full_rasters_bit_exact_and_voltage_monitor():
    """Monitors: [T,B,n] spike and voltage recordings equal the oracle's, step by step."""
    from bindsnet_b200.network.monitors import Monitor
    from oracle.oracle import OracleBackend

    outs = []
    for dev in ("cuda", "cpu"):
        fx = helpers.Fixture("dc2015_onespike")
        net, inputs, kw, T = fx.build(dev)
        net.add_monitor(Monitor(net.layers["Ae"], ["s", "v"], time=T, device=dev), "ae")
        net.add_monitor(Monitor(net.layers["Ai"], ["s"], time=T, device=dev), "ai")
        x = {k: v.to(dev) for k, v in inputs.items()}
        if dev == "cpu":
            with OracleBackend():
                net.run(inputs=x, time=T, one_spike_seed=3)
        else:
            net.run(inputs=x, time=T, one_spike_seed=3)
            net.check_errors()
        outs.append({k: net.monitors[m].get(k).cpu().numpy() for m, k in (("ae", "s"), ("ae", "v"), ("ai", "s"))})
    for k in outs[0]:
        assert outs[0][k].shape == outs[1][k].shape and outs[0][k].shape[0] == 150
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_two_windows_without_reset_and_batch_change():
    """State persists across run() calls; a batch-size change re-allocates it (network.py:342-353)."""
    from bindsnet_b200.models import DiehlAndCook2015
    from oracle.oracle import OracleBackend

    g = torch.Generator().manual_seed(5)
    w0 = 0.3 * torch.rand(784, 50, generator=g)
    xa = torch.bernoulli(0.05 * torch.ones(40, 6, 1, 28, 28), generator=g).byte()
    xb = torch.bernoulli(0.05 * torch.ones(30, 3, 1, 28, 28), generator=g).byte()
    snaps = []
    for dev in ("cuda", "cpu"):
        net = DiehlAndCook2015(n_inpt=784, n_neurons=50, batch_size=6, inpt_shape=(1, 28, 28), inh=120.0)
        with torch.no_grad():
            net.connections[("X", "Ae")].w.copy_(w0)
        net.to(dev)
        ctx = OracleBackend() if dev == "cpu" else None
        if ctx: ctx.__enter__()
        net.run({"X": xa.to(dev)}, time=40, one_spike_seed=1)
        net.run({"X": xa.to(dev)}, time=40, one_spike_seed=2)
        net.run({"X": xb.to(dev)}, time=30, one_spike_seed=3)
        if ctx: ctx.__exit__(None, None, None)
        else: net.check_errors()
        assert net.layers["Ae"].v.shape == (3, 50)
        snaps.append(helpers.snapshot(net))
    helpers.assert_bit_identical(snaps[0], snaps[1], "multi-window")


def test_single_operator_entry_points():
    """Connection.compute / update / normalize through the C ABI equal the oracle's."""
    from bindsnet_b200.network import nodes, topology
    from bindsnet_b200.learning import PostPre
    from oracle.oracle import OracleBackend

    res = []
    for dev in ("cuda", "cpu"):
        g = torch.Generator().manual_seed(9)
        X = nodes.Input(n=70, traces=True); Y = nodes.LIFNodes(n=45, traces=True)
        C = topology.Connection(X, Y, w=torch.rand(70, 45, generator=g), b=torch.rand(45, generator=g),
                                update_rule=PostPre, nu=(1e-2, 2e-2), reduction=torch.sum, wmin=0.0, wmax=1.0, norm=10.0)
        for l in (X, Y):
            l.compute_decays(1.0); l.set_batch_size(5)
        X.s = torch.bernoulli(0.3 * torch.ones(5, 70), generator=g).bool(); X.x = torch.rand(5, 70, generator=g)
        Y.s = torch.bernoulli(0.3 * torch.ones(5, 45), generator=g).bool(); Y.x = torch.rand(5, 45, generator=g)
        for m in (X, Y, C):
            m.to(dev)
        ctx = OracleBackend() if dev == "cpu" else None
        if ctx: ctx.__enter__()
        out = C.compute(X.s)
        C.update(learning=True)
        w_after_update = C.w.detach().clone()
        C.normalize()
        if ctx: ctx.__exit__(None, None, None)
        res.append([t.cpu().numpy() for t in (out, w_after_update, C.w.detach())])
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    ref = (res[1][0])
    assert ref.shape == (5, 45)


def test_nonbinary_input_is_reported():
    from bindsnet_b200 import _backend
    from bindsnet_b200.models import TwoLayerNetwork

    net = TwoLayerNetwork(n_inpt=16, n_neurons=8).to("cuda")
    x = torch.zeros(5, 1, 16, dtype=torch.uint8, device="cuda"); x[2, 0, 3] = 7
    net.run({"X": x}, time=5)
    with pytest.raises(_backend.BackendError):
        net.check_errors()


@pytest.mark.parametrize("tier", [0, 1])
def test_spike_counter_on_gpu(tier):
    """In-kernel spike counts (fused: registers, generic: global) == raster summed over time."""
    from bindsnet_b200.models import DiehlAndCook2015
    from bindsnet_b200.network.monitors import Monitor, SpikeCounter

    g = torch.Generator().manual_seed(8)
    net = DiehlAndCook2015(n_inpt=784, n_neurons=96, batch_size=8, inpt_shape=(1, 28, 28), inh=120.0).to("cuda")
    net.force_tier = tier
    net.add_monitor(Monitor(net.layers["Ae"], ["s"], time=80, device="cuda"), "raster")
    net.add_monitor(SpikeCounter(net.layers["Ae"]), "count")
    net.add_monitor(SpikeCounter(net.layers["Ai"]), "count_i")
    net.add_monitor(Monitor(net.layers["Ai"], ["s"], time=80, device="cuda"), "raster_i")
    for _ in range(2):
        x = torch.bernoulli(0.06 * torch.ones(80, 8, 1, 28, 28), generator=g).byte().cuda()
        net.run({"X": x}, time=80, one_spike_seed=2)
        net.check_errors()
        assert torch.equal(net.monitors["count"].get("s"), net.monitors["raster"].get("s").sum(0).to(torch.int32))
        assert torch.equal(net.monitors["count_i"].get("s"), net.monitors["raster_i"].get("s").sum(0).to(torch.int32))
        assert net.monitors["count"].get("s").sum() > 0
