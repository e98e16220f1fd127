"""GPU twins (need a B200) of the host-side additions made after the round's GPU time was spent: ``reward_fn``,
per-connection ``a_plus`` / ``a_minus`` and per-step index clamps, ``IncreasingInhibitionNetwork`` and
``LocallyConnectedNetwork`` — the same runs as the CPU tests (which compare with the live reference and drive the
kernels' CUDA sources on the emulation), here on the kernels against the oracle, bit for bit.  No kernel changed for
them; the file runs last (name) because its first execution is the driver's."""
import pytest

import cases
import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rule", ["MSTDP", "MSTDPET"])
def test_reward_fn_on_the_kernels_bit_exact_vs_oracle(rule):
    import test_reward_fn as t

    t._kernel_vs_oracle(rule, None, lambda x: x.cuda())


def _pair_of_runs(make, T, run_kw):
    from oracle.oracle import OracleBackend

    gpu = make()
    net, x = gpu[0], gpu[1]
    net.to("cuda")
    helpers.add_spike_monitors(net, T, device="cuda")
    net.run(inputs={"X": x.cuda()}, time=T, **run_kw(gpu))
    net.check_errors()
    cpu = make()
    ref, x2 = cpu[0], cpu[1]
    helpers.add_spike_monitors(ref, T)
    with OracleBackend() as ob:
        ref.run(inputs={"X": x2}, time=T, **run_kw(cpu))
        assert ob.err == 0
    return (helpers.snapshot(net), helpers.spike_counts(net, T)), (helpers.snapshot(ref), helpers.spike_counts(ref, T))


@pytest.mark.parametrize("which", ["dicts", "clamps"])
def test_run_kwargs_on_the_kernels_bit_exact_vs_oracle(which):
    import test_run_kwargs_live as t

    ns = cases.namespace("b200")
    if which == "dicts":
        a, b = _pair_of_runs(lambda: t._two_mstdp(ns), t.T, lambda made: t.KW)
    else:
        a, b = _pair_of_runs(lambda: t._clamped(ns), t.T, lambda made: made[2])
    helpers.assert_bit_identical(a[0], b[0], f"{which} state")
    helpers.assert_bit_identical(a[1], b[1], f"{which} spike counts")


@pytest.mark.parametrize("which", ["increasing", "local"])
def test_canned_models_on_the_kernels_bit_exact_vs_oracle(which):
    import test_models_live as t

    ns = cases.namespace("b200")
    a, b = _pair_of_runs(lambda: t._make(ns, which), t.T, lambda made: {"one_spike_seed": t.SEED})
    assert int(b[1]["L/Y/count"].sum()) > 10
    helpers.assert_bit_identical(a[0], b[0], f"{which} state")
    helpers.assert_bit_identical(a[1], b[1], f"{which} spike counts")


def test_network_monitor_on_the_kernels_bit_exact_vs_oracle():
    """NetworkMonitor makes the run step-wise (one-step windows, the end-of-run normalize as a single operator after
    the last record): recordings on the device equal the oracle's."""
    import torch

    import test_network_monitor as t
    from oracle.oracle import OracleBackend

    ns = cases.namespace("b200")
    net, x = t._net(ns)
    net.to("cuda")
    mon = ns.monitors.NetworkMonitor(net)
    net.add_monitor(mon, "all")
    net.run(inputs={"X": x.cuda()}, time=t.T)
    net.check_errors()
    ref = t._record(ns, None, OracleBackend)
    for key in ref.get():
        for v in ref.get()[key]:
            assert torch.equal(mon.get()[key][v].cpu(), ref.get()[key][v]), (key, v)


def test_srm0_nodes_on_the_device():
    """SRM0Nodes' spike draw is torch's CUDA generator (no oracle can follow it): the window must run on the device
    through the scripted tier — built-in PostPre update and normalize on their kernels with the host population as
    the target — and behave like the CPU run in distribution (rates within a wide band, weights inside their bounds,
    columns normalised)."""
    import torch

    import test_srm0_live as t
    from oracle.oracle import OracleBackend

    ns = cases.namespace("b200")
    net, x = t._net(ns)
    net.to("cuda")
    torch.manual_seed(7)
    net.run(inputs={"X": x.cuda()}, time=t.T)
    net.check_errors()
    Y, w = net.layers["Y"], net.connections[("X", "Y")].w
    assert Y.s.is_cuda and Y.v.is_cuda and w.is_cuda
    cpu, x2 = t._net(ns)
    torch.manual_seed(7)
    with OracleBackend():
        cpu.run(inputs={"X": x2}, time=t.T)
    assert float(w.min()) >= 0.0 and float(w.max()) <= 3.0
    assert torch.allclose(w.sum(0).cpu(), torch.full((15,), 40.0), rtol=1e-4)
    assert float(Y.summed.sum()) > 0 and abs(float(Y.summed.mean()) / float(cpu.layers["Y"].summed.mean()) - 1.0) < 0.5
