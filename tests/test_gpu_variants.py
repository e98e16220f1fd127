"""Fused DiehlAndCook2015 kernel: option coverage beyond the golden cases (needs a B200).

The golden cases pin the default configuration against the live reference; these variants switch
on the rarely used options of the same graph — WeightDependentPostPre, mean reduction, additive
traces, voltage lower bounds, batches above 128 (8-word sample masks), weight decay — and compare
the CUDA path with the oracle bit for bit (state, weights, spike counts), like
test_gpu_parity.test_fused_kernel_selected_and_bit_exact_vs_oracle does for the golden cases."""
import numpy as np
import pytest
import torch

import cases
import helpers

pytestmark = pytest.mark.gpu


def _graph(ns, n, B, *, rule, reduction, nu=(1e-4, 1e-2), additive=False, lbound=None, weight_decay=0.0,
           one_spike=True, exc=22.5, inh=120.0, norm=78.4, w_seed=7, tiny=0.0, huge=0.0):
    """The DiehlAndCook2015 wiring (models.py:94-244) from classic Connection objects, so that the
    learning rule and its options can be chosen freely."""
    net = ns.Network(dt=1.0, batch_size=B)
    X = ns.nodes.Input(n=784, shape=(1, 28, 28), traces=True, tc_trace=20.0, traces_additive=additive)
    E = ns.nodes.DiehlAndCookNodes(n=n, traces=True, rest=-65.0, reset=-60.0, thresh=-52.0, refrac=5, tc_decay=100.0,
                                   tc_trace=20.0, theta_plus=0.05, tc_theta_decay=1e7, traces_additive=additive,
                                   lbound=lbound, one_spike=one_spike)
    I = ns.nodes.LIFNodes(n=n, traces=False, rest=-60.0, reset=-45.0, thresh=-40.0, tc_decay=10.0, refrac=2,
                          lbound=lbound)
    w = cases._w((784, n), w_seed, 0.3)
    if tiny:   # a fraction of the weights next to wmin: the pre term drives them below it, the clamp becomes active
        g = torch.Generator().manual_seed(w_seed + 1)
        w = torch.where(torch.rand(w.shape, generator=g) < tiny, w * 1e-3, w)
    if huge:   # ... and next to wmax: the post term drives them above it
        g = torch.Generator().manual_seed(w_seed + 2)
        w = torch.where(torch.rand(w.shape, generator=g) < huge, 1.0 - w * 1e-2, w)
    cxe = ns.topology.Connection(source=X, target=E, w=w, update_rule=rule, nu=nu, reduction=reduction, wmin=0.0, wmax=1.0,
                                 norm=norm, weight_decay=weight_decay)
    cei = ns.topology.Connection(source=E, target=I, w=exc * torch.diag(torch.ones(n)), wmin=0.0, wmax=exc)
    cie = ns.topology.Connection(source=I, target=E, w=-inh * (torch.ones(n, n) - torch.diag(torch.ones(n))), wmin=-inh, wmax=0.0)
    net.add_layer(X, "X"); net.add_layer(E, "Ae"); net.add_layer(I, "Ai")
    net.add_connection(cxe, "X", "Ae"); net.add_connection(cei, "Ae", "Ai"); net.add_connection(cie, "Ai", "Ae")
    return net


def _variants():
    L = cases.namespace("b200").learning
    return {
        # name: (graph kwargs, B, n, T, input kind)
        "wdep_mean_lists": (dict(rule=L.WeightDependentPostPre, reduction=torch.mean, nu=(1e-2, 5e-2)), 16, 64, 90, "poisson"),
        "wdep_sum_dense": (dict(rule=L.WeightDependentPostPre, reduction=torch.sum, nu=(2e-3, 1e-2)), 8, 40, 70, "bernoulli"),
        "postpre_mean": (dict(rule=L.PostPre, reduction=torch.mean, nu=(1e-3, 1e-1)), 12, 52, 80, "poisson"),
        "additive_traces": (dict(rule=L.PostPre, reduction=torch.sum, additive=True), 9, 36, 70, "poisson"),
        "lbound_decay": (dict(rule=L.PostPre, reduction=torch.sum, lbound=-70.0, weight_decay=1e-3), 6, 28, 60, "poisson"),
        "batch160_lists": (dict(rule=L.PostPre, reduction=torch.sum), 160, 48, 50, "poisson"),
        "multi_spike_lists": (dict(rule=L.PostPre, reduction=torch.sum, one_spike=False, inh=17.5), 24, 64, 80, "poisson"),
        # the lean option set also runs on the column-group kernel (tier 3): odd batch / tile shapes, dense input slots
        "lean_b40_n100": (dict(rule=L.PostPre, reduction=torch.sum), 40, 100, 70, "poisson"),
        "lean_b128_n600": (dict(rule=L.PostPre, reduction=torch.sum, nu=(1e-3, 5e-2)), 128, 600, 40, "poisson"),
        "lean_dense_slots": (dict(rule=L.PostPre, reduction=torch.sum, nu=(2e-3, 1e-2)), 8, 40, 70, "bernoulli"),
        "lean_no_post": (dict(rule=L.PostPre, reduction=torch.sum, nu=(1e-3, 0.0)), 16, 64, 60, "poisson"),
        "lean_weak_inh": (dict(rule=L.PostPre, reduction=torch.sum, inh=3.0, exc=30.0), 16, 64, 90, "poisson"),
        # weights at the lower bound while pre and post term meet on one row: the reference clamps once, after both
        "lean_clamped_rows": (dict(rule=L.PostPre, reduction=torch.sum, nu=(2e-3, 1e-2), one_spike=False, inh=3.0, tiny=0.3), 12, 64, 40,
                              "poisson"),
        "clamped_rows_wdep": (dict(rule=L.WeightDependentPostPre, reduction=torch.sum, nu=(2e-3, 1e-2), one_spike=False, inh=3.0, tiny=0.3),
                              12, 64, 40, "poisson"),
    }

TIER3 = {"multi_spike_lists", "lean_b40_n100", "lean_b128_n600", "lean_dense_slots", "lean_no_post", "lean_weak_inh", "lean_clamped_rows"}


def _build(name, device):
    ns = cases.namespace("b200")
    kw, B, n, T, kind = _variants()[name]
    torch.manual_seed(99)
    net = _graph(ns, n, B, **kw)
    seed = 1000 + sum(map(ord, name))
    if kind == "poisson":
        x = cases._poisson_inputs(ns, T, B, (1, 28, 28), seed)
    else:
        x = cases._bernoulli_inputs(T, B, (1, 28, 28), 0.05, seed)
    if device != "cpu":
        net.to(device)
    return net, {"X": x}, T


@pytest.mark.parametrize("name", list(_variants()))
def test_fused_variant_bit_exact_vs_oracle(name):
    from bindsnet_b200 import _backend
    from oracle.oracle import OracleBackend

    net, inputs, T = _build(name, "cuda")
    helpers.add_spike_monitors(net, T, device="cuda")
    net.run(inputs={k: v.cuda() for k, v in inputs.items()}, time=T, one_spike_seed=cases.ONE_SPIKE_SEED)
    net.check_errors()
    assert _backend.last_tier == 2, "the fused kernel was not selected for this graph"
    s_gpu, c_gpu = helpers.snapshot(net), helpers.spike_counts(net, T)

    ref, inputs, T = _build(name, "cpu")
    helpers.add_spike_monitors(ref, T)
    with OracleBackend() as ob:
        ref.run(inputs=inputs, time=T, one_spike_seed=cases.ONE_SPIKE_SEED)
        assert ob.err == 0
    s_cpu, c_cpu = helpers.snapshot(ref), helpers.spike_counts(ref, T)
    assert sum(int(v.sum()) for k, v in c_cpu.items() if k.endswith("Ae/count")) > 0, "variant produced no Ae spikes: nothing tested"
    helpers.assert_bit_identical(s_gpu, s_cpu, f"{name} state (fused)")
    helpers.assert_bit_identical(c_gpu, c_cpu, f"{name} spike counts (fused)")

    # same network through the generic kernel (and the column-group fused kernel where it matches)
    for tier in ((1, 3) if name in TIER3 else (1,)):
        net2, inputs, T = _build(name, "cuda")
        net2.force_tier = tier
        helpers.add_spike_monitors(net2, T, device="cuda")
        net2.run(inputs={k: v.cuda() for k, v in inputs.items()}, time=T, one_spike_seed=cases.ONE_SPIKE_SEED)
        net2.check_errors()
        assert _backend.last_tier == tier
        helpers.assert_bit_identical(helpers.snapshot(net2), s_cpu, f"{name} state (tier {tier})")
