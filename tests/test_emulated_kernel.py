"""The generic window kernel's REAL CUDA source (bindsnet_b200/csrc/snn_generic.cu, snn_phases.cuh, snn_common.cuh,
snn_api.cu), compiled for the host on a small emulation of the CUDA execution model (tests/emu/: one OS thread per CTA,
cooperatively scheduled fibers per thread, warp / CTA collectives as rendezvous, host atomics for the grid barrier), and
run on the golden cases — bit for bit against the oracle, like the `-m gpu` parity tests do on the B200.

What this tier checks without a GPU: the kernel's work decomposition (sample chunks, learning units, several units per
CTA, several CTAs), indexing, summation orders, the staging / prefetch logic and the barrier protocol.  What it cannot
check: anything that depends on the GPU's memory model or timing — that stays with the `-m gpu` tests."""
import os
import sys

import pytest

import cases
import helpers

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

SKIP = ("dc2015_c2", "dc2015_metric_t40", "dc2015_metric_t250", "conv_mstdp_c4", "conv_mstdp_c4_b128")   # minutes under emulation
SMALL = [c for c in cases.CASES if c not in SKIP]


def _run_emulated(name, env=None, tier=1):
    import emu

    fx = helpers.Fixture(name)
    net, inputs, kw, T = fx.build("cpu")
    net.force_tier = tier      # 1: the generic window kernel, 2: the fused DiehlAndCook2015 window kernel
    helpers.add_spike_monitors(net, T)
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        with emu.EmuBackend() as eb:
            net.run(inputs=inputs, time=T, one_spike_seed=cases.ONE_SPIKE_SEED, **kw)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert eb.err == 0
    assert emu.last_tier == tier, f"window went to tier {emu.last_tier}, not {tier}"
    return helpers.snapshot(net), helpers.spike_counts(net, T)


@pytest.mark.parametrize("name", SMALL)
def test_emulated_generic_kernel_bit_exact_vs_oracle(name):
    s_emu, c_emu = _run_emulated(name)
    _, s_cpu, c_cpu = helpers.run_case_oracle(name)
    helpers.assert_bit_identical(s_emu, s_cpu, f"{name} state (emulated kernel)")
    helpers.assert_bit_identical(c_emu, c_cpu, f"{name} spike counts (emulated kernel)")


@pytest.mark.parametrize("name", ["lif_postpre_batch", "dc2015_onespike", "conv_mstdp", "mstdp_dense"])
@pytest.mark.parametrize("sms", ["1", "7"])
def test_emulated_kernel_is_independent_of_the_grid_size(name, sms):
    """One CTA doing every unit in turn and 14 CTAs sharing them give the same bits: nothing depends on which CTA ran what."""
    s_emu, c_emu = _run_emulated(name, env={"SNN_EMU_SMS": sms})
    _, s_cpu, c_cpu = helpers.run_case_oracle(name)
    helpers.assert_bit_identical(s_emu, s_cpu, f"{name} state (emulated kernel, {sms} SMs)")
    helpers.assert_bit_identical(c_emu, c_cpu, f"{name} spike counts (emulated kernel, {sms} SMs)")


@pytest.mark.parametrize("name", ["lif_postpre_batch", "dc2015_onespike", "conv_mstdp", "mstdp_dense"])
@pytest.mark.parametrize("seed", ["1"])
def test_emulated_kernel_under_random_thread_interleavings(name, seed):
    """SNN_EMU_SHUFFLE: the emulator picks the next thread at random at every synchronisation point instead of round
    robin — the warps of a CTA interleave differently on every run, so a missing __syncthreads / __syncwarp between a
    shared-memory write and its readers shows up as a mismatch.  (A logic-level stand-in for racecheck, not a
    replacement: it knows nothing of the GPU's memory model.)"""
    s_emu, c_emu = _run_emulated(name, env={"SNN_EMU_SHUFFLE": seed, "SNN_EMU_SMS": "2"})
    _, s_cpu, c_cpu = helpers.run_case_oracle(name)
    helpers.assert_bit_identical(s_emu, s_cpu, f"{name} state (emulated kernel, shuffled schedule {seed})")
    helpers.assert_bit_identical(c_emu, c_cpu, f"{name} spike counts (emulated kernel, shuffled schedule {seed})")


@pytest.mark.parametrize("seed", list(range(24)))
def test_emulated_kernel_on_random_networks(seed):
    """The randomly drawn networks of test_oracle_fuzz_vs_reference.py (node kinds x rules x reductions x options x batch
    sizes) through the emulated kernel, bit for bit against the oracle — the CPU twin of tests/test_zz_gpu_fuzz_vs_oracle.py."""
    import emu
    import test_oracle_fuzz_vs_reference as fuzz
    from oracle.oracle import OracleBackend

    spec = fuzz._draw(seed)
    ns = cases.namespace("b200")
    outs = []
    for backend in (emu.EmuBackend, OracleBackend):
        net, x = fuzz._build(ns, spec)
        net.force_tier = 1
        helpers.add_spike_monitors(net, spec["T"])
        with backend() as be:
            net.run(inputs={"X": x}, time=spec["T"])
            assert be.err == 0
        outs.append((helpers.snapshot(net), helpers.spike_counts(net, spec["T"])))
    what = f"seed {seed} {spec['kind']} {spec['rule']} B={spec['B']}"
    helpers.assert_bit_identical(outs[0][0], outs[1][0], what + " state (emulated kernel)")
    helpers.assert_bit_identical(outs[0][1], outs[1][1], what + " spike counts (emulated kernel)")


@pytest.mark.parametrize("shape,clamp,norm,norm_abs", [((784, 160), True, 78.4, 0), ((100, 37), True, None, 1), ((53, 129), False, 5.0, 1),
                                                       ((7, 33), True, 2.0, 0), ((260, 64), True, 11.0, 1)])
def test_emulated_window_combine_bit_exact_vs_oracle(shape, clamp, norm, norm_abs):
    """The multi-GPU window combine (csrc/snn_combine.cuh: W = clamp(W0 + sum dW), normalize(), eight rows in flight per
    warp) under emulation against snn_oracle_delta_apply — out of place and in place (the fused form, with theta)."""
    import ctypes as C

    import numpy as np
    import torch

    import emu
    from oracle import oracle

    g = torch.Generator().manual_seed(11)
    w0 = (0.3 * torch.rand(*shape, generator=g)).contiguous()
    dsum = (0.08 * torch.randn(*shape, generator=g)).contiguous()
    ref = torch.empty_like(w0)
    assert oracle.lib().snn_oracle_delta_apply(ref.data_ptr(), w0.data_ptr(), dsum.data_ptr(), shape[0], shape[1], int(clamp), C.c_float(0.0),
                                               C.c_float(1.0), int(norm is not None), norm_abs, C.c_float(norm or 0.0)) == 0
    L = emu.lib()
    out = torch.full_like(w0, 7.0)
    assert L.snn_b200_delta_apply(out.data_ptr(), w0.data_ptr(), dsum.data_ptr(), shape[0], shape[1], int(clamp), 0.0, 1.0,
                                  int(norm is not None), norm_abs, norm or 0.0, None) == 0
    assert np.array_equal(out.numpy().view(np.uint32), ref.numpy().view(np.uint32))
    inplace = w0.clone()
    th, dth = torch.rand(shape[1], generator=g), torch.rand(shape[1], generator=g)
    th_ref = th + dth
    assert L.snn_b200_delta_apply_fused(inplace.data_ptr(), dsum.data_ptr(), shape[0], shape[1], int(clamp), 0.0, 1.0, int(norm is not None),
                                        norm_abs, norm or 0.0, th.data_ptr(), dth.data_ptr(), shape[1], None) == 0
    assert np.array_equal(inplace.numpy().view(np.uint32), ref.numpy().view(np.uint32))
    assert torch.equal(th, th_ref)


def test_emulated_kernel_large_batch_wide_layer_paths():
    """DiehlAndCook2015 with n = 1100 (35 bit words per sample: two gather blocks, per-sample any-spike flags), B = 70
    (three sample chunks per tile, the eager weight-row prefetch of the learning phase) — the code paths of BASELINE
    configs 2-3 that the small goldens do not reach — through the emulated kernel, bit for bit against the oracle."""
    import torch

    import emu
    from bindsnet_b200.models import DiehlAndCook2015
    from oracle.oracle import OracleBackend

    n, B, T = 1100, 70, 24

    def build():
        torch.manual_seed(5)
        net = DiehlAndCook2015(n_inpt=784, n_neurons=n, batch_size=B, inpt_shape=(1, 28, 28), dt=1.0, nu=(1e-3, 1e-2), norm=78.4,
                               theta_plus=0.05, exc=22.5, inh=120.0)
        g = torch.Generator().manual_seed(9)
        return net, torch.bernoulli(0.12 * torch.ones(T, B, 1, 28, 28), generator=g).byte()

    outs = []
    for backend, tier in ((emu.EmuBackend, 1), (emu.EmuBackend, 2), (OracleBackend, 0)):
        net, x = build()
        net.force_tier = tier
        helpers.add_spike_monitors(net, T)
        with backend() as be:
            net.run({"X": x}, time=T, one_spike_seed=3)
            assert be.err == 0
        outs.append((helpers.snapshot(net), helpers.spike_counts(net, T)))
    assert int(outs[2][1]["L/Ae/count"].sum()) > 100 and int(outs[2][1]["L/Ai/count"].sum()) > 100
    for k, what in ((0, "generic"), (1, "fused")):
        helpers.assert_bit_identical(outs[k][0], outs[2][0], f"state (emulated {what} kernel)")
        helpers.assert_bit_identical(outs[k][1], outs[2][1], f"spike counts (emulated {what} kernel)")


def test_emulated_kernel_monitors_and_state_carry_over():
    """CPU twins of three `-m gpu` tests: spike and voltage recordings step by step, the in-kernel spike counter, state
    carried across windows and a batch-size change — emulated kernel against the oracle."""
    import numpy as np
    import torch

    import emu
    from bindsnet_b200.models import DiehlAndCook2015
    from bindsnet_b200.network.monitors import Monitor, SpikeCounter
    from oracle.oracle import OracleBackend

    g = torch.Generator().manual_seed(5)
    w0 = 0.3 * torch.rand(784, 50, generator=g)
    xa = torch.bernoulli(0.05 * torch.ones(40, 6, 1, 28, 28), generator=g).byte()
    xb = torch.bernoulli(0.05 * torch.ones(30, 3, 1, 28, 28), generator=g).byte()
    outs = []
    for backend, tier in ((emu.EmuBackend, 1), (emu.EmuBackend, 2), (OracleBackend, 0)):
        net = DiehlAndCook2015(n_inpt=784, n_neurons=50, batch_size=6, inpt_shape=(1, 28, 28), inh=120.0)
        net.force_tier = tier
        with torch.no_grad():
            net.connections[("X", "Ae")].w.copy_(w0)
        net.add_monitor(Monitor(net.layers["Ae"], ["s", "v"], time=40), "ae")
        net.add_monitor(SpikeCounter(net.layers["Ae"]), "count")
        rec = {}
        with backend() as be:
            net.run({"X": xa}, time=40, one_spike_seed=1)
            rec["s1"], rec["v1"] = net.monitors["ae"].get("s").clone().numpy(), net.monitors["ae"].get("v").clone().numpy()
            assert torch.equal(net.monitors["count"].get("s"), net.monitors["ae"].get("s").sum(0).to(torch.int32))
            net.run({"X": xa}, time=40, one_spike_seed=2)           # state carried over
            rec["s2"] = net.monitors["ae"].get("s").clone().numpy()
            net.run({"X": xb}, time=30, one_spike_seed=3)           # batch size 6 -> 3: state re-allocated
            assert be.err == 0
        assert net.layers["Ae"].v.shape == (3, 50)
        outs.append((helpers.snapshot(net), rec))
    for q, what in ((0, "generic"), (1, "fused")):
        helpers.assert_bit_identical(outs[q][0], outs[2][0], f"multi-window state (emulated {what} kernel)")
        for k in outs[q][1]:
            assert np.array_equal(outs[q][1][k], outs[2][1][k]), (what, k)
    assert outs[2][1]["s2"].sum() > 0


# ---- the fused DiehlAndCook2015 window kernel (tier 2: the metric's kernel) under emulation ---------------------------
# Its bulk copies (cp.async.bulk + mbarrier) run on the emulator's model of them: the copy happens at issue — the earliest
# moment the hardware could overwrite the destination —, waits yield until the phase has completed.
FUSED_GOLDEN = ["dc2015_multi", "dc2015_onespike", "dc2015_eval", "dc2015_b1_t1", "dc2015_silent"]


@pytest.mark.parametrize("name", FUSED_GOLDEN)
def test_emulated_fused_kernel_bit_exact_vs_oracle(name):
    s_emu, c_emu = _run_emulated(name, tier=2)
    _, s_cpu, c_cpu = helpers.run_case_oracle(name)
    helpers.assert_bit_identical(s_emu, s_cpu, f"{name} state (emulated fused kernel)")
    helpers.assert_bit_identical(c_emu, c_cpu, f"{name} spike counts (emulated fused kernel)")


def _fused_variant(name, env=None, fused_tier=2):
    import emu
    import test_gpu_variants as V
    from oracle.oracle import OracleBackend

    outs = []
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        for backend, tier in ((emu.EmuBackend, fused_tier), (OracleBackend, 0)):
            net, inputs, T = V._build(name, "cpu")
            net.force_tier = tier
            helpers.add_spike_monitors(net, T)
            with backend() as be:
                net.run(inputs=inputs, time=T, one_spike_seed=cases.ONE_SPIKE_SEED)
                assert be.err == 0
            if backend is emu.EmuBackend:
                assert emu.last_tier == fused_tier, f"window went to tier {emu.last_tier}, not {fused_tier}"
            outs.append((helpers.snapshot(net), helpers.spike_counts(net, T)))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert int(outs[1][1]["L/Ae/count"].sum()) > 0
    return outs


def _variant_names():
    import test_gpu_variants as V

    return list(V._variants())


@pytest.mark.parametrize("name", _variant_names())
def test_emulated_fused_kernel_option_variants(name):
    """The option variants of tests/test_gpu_variants.py (WeightDependentPostPre, mean reduction, additive traces, voltage
    bounds, weight decay, batches above 128, multi-spike, dense input slots ...) through the emulated fused kernel."""
    outs = _fused_variant(name)
    helpers.assert_bit_identical(outs[0][0], outs[1][0], f"{name} state (emulated fused kernel)")
    helpers.assert_bit_identical(outs[0][1], outs[1][1], f"{name} spike counts (emulated fused kernel)")


@pytest.mark.parametrize("name", ["wdep_mean_lists", "batch160_lists", "lean_dense_slots", "lean_b40_n100"])
@pytest.mark.parametrize("seed", ["1", "2"])
def test_emulated_fused_kernel_under_random_thread_interleavings(name, seed):
    outs = _fused_variant(name, env={"SNN_EMU_SHUFFLE": seed})
    helpers.assert_bit_identical(outs[0][0], outs[1][0], f"{name} state (emulated fused kernel, shuffled schedule {seed})")
    helpers.assert_bit_identical(outs[0][1], outs[1][1], f"{name} spike counts (emulated fused kernel, shuffled schedule {seed})")


# ---- the column-group fused kernel (tier 3, opt-in) under emulation ---------------------------------------------------
# The kernel compute-sanitizer's racecheck pointed at in round 1 (hazards fixed since): named barriers per column group,
# a bulk-copy ring and a relaxed-polling hand-off between CTAs — shuffled thread schedules are the part that matters here.
@pytest.mark.parametrize("name", FUSED_GOLDEN)
def test_emulated_column_group_kernel_bit_exact_vs_oracle(name):
    s_emu, c_emu = _run_emulated(name, tier=3)
    _, s_cpu, c_cpu = helpers.run_case_oracle(name)
    helpers.assert_bit_identical(s_emu, s_cpu, f"{name} state (emulated column-group kernel)")
    helpers.assert_bit_identical(c_emu, c_cpu, f"{name} spike counts (emulated column-group kernel)")


def _tier3_names():
    import test_gpu_variants as V

    return sorted(V.TIER3)


@pytest.mark.parametrize("name", _tier3_names())
@pytest.mark.parametrize("seed", [None, "1", "2"])
def test_emulated_column_group_kernel_variants_and_interleavings(name, seed):
    """The shapes tests/test_gpu_variants.py sends through tier 3 on the GPU, here in program order (seed None) and under
    two shuffled thread schedules."""
    outs = _fused_variant(name, env={"SNN_EMU_SHUFFLE": seed} if seed else None, fused_tier=3)
    helpers.assert_bit_identical(outs[0][0], outs[1][0], f"{name} state (emulated column-group kernel, schedule {seed})")
    helpers.assert_bit_identical(outs[0][1], outs[1][1], f"{name} spike counts (emulated column-group kernel, schedule {seed})")


def test_emulated_delta_window_and_combine():
    """CPU twin of test_gpu_ops.test_delta_window_writes_the_change_and_leaves_the_weights: the fused kernel's delta window
    (snn_run_opts_t.delta_w / delta_theta: the multi-GPU combine's input) writes W_end - W_start / theta_end - theta_start and
    leaves W / theta alone; the in-place combine equals the snapshot-based one — all under emulation."""
    import torch

    import emu
    from bindsnet_b200.models import DiehlAndCook2015

    def make():
        torch.manual_seed(3)
        net = DiehlAndCook2015(n_inpt=784, n_neurons=96, batch_size=8, inpt_shape=(1, 28, 28), norm=78.4, theta_plus=0.05)
        net.force_tier = 2
        return net

    x = torch.bernoulli(0.04 * torch.ones(80, 8, 1, 28, 28), generator=torch.Generator().manual_seed(5)).byte()
    a, b = make(), make()
    w0, th0 = a.connections[("X", "Ae")].w.detach().clone(), a.layers["Ae"].theta.clone()
    wb, thb = b.connections[("X", "Ae")].w.detach(), b.layers["Ae"].theta
    flat = torch.full((wb.numel() + thb.numel(),), float("nan"))
    dw, dth = flat[:wb.numel()].view_as(wb), flat[wb.numel():]
    with emu.EmuBackend() as be:
        a.run({"X": x}, time=80, one_spike_seed=9, b200_normalize=False)
        b.run({"X": x}, time=80, one_spike_seed=9, b200_normalize=False, b200_delta=(dw, dth))
        assert be.err == 0
    assert torch.equal(wb, w0) and torch.equal(thb, th0), "a delta window must not touch W / theta"
    wa, tha = a.connections[("X", "Ae")].w.detach(), a.layers["Ae"].theta
    assert torch.equal(dw, wa - w0) and torch.equal(dth, tha - th0)
    assert float(dw.abs().sum()) > 0 and float(dth.abs().sum()) > 0, "nothing learned: nothing tested"
    for lname in ("Ae", "Ai"):
        assert torch.equal(a.layers[lname].v, b.layers[lname].v) and torch.equal(a.layers[lname].s, b.layers[lname].s)
    L = emu.lib()
    ref_w = torch.empty_like(w0)
    two = (2.0 * flat).contiguous()
    assert L.snn_b200_delta_apply(ref_w.data_ptr(), w0.data_ptr(), two.data_ptr(), 784, 96, 1, 0.0, 1.0, 1, 0, 78.4, None) == 0
    thc = thb.clone()
    assert L.snn_b200_delta_apply_fused(wb.data_ptr(), two.data_ptr(), 784, 96, 1, 0.0, 1.0, 1, 0, 78.4, thc.data_ptr(),
                                        two[wb.numel():].data_ptr(), 96, None) == 0
    assert torch.equal(wb, ref_w) and torch.equal(thc, th0 + 2.0 * dth)


# ---- the small kernels of the library (snn_ops.cu, snn_encode.cu, snn_readout.cu) under emulation -----------------------

def test_emulated_single_operator_entry_points():
    """CPU twin of test_gpu_parity.test_single_operator_entry_points and test_gpu_ops.test_conv2d_single_operators...:
    Connection.compute / update / normalize and Conv2dConnection.compute / normalize through the emulated kernels equal the
    oracle's bit for bit."""
    import numpy as np
    import torch

    import emu
    from bindsnet_b200.learning import PostPre
    from bindsnet_b200.network import nodes, topology
    from oracle.oracle import OracleBackend

    res = []
    for backend in (emu.EmuBackend, OracleBackend):
        g = torch.Generator().manual_seed(9)
        X = nodes.Input(n=70, traces=True); Y = nodes.LIFNodes(n=45, traces=True)
        Cn = topology.Connection(X, Y, w=torch.rand(70, 45, generator=g), b=torch.rand(45, generator=g),
                                 update_rule=PostPre, nu=(1e-2, 2e-2), reduction=torch.sum, wmin=0.0, wmax=1.0, norm=10.0)
        for l in (X, Y):
            l.compute_decays(1.0); l.set_batch_size(5)
        X.s = torch.bernoulli(0.3 * torch.ones(5, 70), generator=g).bool(); X.x = torch.rand(5, 70, generator=g)
        Y.s = torch.bernoulli(0.3 * torch.ones(5, 45), generator=g).bool(); Y.x = torch.rand(5, 45, generator=g)
        Xc = nodes.Input(shape=[2, 11, 9], traces=True); H = nodes.LIFNodes(shape=[5, 6, 5], traces=True)
        cc = topology.Conv2dConnection(Xc, H, kernel_size=(3, 3), stride=2, padding=1, wmin=-1.0, wmax=1.0, norm=0.4,
                                       w=torch.rand(5, 2, 3, 3, generator=g) - 0.3, b=torch.rand(5, generator=g))
        sc = torch.bernoulli(0.3 * torch.ones(7, 2, 11, 9), generator=g).byte()
        with backend():
            out = Cn.compute(X.s)
            Cn.update(learning=True)
            w_after_update = Cn.w.detach().clone()
            Cn.normalize()
            outc = cc.compute(sc)
            cc.normalize()
        res.append([t.detach().numpy().copy() for t in (out, w_after_update, Cn.w, outc, cc.w)])
    for a, b in zip(*res):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert res[0][0].shape == (5, 45) and res[0][3].shape == (7, 5, 6, 5)


def test_emulated_readout_kernels_match_the_torch_formulas():
    """CPU twin of test_gpu_ops.test_readout_kernels_match_the_torch_formulas: snn_b200_assign_labels / snn_b200_predict
    under emulation against bindsnet_b200.evaluation's CPU formulas (pinned against the live reference elsewhere)."""
    import torch

    import emu
    from bindsnet_b200 import _backend, evaluation as ev

    g = torch.Generator().manual_seed(8)
    S, n, L = 64, 300, 10
    counts = torch.poisson(3.0 * torch.rand(S, n, generator=g), generator=g).int().contiguous()
    labels = torch.randint(0, L, (S,), generator=g)
    a_c, p_c, r_c = ev.assign_labels(counts, labels, L)
    with emu.EmuBackend():
        rates = torch.zeros(n, L); prop = torch.empty(n, L); asg = torch.empty(n, dtype=torch.int64)
        _backend.assign_labels(counts, labels.contiguous(), L, 1.0, rates, prop, asg)
        assert torch.equal(asg, a_c) and torch.allclose(prop, p_c, atol=1e-6) and torch.allclose(rates, r_c, atol=1e-5)
        rates2 = rates.clone(); prop2 = torch.empty(n, L); asg2 = torch.empty(n, dtype=torch.int64)
        _backend.assign_labels(counts.flip(0).contiguous(), labels.contiguous(), L, 0.8, rates2, prop2, asg2)
        a2_c, p2_c, r2_c = ev.assign_labels(counts.flip(0), labels, L, rates=r_c.clone(), alpha=0.8)
        assert torch.equal(asg2, a2_c) and torch.allclose(rates2, r2_c, atol=1e-5)
        pred = torch.empty(S, dtype=torch.int64)
        _backend.predict(counts, asg, None, L, pred)
        assert torch.equal(pred, ev.all_activity(counts, a_c, L))
        predw = torch.empty(S, dtype=torch.int64)
        _backend.predict(counts, asg, prop.contiguous(), L, predw)
        assert (predw != ev.proportion_weighting(counts, a_c, p_c, L)).sum() <= 1   # weighted float sums: a near-tie may fall the other way


def test_emulated_encoders():
    """The on-device encoders (csrc/snn_encode.cu) under emulation: Bernoulli rates, Poisson rates against the CPU
    restatement of the reference's encoder, and the counter-based stream property (a train depends on seed and element only)."""
    import numpy as np
    import torch

    import emu
    from bindsnet_b200 import _backend
    from bindsnet_b200.encoding import poisson

    with emu.EmuBackend():
        p = torch.tensor([0.0, 0.05, 0.3, 1.0]).repeat_interleave(1500).contiguous()
        out = torch.empty(100, p.numel(), dtype=torch.uint8)
        _backend.encode_bernoulli(p, 100, 3, out)
        got = out.float().view(100, 4, 1500).mean(dim=(0, 2)).numpy()
        assert got[0] == 0.0 and got[3] == 1.0
        assert abs(got[1] - 0.05) < 5 * np.sqrt(0.05 * 0.95 / 1.5e5) and abs(got[2] - 0.3) < 5 * np.sqrt(0.3 * 0.7 / 1.5e5)
        rates = torch.tensor([0.0, 10.0, 64.0, 128.0]).repeat_interleave(1200).contiguous()
        T = 200
        dev = torch.empty(T, rates.numel(), dtype=torch.uint8)
        _backend.encode_poisson(rates, T, 1.0, 77, dev)
        again = torch.empty_like(dev)
        _backend.encode_poisson(rates, T, 1.0, 77, again)
        other = torch.empty_like(dev)
        _backend.encode_poisson(rates, T, 1.0, 78, other)
        assert torch.equal(dev, again) and not torch.equal(dev, other)
        first = torch.empty(T, 1200, dtype=torch.uint8)                 # the first 1200 elements alone: the same trains
        _backend.encode_poisson(rates[:1200].contiguous(), T, 1.0, 77, first)
        assert torch.equal(first, dev[:, :1200])
    torch.manual_seed(5)
    ref = poisson(rates, time=T, dt=1.0).numpy()                        # CPU restatement of encodings.py:99-156
    d = dev.numpy()
    for k, r in enumerate([0.0, 10.0, 64.0, 128.0]):
        a, b = ref[:, k * 1200:(k + 1) * 1200].sum(0).astype(np.float64), d[:, k * 1200:(k + 1) * 1200].sum(0).astype(np.float64)
        if r == 0.0:
            assert a.sum() == 0 and b.sum() == 0
            continue
        se = np.sqrt(a.var() / 1200 + b.var() / 1200) + 1e-9
        assert abs(a.mean() - b.mean()) < 5 * se + 1e-3, f"{r} Hz: mean count {a.mean():.4f} vs {b.mean():.4f}"


def test_emulated_scripted_tier():
    """The scripted tier (user-defined Nodes / LearningRule subclasses stepped from Python, built-in pieces on their
    single-operator kernels) with those kernels under emulation: same result as the built-in network it restates."""
    import torch

    import emu
    import test_scripted_tier as ST

    g = torch.Generator().manual_seed(2)
    w0 = 0.9 * torch.rand(50, 30, generator=g)
    x = torch.bernoulli(0.15 * torch.ones(60, 3, 50), generator=g).byte()
    a, b = ST._net(True, w0), ST._net(False, w0)
    assert a._scripted_required() and not b._scripted_required()
    with emu.EmuBackend() as be:
        a.run({"X": x}, time=60)
        b.run({"X": x}, time=60)
        assert be.err == 0
    sa, sb = a.monitors["Y"].get("s"), b.monitors["Y"].get("s")
    assert int(sb.sum()) > 20 and torch.equal(sa, sb)
    wa, wb = a.connections[("X", "Y")].w, b.connections[("X", "Y")].w
    assert float((wa - wb).abs().max() / wb.abs().max()) < 1e-5


@pytest.mark.parametrize("name,tier", [("dc2015_c2", 2), ("dc2015_c2", 1), ("dc2015_metric_t40", 2), ("dc2015_metric_t40", 1),
                                       ("dc2015_metric_t250", 2), ("dc2015_c2", 3), ("dc2015_metric_t40", 3),
                                       ("conv_mstdp_c4", 1), ("conv_mstdp_c4_b128", 1)])
def test_emulated_kernels_at_the_baseline_configurations(name, tier):
    """BASELINE.json config 2 (n = 400, B = 32, T = 250) and the metric configuration (n = 1600, B = 128; T = 40 and the full
    250-step window) through the emulated kernels — tier 2 is the kernel the metric is quoted on —, and config 4 (the
    convolutional MSTDP network, B = 32 and B = 128) through the generic kernel, bit for bit against the oracle, which the golden fixtures pin against the live reference at exactly these shapes."""
    s_emu, c_emu = _run_emulated(name, tier=tier)
    _, s_cpu, c_cpu = helpers.run_case_oracle(name)
    helpers.assert_bit_identical(s_emu, s_cpu, f"{name} state (emulated kernel, tier {tier})")
    helpers.assert_bit_identical(c_emu, c_cpu, f"{name} spike counts (emulated kernel, tier {tier})")
