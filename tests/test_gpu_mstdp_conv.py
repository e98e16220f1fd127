"""SURVEY.md §8a rows A11-A13 on the GPU (needs a B200): Conv2dConnection.compute / normalize and
MSTDP (dense and conv2d) run in the generic window kernel.  Checked like every other path: bit-exact
against the oracle (state, weights, rule state, spike counts) and within the north_star's tolerances
against the goldens produced by the live reference."""
import numpy as np
import pytest
import torch

import cases
import helpers

pytestmark = pytest.mark.gpu

NEW = ["mstdp_dense", "conv_mstdp", "conv_stride_norm", "conv_mstdp_c4", "conv_mstdp_c4_b128", "mstdp_mean_decay", "conv_bias_stride"]


def _rule_state(net):
    out = {}
    for (s, t), conn in net.connections.items():
        rule = getattr(conn, "update_rule", None)
        for name in ("p_plus", "p_minus"):
            v = getattr(rule, name, None)
            if isinstance(v, torch.Tensor):
                out[f"R/{s}->{t}/{name}"] = v.detach().float().cpu().numpy().copy()
        if type(rule).__name__ == "MSTDP" and isinstance(getattr(rule, "p_plus", None), torch.Tensor):
            out[f"R/{s}->{t}/eligibility"] = rule.eligibility.detach().float().cpu().numpy().copy()
    return out


def _run(name, device, windows=None):
    from oracle.oracle import OracleBackend

    fx = helpers.Fixture(name)
    net, inputs, kw, T = fx.build("cpu" if device == "oracle" else device)
    helpers.add_spike_monitors(net, T, device="cpu" if device == "oracle" else device)
    x = {k: (v if device == "oracle" else v.cuda()) for k, v in inputs.items()}
    spans = windows or [T]

    def go():
        t0 = 0
        for span in spans:
            net.run(inputs={k: v[t0:t0 + span] for k, v in x.items()}, time=span, one_spike_seed=cases.ONE_SPIKE_SEED, **kw)
            t0 += span

    if device == "oracle":
        with OracleBackend() as ob:
            go()
            assert ob.err == 0
    else:
        go()
        net.check_errors()
    state = helpers.snapshot(net)
    state.update(_rule_state(net))
    return fx, state, (helpers.spike_counts(net, T) if windows is None else None)


@pytest.mark.parametrize("name", NEW)
def test_new_rows_bit_exact_vs_oracle_and_close_to_reference(name):
    fx, s_gpu, c_gpu = _run(name, "cuda")
    _, s_cpu, c_cpu = _run(name, "oracle")
    assert sum(int(v.sum()) for k, v in c_cpu.items() if k.endswith("/count")) > 0
    helpers.assert_bit_identical(s_gpu, s_cpu, f"{name} state")
    helpers.assert_bit_identical(c_gpu, c_cpu, f"{name} spike counts")
    helpers.assert_close_to_golden(fx, {k: v for k, v in s_gpu.items() if not k.startswith("R/")}, c_gpu)


@pytest.mark.parametrize("name", ["mstdp_dense", "conv_mstdp"])
def test_rule_state_carries_across_windows_of_odd_and_even_length(name):
    """The rule's double-buffered state must land in the caller's tensors whatever the window length."""
    spans = [7, 6, 1, 10]
    _, s_gpu, _ = _run(name, "cuda", windows=spans)
    _, s_cpu, _ = _run(name, "oracle", windows=spans)
    helpers.assert_bit_identical(s_gpu, s_cpu, f"{name} after windows {spans}")
