"""Randomised kernel-vs-oracle parity (needs a B200): the networks of test_oracle_fuzz_vs_reference.py — drawn from a
seeded generator over node kinds, learning rules, reductions, options and batch sizes — through the generic window kernel,
bit for bit against the oracle (state, weights, spike counts).  The CPU twin of this file pins the oracle on the same
networks against the live reference.

Runs last (file name) on purpose: it was written after the round's GPU time was spent, so its first execution is the
driver's; everything before it is the suite that ran on the B200 during the round."""
import pytest

import cases
import helpers
import test_oracle_fuzz_vs_reference as fuzz

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", list(range(24)))
def test_random_network_generic_kernel_bit_exact_vs_oracle(seed):
    from bindsnet_b200 import _backend
    from oracle.oracle import OracleBackend

    spec = fuzz._draw(seed)
    ns = cases.namespace("b200")
    T = spec["T"]

    gpu, x = fuzz._build(ns, spec)
    gpu.to("cuda")
    gpu.force_tier = 1
    helpers.add_spike_monitors(gpu, T, device="cuda")
    gpu.run(inputs={"X": x.cuda()}, time=T)
    gpu.check_errors()
    assert _backend.last_tier == 1
    s_gpu, c_gpu = helpers.snapshot(gpu), helpers.spike_counts(gpu, T)

    cpu, x2 = fuzz._build(ns, spec)
    helpers.add_spike_monitors(cpu, T)
    with OracleBackend() as ob:
        cpu.run(inputs={"X": x2}, time=T)
        assert ob.err == 0
    s_cpu, c_cpu = helpers.snapshot(cpu), helpers.spike_counts(cpu, T)
    what = f"seed {seed} {spec['kind']} {spec['rule']} B={spec['B']}"
    helpers.assert_bit_identical(s_gpu, s_cpu, what + " state")
    helpers.assert_bit_identical(c_gpu, c_cpu, what + " spike counts")
