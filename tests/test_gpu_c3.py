"""BASELINE.json configs[2] (DiehlAndCook2015, n_neurons = 6400, batch 256) on the GPU (needs a B200).

The live reference cannot run this shape (42 GB of temporaries), so the oracle — pinned
against the live reference on the smaller configurations — is the checker: the CUDA path must agree with it bit
for bit (state, weights, spike counts) over a window long enough for Ae to spike, learn and inhibit."""
import pytest
import torch

import cases
import helpers

pytestmark = pytest.mark.gpu

N, B, T = 6400, 256, 70


def _build(device):
    from bindsnet_b200.models import DiehlAndCook2015

    torch.manual_seed(3)
    net = DiehlAndCook2015(n_inpt=784, n_neurons=N, batch_size=B, inpt_shape=(1, 28, 28), dt=1.0, nu=(1e-4, 1e-2), norm=78.4,
                           theta_plus=0.05, exc=22.5, inh=120.0)
    x = cases._poisson_inputs(cases.namespace("b200"), T, B, (1, 28, 28), 77)
    if device != "cpu":
        net.to(device)
    return net, {"X": x}


@pytest.mark.parametrize("tier", [0, 1])
def test_config3_bit_exact_vs_oracle(tier):
    from bindsnet_b200 import _backend
    from oracle.oracle import OracleBackend

    net, inputs = _build("cuda")
    net.force_tier = tier
    helpers.add_spike_monitors(net, T, device="cuda")
    net.run(inputs={k: v.cuda() for k, v in inputs.items()}, time=T, one_spike_seed=cases.ONE_SPIKE_SEED)
    net.check_errors()
    used = _backend.last_tier
    s_gpu, c_gpu = helpers.snapshot(net), helpers.spike_counts(net, T)
    del net
    torch.cuda.empty_cache()

    ref, inputs = _build("cpu")
    helpers.add_spike_monitors(ref, T)
    with OracleBackend() as ob:
        ref.run(inputs=inputs, time=T, one_spike_seed=cases.ONE_SPIKE_SEED)
        assert ob.err == 0
    s_cpu, c_cpu = helpers.snapshot(ref), helpers.spike_counts(ref, T)
    assert int(c_cpu["L/Ae/count"].sum()) > 0, "no Ae spikes: nothing tested"
    helpers.assert_bit_identical(s_gpu, s_cpu, f"config 3 state (tier {used})")
    helpers.assert_bit_identical(c_gpu, c_cpu, f"config 3 spike counts (tier {used})")
    if tier == 1:
        assert used == 1
