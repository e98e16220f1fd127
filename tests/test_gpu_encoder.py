"""On-device spike encoders (csrc/snn_encode.cu) against the reference's encoder semantics
(bindsnet/encoding/encodings.py:50-96, 99-156; restated on the CPU in bindsnet_b200.encoding).  The reference draws from
torch's global generator, so the comparison is in distribution: firing rates, inter-spike-interval statistics, first
spike times.  Needs a B200."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RATES = [0.0, 2.0, 10.0, 31.0, 64.0, 128.0]   # Hz; 128 = --intensity of examples/mnist/batch_eth_mnist.py
PER_RATE, T = 6000, 250


def _isis(spikes):   # spikes: [T, n] 0/1 numpy -> concatenated inter-spike intervals, first spike times (1-based)
    isis, first = [], []
    for col in spikes.T:
        t = np.flatnonzero(col)
        if t.size:
            first.append(t[0] + 1)
            isis.append(np.diff(t))
    return (np.concatenate(isis) if isis else np.zeros(0)), np.asarray(first)


def test_poisson_encoder_matches_reference_distribution():
    from bindsnet_b200.encoding import poisson

    rate = torch.tensor(RATES).repeat_interleave(PER_RATE)
    torch.manual_seed(5)
    ref = poisson(rate, time=T, dt=1.0).numpy()                      # CPU restatement of encodings.py:99-156
    dev = poisson(rate.cuda(), time=T, dt=1.0, seed=77).cpu().numpy()
    assert dev.shape == ref.shape == (T, rate.numel()) and dev.dtype == np.uint8 and set(np.unique(dev)) <= {0, 1}
    for k, r in enumerate(RATES):
        a, b = ref[:, k * PER_RATE:(k + 1) * PER_RATE], dev[:, k * PER_RATE:(k + 1) * PER_RATE]
        if r == 0.0:
            assert b.sum() == 0 and a.sum() == 0
            continue
        # spike counts per train: same mean within the sampling error of the two estimates (5 sigma)
        ca, cb = a.sum(0).astype(np.float64), b.sum(0).astype(np.float64)
        se = np.sqrt(ca.var() / PER_RATE + cb.var() / PER_RATE) + 1e-9
        assert abs(ca.mean() - cb.mean()) < 5 * se + 1e-3, f"{r} Hz: mean count {ca.mean():.4f} vs {cb.mean():.4f}"
        ia, fa = _isis(a)
        ib, fb = _isis(b)
        if ia.size > 500 and ib.size > 500:
            se = np.sqrt(ia.var() / ia.size + ib.var() / ib.size)
            assert abs(ia.mean() - ib.mean()) < 5 * se, f"{r} Hz: mean ISI {ia.mean():.3f} vs {ib.mean():.3f}"
            assert abs(ia.std() - ib.std()) < 0.08 * ia.std() + 0.05, f"{r} Hz: ISI spread {ia.std():.3f} vs {ib.std():.3f}"
            assert ib.min() >= 1   # zero intervals are bumped to one step (encodings.py:145)
            # two-sample Kolmogorov-Smirnov distance between the ISI distributions
            grid = np.arange(1, int(max(ia.max(), ib.max())) + 1)
            Fa = np.searchsorted(np.sort(ia), grid, side="right") / ia.size
            Fb = np.searchsorted(np.sort(ib), grid, side="right") / ib.size
            assert np.abs(Fa - Fb).max() < 1.95 * np.sqrt((ia.size + ib.size) / (ia.size * ib.size)) + 0.01
        if fa.size > 500 and fb.size > 500:
            se = np.sqrt(fa.var() / fa.size + fb.var() / fb.size)
            assert abs(fa.mean() - fb.mean()) < 5 * se + 1e-3, f"{r} Hz: first spike {fa.mean():.3f} vs {fb.mean():.3f}"


def test_poisson_encoder_is_a_function_of_seed_and_element():
    from bindsnet_b200.encoding import poisson

    rate = (128.0 * torch.rand(3, 1, 28, 28)).cuda()
    a = poisson(rate, time=100, seed=1)
    b = poisson(rate, time=100, seed=1)
    c = poisson(rate, time=100, seed=2)
    assert a.shape == (100, 3, 1, 28, 28) and torch.equal(a, b) and not torch.equal(a, c)
    # an element's train does not depend on what else is in the batch (counter-based stream per element)
    d = poisson(rate[:1], time=100, seed=1)
    assert torch.equal(d[:, 0], a[:, 0])


def test_bernoulli_encoder_rates_and_window_consumption():
    from bindsnet_b200.encoding import bernoulli
    from bindsnet_b200.models import DiehlAndCook2015

    p = torch.tensor([0.0, 0.05, 0.3, 1.0]).repeat_interleave(5000).cuda()
    s = bernoulli(p, time=200, seed=3).float()
    got = s.view(200, 4, 5000).mean(dim=(0, 2)).cpu().numpy()
    assert got[0] == 0.0 and got[3] == 1.0
    assert abs(got[1] - 0.05) < 5 * np.sqrt(0.05 * 0.95 / 1e6) and abs(got[2] - 0.3) < 5 * np.sqrt(0.3 * 0.7 / 1e6)
    s2 = bernoulli(4.0 * p, time=50, max_prob=0.5, seed=3).float()   # datum.max() > 1: normalised, then scaled (encodings.py:80-84)
    got2 = s2.view(50, 4, 5000).mean(dim=(0, 2)).cpu().numpy()
    assert abs(got2[3] - 0.5) < 0.01 and abs(got2[2] - 0.15) < 0.01
    # the encoded tensor feeds Network.run directly, no host round trip
    from bindsnet_b200.encoding import poisson
    net = DiehlAndCook2015(n_inpt=784, n_neurons=64, batch_size=4, inpt_shape=(1, 28, 28), inh=120.0).to("cuda")
    x = poisson(128.0 * torch.rand(4, 1, 28, 28, device="cuda"), time=60, seed=9)
    net.run({"X": x}, time=60)
    net.check_errors()
    assert int(net.layers["X"].s.sum()) == int(x[-1].sum())
