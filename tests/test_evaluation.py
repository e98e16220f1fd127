"""bindsnet_b200.evaluation against the live reference's evaluation functions (evaluation/evaluation.py:8-61,
99-180) on the CPU, from rasters and from the per-sample counts the window kernels deliver."""
import pytest
import torch

import cases

try:
    cases.namespace("reference")
    import bindsnet.evaluation.evaluation as ref_eval   # needs sklearn (present in the image)
except Exception:  # pragma: no cover
    ref_eval = None

pytestmark = pytest.mark.skipif(ref_eval is None, reason="live reference not available")


def test_label_assignment_and_readout_match_the_reference():
    from bindsnet_b200 import evaluation as ev

    g = torch.Generator().manual_seed(3)
    S, T, n, L = 60, 30, 50, 10
    spikes = (torch.rand(S, T, n, generator=g) < 0.1 * torch.rand(1, 1, n, generator=g) + 0.02).byte()
    labels = torch.randint(0, L - 1, (S,), generator=g)            # the last class never occurs (n_labeled == 0 branch)
    a_ref, p_ref, r_ref = ref_eval.assign_labels(spikes.float(), labels, L)
    a, p, r = ev.assign_labels(spikes, labels, L)
    assert torch.equal(a, a_ref) and torch.allclose(p, p_ref, atol=1e-6) and torch.allclose(r, r_ref, atol=1e-6)
    # from counts (what SpikeCounter hands over), second batch with the running rates and alpha
    spikes2 = (torch.rand(S, T, n, generator=g) < 0.08).byte()
    a_ref2, p_ref2, r_ref2 = ref_eval.assign_labels(spikes2.float(), labels, L, rates=r_ref.clone(), alpha=0.9)
    a2, p2, r2 = ev.assign_labels(spikes2.sum(1).int(), labels, L, rates=r.clone(), alpha=0.9)
    assert torch.equal(a2, a_ref2) and torch.allclose(p2, p_ref2, atol=1e-6) and torch.allclose(r2, r_ref2, atol=1e-6)
    assert torch.equal(ev.all_activity(spikes2.sum(1).int(), a2, L), ref_eval.all_activity(spikes2.float(), a_ref2, L))
    assert torch.equal(ev.proportion_weighting(spikes2, a2, p2, L), ref_eval.proportion_weighting(spikes2.float(), a_ref2, p_ref2, L))
