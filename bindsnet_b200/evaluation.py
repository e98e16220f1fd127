"""Label assignment and classification — the step after the hot path (reference: bindsnet/evaluation/evaluation.py
``assign_labels`` :8-61, ``all_activity`` :99-136, ``proportion_weighting`` :139-180; consumer loop
examples/mnist/batch_eth_mnist.py:212-264,280-284).

Same signatures and results as the reference.  ``spikes`` may be the reference's ``[n_samples, time, n_neurons]``
raster or — the point of SURVEY.md §8f rank 2 — the ``[n_samples, n_neurons]`` per-sample spike COUNTS that the window
kernels accumulate themselves (``SpikeCounter`` / ``snn_layer_t.rec_count``), so that no raster is ever written.
CUDA tensors run on the kernels of ``csrc/snn_readout.cu``; CPU tensors on the torch formulas below (same arithmetic:
counts are integers, so every sum is exact)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def _counts(spikes: torch.Tensor) -> torch.Tensor:
    """[n_samples, n_neurons] int32 spike counts from a raster (summed over time, evaluation.py:41) or from counts."""
    if spikes.dim() == 3:
        spikes = spikes.sum(1)
    elif spikes.dim() != 2:
        raise ValueError("spikes must be [n_samples, time, n_neurons] or [n_samples, n_neurons]")
    return spikes.to(torch.int32).contiguous()


def assign_labels(spikes: torch.Tensor, labels: torch.Tensor, n_labels: int, rates: Optional[torch.Tensor] = None,
                  alpha: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Class assignments ``[n_neurons]``, per-class spike proportions and per-class firing rates ``[n_neurons,
    n_labels]`` (evaluation.py:8-61); ``rates`` is updated in place like the reference does."""
    counts = _counts(spikes)
    n = counts.shape[1]
    if rates is None:
        rates = torch.zeros((n, n_labels), device=counts.device)
    labels = labels.to(counts.device, torch.int64).contiguous()
    if counts.is_cuda:
        from . import _backend

        if rates.dtype != torch.float32 or not rates.is_contiguous():
            raise TypeError("rates must be a contiguous float32 tensor (it is updated in place)")
        proportions = torch.empty_like(rates)
        assignments = torch.empty(n, dtype=torch.int64, device=counts.device)
        _backend.assign_labels(counts, labels, n_labels, alpha, rates, proportions, assignments)
        return assignments, proportions, rates
    c = counts.float()
    for i in range(n_labels):
        sel = labels == i
        n_labeled = sel.sum().float()
        if n_labeled > 0:
            rates[:, i] = alpha * rates[:, i] + c[sel].sum(0) / n_labeled
    proportions = rates / rates.sum(1, keepdim=True)
    proportions[proportions != proportions] = 0
    return torch.max(proportions, 1)[1], proportions, rates


def _predict(spikes, assignments, proportions, n_labels):
    counts = _counts(spikes)
    assignments = assignments.to(counts.device, torch.int64).contiguous()
    if counts.is_cuda:
        from . import _backend

        out = torch.empty(counts.shape[0], dtype=torch.int64, device=counts.device)
        p = None if proportions is None else proportions.to(counts.device, torch.float32).contiguous()
        _backend.predict(counts, assignments, p, n_labels, out)
        return out
    c = counts.float()
    rates = torch.zeros((counts.shape[0], n_labels))
    for i in range(n_labels):
        sel = assignments == i
        n_assigns = sel.sum().float()
        if n_assigns > 0:
            w = c if proportions is None else proportions[:, i] * c
            rates[:, i] = w[:, sel].sum(1) / n_assigns
    return torch.max(rates, 1)[1]


def all_activity(spikes: torch.Tensor, assignments: torch.Tensor, n_labels: int) -> torch.Tensor:
    """Label with the highest mean activity of its assigned neurons (evaluation.py:99-136)."""
    return _predict(spikes, assignments, None, n_labels)


def proportion_weighting(spikes: torch.Tensor, assignments: torch.Tensor, proportions: torch.Tensor, n_labels: int) -> torch.Tensor:
    """The same, every neuron weighted by its class proportion (evaluation.py:139-180)."""
    return _predict(spikes, assignments, proportions, n_labels)


def _steps_with_spikes(activity: torch.Tensor):
    """Per step that has spikes, the firing neurons in ascending order — one ``nonzero`` (and one device → host copy)
    per example instead of one per step."""
    T = activity.shape[0]
    idx = torch.nonzero(activity.reshape(T, -1)).cpu()           # row-major: by step, then by neuron
    steps = {}
    for t, j in idx.tolist():
        steps.setdefault(t, []).append(j)
    return [steps[t] for t in sorted(steps)]


def ngram(spikes: torch.Tensor, ngram_scores, n_labels: int, n: int) -> torch.Tensor:
    """Class per example from the scores of the length-``n`` runs in its firing order (evaluation.py:183-217):
    neurons are read step by step, ascending within a step; like the reference, the last ``n``-gram of an example is
    not scored (its loop stops at ``len - n``)."""
    out = []
    for activity in spikes:
        order = [j for step in _steps_with_spikes(activity) for j in step]
        score = torch.zeros(n_labels, device=spikes.device)
        for k in range(len(order) - n):
            hit = ngram_scores.get(tuple(order[k:k + n]))
            if hit is not None:
                score += hit
        out.append(int(torch.argmax(score)))
    return torch.tensor(out, device=spikes.device).long()


def update_ngram_scores(spikes: torch.Tensor, labels: torch.Tensor, n_labels: int, n: int, ngram_scores):
    """Adds, for every ``n`` consecutive steps-with-spikes of every example, one count of the example's label to each
    sequence that takes one firing neuron from each of those steps (evaluation.py:220-258).  ``ngram_scores`` is
    updated in place and returned."""
    from itertools import product

    for i, activity in enumerate(spikes):
        steps = _steps_with_spikes(activity)
        label = int(labels[i])
        for start in range(len(steps) - n + 1):
            for seq in product(*steps[start:start + n]):
                if seq not in ngram_scores:
                    ngram_scores[seq] = torch.zeros(n_labels, device=spikes.device)
                ngram_scores[seq][label] += 1
    return ngram_scores


def logreg_fit(spikes: torch.Tensor, labels: torch.Tensor, logreg):
    """(Re)fits a scikit-learn ``LogisticRegression`` on time-summed spikes (evaluation.py:64-79); device tensors are
    brought to the host, where scikit-learn works."""
    logreg.fit(spikes.detach().cpu(), labels.detach().cpu())
    return logreg


def logreg_predict(spikes: torch.Tensor, logreg) -> torch.Tensor:
    """Classes from a fitted model, ``-1`` for every example while it is unfitted (evaluation.py:82-96)."""
    if getattr(logreg, "coef_", None) is None:
        return -1 * torch.ones(spikes.size(0)).long()
    return torch.Tensor(logreg.predict(spikes.detach().cpu())).long()
