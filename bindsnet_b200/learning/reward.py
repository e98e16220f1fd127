"""Reward shaping for reward-modulated rules — mirror of ``bindsnet.learning.reward``
(reference: bindsnet/learning/reward.py:6-88).  Host logic only: ``Network.run`` asks the network's
``reward_fn`` once per window for the reward the MSTDP / MSTDPET kernels are launched with
(network.py:325-326); the user calls ``update`` once per episode."""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch


class AbstractReward(ABC):
    """Interface of a reward modifier (reward.py:6-26): ``compute(**run_kwargs)`` returns the reward
    of the coming window, ``update(**kwargs)`` advances whatever the modifier keeps between episodes."""

    @abstractmethod
    def compute(self, **kwargs):
        ...

    @abstractmethod
    def update(self, **kwargs) -> None:
        ...


class MovingAvgRPE(AbstractReward):
    """Reward prediction error against an exponential moving average of earlier rewards
    (reward.py:29-88).  All arithmetic is fp32 tensor arithmetic, as in the reference, so the
    scalar handed to the kernels is the same."""

    def __init__(self, **kwargs) -> None:
        self.reward_predict = torch.tensor(0.0)          # per-step prediction (reward.py:40)
        self.reward_predict_episode = torch.tensor(0.0)  # per-episode prediction (:41)
        self.rewards_predict_episode = []                # history of the latter (:42-44)

    def compute(self, **kwargs) -> torch.Tensor:
        # reward.py:57-59: the error is taken against the per-step prediction
        return kwargs["reward"] - self.reward_predict

    def update(self, **kwargs) -> None:
        # reward.py:73-88
        total = kwargs["accumulated_reward"]
        steps = torch.tensor(kwargs["steps"]).float()
        window = torch.tensor(kwargs.get("ema_window", 10.0))
        keep, take = 1 - 1 / window, 1 / window
        self.reward_predict = keep * self.reward_predict + take * (total / steps)
        self.reward_predict_episode = keep * self.reward_predict_episode + take * total
        self.rewards_predict_episode.append(self.reward_predict_episode.item())
