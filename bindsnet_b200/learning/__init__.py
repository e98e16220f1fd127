"""Mirror of ``bindsnet.learning`` (reference: bindsnet/learning/__init__.py)."""
from .learning import (
    Hebbian,
    LearningRule,
    MSTDP,
    MSTDPET,
    NoOp,
    PostPre,
    Rmax,
    WeightDependentPostPre,
)

from .reward import AbstractReward, MovingAvgRPE

__all__ = [
    "AbstractReward", "MovingAvgRPE",
    "LearningRule", "NoOp", "PostPre", "WeightDependentPostPre", "Hebbian", "MSTDP", "MSTDPET", "Rmax",
]
