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

__all__ = [
    "LearningRule", "NoOp", "PostPre", "WeightDependentPostPre", "Hebbian", "MSTDP", "MSTDPET", "Rmax",
]
