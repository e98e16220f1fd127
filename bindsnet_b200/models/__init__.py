"""Mirror of ``bindsnet.models`` (reference: bindsnet/models/__init__.py)."""
from .models import (
    DiehlAndCook2015,
    DiehlAndCook2015v2,
    IncreasingInhibitionNetwork,
    LocallyConnectedNetwork,
    TwoLayerNetwork,
)

__all__ = [
    "TwoLayerNetwork", "DiehlAndCook2015", "DiehlAndCook2015v2", "IncreasingInhibitionNetwork",
    "LocallyConnectedNetwork",
]
