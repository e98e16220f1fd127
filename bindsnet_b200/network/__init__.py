"""Mirror of ``bindsnet.network`` (reference: bindsnet/network/__init__.py)."""
from . import monitors, nodes, topology, topology_features
from .network import Network, load

__all__ = ["Network", "load", "nodes", "topology", "topology_features", "monitors"]
