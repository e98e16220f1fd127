"""bindsnet_b200 — a Blackwell-native simulation core behind BindsNET's plug-in API.

Scope: the ``Network.run()`` hot path (LIF / DiehlAndCook neuron update, spike x weight
current injection, PostPre-family STDP) executed by hand-written sm_100a kernels in
``bindsnet_b200/csrc`` behind the C ABI of ``include/snn_b200.h``.  Import order mirrors the
reference's (network before learning/models).
"""
from . import _abi  # noqa: F401
from . import network  # noqa: F401  (must precede learning: circular import order of the reference)
from . import learning  # noqa: F401
from . import models  # noqa: F401
from . import encoding  # noqa: F401
from . import pipeline  # noqa: F401

__version__ = "0.1.0"
