"""Canned networks — mirror of ``bindsnet/models/models.py`` (wiring only).

``TwoLayerNetwork`` (models.py:21-91), ``DiehlAndCook2015`` (:94-244), ``DiehlAndCook2015v2`` (:247-346),
``IncreasingInhibitionNetwork`` (:349-454) and ``LocallyConnectedNetwork`` (:457-584) build exactly the graphs the reference builds, out of this
package's ``Nodes`` / ``Connection`` / ``MulticompartmentConnection`` objects, with the same
constructor signatures and defaults.
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence, Union

import numpy as np
import torch
from torch.nn.modules.utils import _pair

from ..learning import PostPre
from ..learning.MCC_learning import PostPre as MMCPostPre
from ..network import Network
from ..network.nodes import DiehlAndCookNodes, Input, LIFNodes
from ..network.topology import Connection, LocalConnection, MulticompartmentConnection
from ..network.topology_features import Weight


class TwoLayerNetwork(Network):
    """``Input -> LIFNodes`` with a PostPre ``Connection`` (reference: models.py:21-91)."""

    def __init__(self, n_inpt: int, n_neurons: int = 100, dt: float = 1.0, wmin: float = 0.0, wmax: float = 1.0,
                 nu: Optional[Union[float, Sequence[float]]] = (1e-4, 1e-2), reduction: Optional[callable] = None,
                 norm: float = 78.4) -> None:
        super().__init__(dt=dt)
        self.n_inpt, self.n_neurons, self.dt = n_inpt, n_neurons, dt
        self.add_layer(Input(n=self.n_inpt, traces=True, tc_trace=20.0), name="X")
        self.add_layer(
            LIFNodes(n=self.n_neurons, traces=True, rest=-65.0, reset=-65.0, thresh=-52.0, refrac=5,
                     tc_decay=100.0, tc_trace=20.0),
            name="Y",
        )
        w = 0.3 * torch.rand(self.n_inpt, self.n_neurons)
        self.add_connection(
            Connection(source=self.layers["X"], target=self.layers["Y"], w=w, update_rule=PostPre, nu=nu,
                       reduction=reduction, wmin=wmin, wmax=wmax, norm=norm),
            source="X", target="Y",
        )


class DiehlAndCook2015(Network):
    """Diehl & Cook (2015): ``X -> Ae <-> Ai`` with MCC connections (reference:
    models.py:94-244).  This is the graph the fused CUDA window kernel recognises."""

    def __init__(self, n_inpt: int, device: str = "cpu", batch_size: int = None, sparse: bool = False,
                 n_neurons: int = 100, exc: float = 22.5, inh: float = 17.5, dt: float = 1.0,
                 nu: Optional[Union[float, Sequence[float]]] = (1e-4, 1e-2), reduction: Optional[callable] = None,
                 wmin: float = 0.0, wmax: float = 1.0, w_dtype: torch.dtype = torch.float32, norm: float = 78.4,
                 theta_plus: float = 0.05, tc_theta_decay: float = 1e7, inpt_shape: Optional[Iterable[int]] = None,
                 inh_thresh: float = -40.0, exc_thresh: float = -52.0) -> None:
        super().__init__(dt=dt)
        self.n_inpt, self.inpt_shape, self.n_neurons, self.exc, self.inh = n_inpt, inpt_shape, n_neurons, exc, inh
        self.dt = dt

        input_layer = Input(n=self.n_inpt, shape=self.inpt_shape, traces=True, tc_trace=20.0)
        exc_layer = DiehlAndCookNodes(
            n=self.n_neurons, traces=True, rest=-65.0, reset=-60.0, thresh=exc_thresh, refrac=5,
            tc_decay=100.0, tc_trace=20.0, theta_plus=theta_plus, tc_theta_decay=tc_theta_decay,
        )
        inh_layer = LIFNodes(
            n=self.n_neurons, traces=False, rest=-60.0, reset=-45.0, thresh=inh_thresh, tc_decay=10.0,
            refrac=2, tc_trace=20.0,
        )

        w = 0.3 * torch.rand(self.n_inpt, self.n_neurons)
        input_exc_conn = MulticompartmentConnection(
            source=input_layer, target=exc_layer, device=device,
            pipeline=[
                Weight("weight", w, value_dtype=w_dtype, range=[wmin, wmax], norm=norm, reduction=reduction,
                       nu=nu, learning_rule=MMCPostPre, sparse=sparse, batch_size=batch_size)
            ],
        )
        w = self.exc * torch.diag(torch.ones(self.n_neurons))
        exc_inh_conn = MulticompartmentConnection(
            source=exc_layer, target=inh_layer, device=device,
            pipeline=[Weight("weight", w, value_dtype=w_dtype, range=[0, self.exc], sparse=sparse)],
        )
        w = -self.inh * (torch.ones(self.n_neurons, self.n_neurons) - torch.diag(torch.ones(self.n_neurons)))
        inh_exc_conn = MulticompartmentConnection(
            source=inh_layer, target=exc_layer, device=device,
            pipeline=[Weight("weight", w, value_dtype=w_dtype, range=[-self.inh, 0], sparse=sparse)],
        )

        self.add_layer(input_layer, name="X")
        self.add_layer(exc_layer, name="Ae")
        self.add_layer(inh_layer, name="Ai")
        self.add_connection(input_exc_conn, source="X", target="Ae")
        self.add_connection(exc_inh_conn, source="Ae", target="Ai")
        self.add_connection(inh_exc_conn, source="Ai", target="Ae")
        if str(device) != "cpu":
            self.to(device)


class DiehlAndCook2015v2(Network):
    """Variant with recurrent lateral inhibition instead of an inhibitory layer, built from
    classic ``Connection`` + ``learning.PostPre`` (reference: models.py:247-346)."""

    def __init__(self, n_inpt: int, n_neurons: int = 100, inh: float = 17.5, dt: float = 1.0,
                 nu: Optional[Union[float, Sequence[float]]] = (1e-4, 1e-2), reduction: Optional[callable] = None,
                 wmin: Optional[float] = 0.0, wmax: Optional[float] = 1.0, norm: float = 78.4,
                 theta_plus: float = 0.05, tc_theta_decay: float = 1e7, inpt_shape: Optional[Iterable[int]] = None,
                 exc_thresh: float = -52.0) -> None:
        super().__init__(dt=dt)
        self.n_inpt, self.inpt_shape, self.n_neurons, self.inh, self.dt = n_inpt, inpt_shape, n_neurons, inh, dt

        self.add_layer(Input(n=self.n_inpt, shape=self.inpt_shape, traces=True, tc_trace=20.0), name="X")
        self.add_layer(
            DiehlAndCookNodes(
                n=self.n_neurons, traces=True, rest=-65.0, reset=-60.0, thresh=exc_thresh, refrac=5,
                tc_decay=100.0, tc_trace=20.0, theta_plus=theta_plus, tc_theta_decay=tc_theta_decay,
            ),
            name="Y",
        )
        w = 0.3 * torch.rand(self.n_inpt, self.n_neurons)
        self.add_connection(
            Connection(source=self.layers["X"], target=self.layers["Y"], w=w, update_rule=PostPre, nu=nu,
                       reduction=reduction, wmin=wmin, wmax=wmax, norm=norm),
            source="X", target="Y",
        )
        w = -self.inh * (torch.ones(self.n_neurons, self.n_neurons) - torch.diag(torch.ones(self.n_neurons)))
        self.add_connection(
            Connection(source=self.layers["Y"], target=self.layers["Y"], w=w, wmin=-self.inh, wmax=0),
            source="Y", target="Y",
        )


class IncreasingInhibitionNetwork(Network):
    """``Input -> DiehlAndCookNodes`` with a PostPre ``Connection`` and a recurrent ``Connection`` whose strength grows
    with the distance between two neurons on the ``sqrt(n) x sqrt(n)`` grid (reference: models.py:349-454).
    The recurrent matrix is the reference's, quirks included: the square root of the Euclidean grid distance,
    divided by its maximum, times ``max_inhib`` plus ``start_inhib`` — on the diagonal too, and with a positive
    sign (:439-449)."""

    def __init__(self, n_input: int, n_neurons: int = 100, start_inhib: float = 1.0, max_inhib: float = 100.0,
                 dt: float = 1.0, nu: Optional[Union[float, Sequence[float]]] = (1e-4, 1e-2),
                 reduction: Optional[callable] = None, wmin: float = 0.0, wmax: float = 1.0, norm: float = 78.4,
                 theta_plus: float = 0.05, tc_theta_decay: float = 1e7, inpt_shape: Optional[Iterable[int]] = None,
                 exc_thresh: float = -52.0) -> None:
        super().__init__(dt=dt)
        self.n_input, self.n_neurons = n_input, n_neurons
        self.n_sqrt = int(np.sqrt(n_neurons))
        self.start_inhib, self.max_inhib, self.dt, self.inpt_shape = start_inhib, max_inhib, dt, inpt_shape

        self.add_layer(Input(n=self.n_input, shape=self.inpt_shape, traces=True, tc_trace=20.0), name="X")
        self.add_layer(
            DiehlAndCookNodes(
                n=self.n_neurons, traces=True, rest=-65.0, reset=-60.0, thresh=exc_thresh, refrac=5,
                tc_decay=100.0, tc_trace=20.0, theta_plus=theta_plus, tc_theta_decay=tc_theta_decay,
            ),
            name="Y",
        )
        w = 0.3 * torch.rand(self.n_input, self.n_neurons)
        self.add_connection(
            Connection(source=self.layers["X"], target=self.layers["Y"], w=w, update_rule=PostPre, nu=nu,
                       reduction=reduction, wmin=wmin, wmax=wmax, norm=norm),
            source="X", target="Y",
        )
        # models.py:436-449, all pairs at once: float64 distances rounded to float32 (what the per-element assignment
        # of the reference does), then the reference's float32 tensor arithmetic
        idx = np.arange(self.n_neurons)
        gx, gy = idx // self.n_sqrt, idx % self.n_sqrt
        dist = np.sqrt(((gx[:, None] - gx[None, :]) ** 2 + (gy[:, None] - gy[None, :]) ** 2).astype(np.float64))
        w = torch.from_numpy(np.sqrt(dist).astype(np.float32))
        w = w / w.max()
        w = (w * self.max_inhib) + self.start_inhib
        self.add_connection(Connection(source=self.layers["Y"], target=self.layers["Y"], w=w), source="Y", target="Y")


class LocallyConnectedNetwork(Network):
    """``Input -> DiehlAndCookNodes`` through a ``LocalConnection`` with PostPre, the output neurons that share a
    receptive field inhibiting each other through a recurrent ``Connection`` (reference: models.py:457-584)."""

    def __init__(self, n_inpt: int, input_shape, kernel_size, stride, n_filters: int, inh: float = 25.0,
                 dt: float = 1.0, nu: Optional[Union[float, Sequence[float]]] = (1e-4, 1e-2),
                 reduction: Optional[callable] = None, theta_plus: float = 0.05, tc_theta_decay: float = 1e7,
                 wmin: float = 0.0, wmax: float = 1.0, norm: Optional[float] = 0.2,
                 exc_thresh: float = -52.0) -> None:
        super().__init__(dt=dt)
        kernel_size, stride = _pair(kernel_size), _pair(stride)
        self.n_inpt, self.input_shape, self.kernel_size, self.stride = n_inpt, input_shape, kernel_size, stride
        self.n_filters, self.inh, self.dt, self.theta_plus = n_filters, inh, dt, theta_plus
        self.tc_theta_decay, self.wmin, self.wmax, self.norm = tc_theta_decay, wmin, wmax, norm

        if kernel_size == input_shape:                                           # models.py:530-536
            conv_size = (1, 1)
        else:
            conv_size = (int((input_shape[0] - kernel_size[0]) / stride[0]) + 1,
                         int((input_shape[1] - kernel_size[1]) / stride[1]) + 1)
        fields = conv_size[0] * conv_size[1]
        n_out = self.n_filters * fields

        X = Input(n=self.n_inpt, traces=True, tc_trace=20.0)
        Y = DiehlAndCookNodes(
            n=n_out, traces=True, rest=-65.0, reset=-60.0, thresh=exc_thresh, refrac=5, tc_decay=100.0,
            tc_trace=20.0, theta_plus=theta_plus, tc_theta_decay=tc_theta_decay,
        )
        local = LocalConnection(
            X, Y, kernel_size=kernel_size, stride=stride, n_filters=n_filters, nu=nu, reduction=reduction,
            update_rule=PostPre, wmin=wmin, wmax=wmax, norm=norm, input_shape=input_shape,
        )
        # models.py:567-579: neuron f * fields + c inhibits every other filter's neuron at the same field c
        f = torch.arange(n_out) // fields
        c = torch.arange(n_out) % fields
        same_field = (c.view(-1, 1) == c.view(1, -1)) & (f.view(-1, 1) != f.view(1, -1))
        w = torch.zeros(n_out, n_out)
        w[same_field] = -inh
        self.add_layer(X, name="X")
        self.add_layer(Y, name="Y")
        self.add_connection(local, source="X", target="Y")
        self.add_connection(Connection(Y, Y, w=w), source="Y", target="Y")
