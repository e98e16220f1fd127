"""User-defined models drop in unchanged (north_star; extension contract docs/source/guide/guide_part_ii.rst:69-73,
learning.py:31-104): a ``Nodes`` subclass with its own torch ``forward`` and a ``LearningRule`` subclass with its own
``update`` run through ``Network.run`` — step by step, the built-in pieces on their single-operator kernels (here: the
oracle backend on the CPU) — and reproduce the built-in LIFNodes + PostPre network they restate."""
import numpy as np
import torch

from bindsnet_b200.learning import LearningRule, PostPre
from bindsnet_b200.network import Network
from bindsnet_b200.network.monitors import Monitor
from bindsnet_b200.network.nodes import Input, LIFNodes, Nodes
from bindsnet_b200.network.topology import Connection
from oracle.oracle import OracleBackend


class MyLIF(Nodes):
    """A user's leaky integrate-and-fire population, written like the reference's own (nodes.py:500-529)."""

    def __init__(self, n, thresh=-52.0, rest=-65.0, reset=-65.0, refrac=5, tc_decay=100.0, **kw):
        super().__init__(n=n, **kw)
        self.register_buffer("rest", torch.tensor(rest)); self.register_buffer("reset", torch.tensor(reset))
        self.register_buffer("thresh", torch.tensor(thresh)); self.register_buffer("refrac", torch.tensor(float(refrac)))
        self.register_buffer("tc_decay", torch.tensor(tc_decay)); self.register_buffer("decay", torch.zeros(()))
        self.register_buffer("v", torch.zeros(0)); self.register_buffer("refrac_count", torch.zeros(0))

    def forward(self, x):
        self.v = self.decay * (self.v - self.rest) + self.rest
        x = x.clone(); x.masked_fill_(self.refrac_count > 0, 0.0)
        self.refrac_count -= self.dt
        self.v += x
        self.s = self.v >= self.thresh
        self.refrac_count.masked_fill_(self.s, float(self.refrac))
        self.v.masked_fill_(self.s, float(self.reset))
        super().forward(x)

    def compute_decays(self, dt):
        super().compute_decays(dt)
        self.decay = torch.exp(-self.dt / self.tc_decay)

    def set_batch_size(self, batch_size):
        super().set_batch_size(batch_size)
        self.v = self.rest * torch.ones(batch_size, *self.shape)
        self.refrac_count = torch.zeros(batch_size, *self.shape)

    def reset_state_variables(self):
        super().reset_state_variables()
        self.v.fill_(float(self.rest)); self.refrac_count.zero_()


class MyPostPre(LearningRule):
    """A user's pair-based STDP, written like learning.py:390-420."""

    def update(self, **kwargs):
        B = self.source.batch_size
        s_pre = self.source.s.view(B, -1).float(); x_pre = self.source.x.view(B, -1)
        s_post = self.target.s.view(B, -1).float(); x_post = self.target.x.view(B, -1)
        w = self.connection.w
        w -= self.nu[0] * torch.einsum("bi,bj->ij", s_pre, x_post)
        w += self.nu[1] * torch.einsum("bi,bj->ij", x_pre, s_post)
        super().update()


def _net(custom: bool, w0, custom_rule=None):
    custom_rule = custom if custom_rule is None else custom_rule
    net = Network(dt=1.0, batch_size=3)
    X = Input(n=50, traces=True)
    Y = (MyLIF if custom else LIFNodes)(n=30, traces=True, thresh=-57.0, refrac=2, tc_decay=60.0)
    C = Connection(X, Y, w=w0.clone(), update_rule=MyPostPre if custom_rule else PostPre, nu=(2e-3, 2e-2), reduction=torch.sum,
                   wmin=0.0, wmax=1.0, norm=12.0)
    net.add_layer(X, "X"); net.add_layer(Y, "Y"); net.add_connection(C, "X", "Y")
    net.add_monitor(Monitor(Y, ["s", "v"], time=60), "Y")
    return net


def test_user_defined_nodes_and_rule_run_through_the_scripted_tier():
    g = torch.Generator().manual_seed(2)
    w0 = 0.9 * torch.rand(50, 30, generator=g)
    x = torch.bernoulli(0.15 * torch.ones(60, 3, 50), generator=g).byte()
    a, b = _net(True, w0), _net(False, w0)
    assert a._scripted_required() and not b._scripted_required()
    with OracleBackend():
        a.run({"X": x}, time=60)
        b.run({"X": x}, time=60)
    sa, sb = a.monitors["Y"].get("s"), b.monitors["Y"].get("s")
    assert sa.shape == sb.shape == (60, 3, 30) and int(sb.sum()) > 20
    assert torch.equal(sa, sb), "spike rasters of the user-defined and the built-in network differ"
    assert torch.allclose(a.monitors["Y"].get("v"), b.monitors["Y"].get("v"), atol=1e-4)
    wa, wb = a.connections[("X", "Y")].w, b.connections[("X", "Y")].w
    assert float((wa - wb).abs().max() / wb.abs().max()) < 1e-5
    assert torch.allclose(a.layers["Y"].x, b.layers["Y"].x, atol=1e-6)


def test_user_defined_nodes_under_a_built_in_rule():
    """A user's population as the target of a built-in connection with a built-in rule: the rule's single-operator
    update reads the population's spikes and traces (bindsnet_b200/network/_plan.py:_fill_endpoint); the result is the
    built-in network's, bit for bit (same kernels on both sides of the update)."""
    g = torch.Generator().manual_seed(4)
    w0 = 0.9 * torch.rand(50, 30, generator=g)
    x = torch.bernoulli(0.15 * torch.ones(60, 3, 50), generator=g).byte()
    a, b = _net(True, w0, custom_rule=False), _net(False, w0)
    assert a._scripted_required()
    with OracleBackend():
        a.run({"X": x}, time=60)
        b.run({"X": x}, time=60)
    assert torch.equal(a.monitors["Y"].get("s"), b.monitors["Y"].get("s")) and int(b.monitors["Y"].get("s").sum()) > 20
    wa, wb = a.connections[("X", "Y")].w, b.connections[("X", "Y")].w
    assert float((wa - wb).abs().max() / wb.abs().max()) < 1e-5
