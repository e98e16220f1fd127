"""Where does the end-to-end window time go?  (diagnostic, not a benchmark)
A: resident inputs; B: + SpikeCounter; C: + AsyncReadback; D: + WindowPrefetcher (H2D overlapped)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bindsnet_b200 import _backend
from bindsnet_b200.network.monitors import SpikeCounter
from bindsnet_b200.pipeline import AsyncReadback, WindowPrefetcher

dev = torch.device("cuda", 0)
net = bench.make_network(dev)
host = [w.pin_memory() for w in bench.synth_windows(bench.POOL, seed=1234)]
resident = [w.to(dev) for w in host]
K = 20
rb = AsyncReadback(depth=2)


def timed(name, body):
    for i in range(3):
        body(i, False)
    while len(rb): rb.pop()
    torch.cuda.synchronize()
    _backend.kernel_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    body(K, True)
    while len(rb): rb.pop()
    e1.record()
    host_s = time.perf_counter() - t0
    torch.cuda.synchronize()
    kern = [a.elapsed_time(b) for a, b in _backend.kernel_events]
    _backend.kernel_events = None
    print(f"{name}: {e0.elapsed_time(e1) / K:.3f} ms/window (device), host enqueue {host_s / K * 1e3:.3f} ms/window, "
          f"window kernels {sum(kern) / max(len(kern), 1):.3f} ms")


def A(n, _):
    if not isinstance(n, int): return
    for i in range(n if _ else 1):
        net.reset_state_variables(); net.run({"X": resident[i % bench.POOL]}, time=bench.T_STEPS)

timed("A resident", A)
net.add_monitor(SpikeCounter(net.layers["Ae"]), "Ae_spikes")
timed("B + SpikeCounter", A)

def C(n, _):
    for i in range(n if _ else 1):
        net.reset_state_variables(); net.run({"X": resident[i % bench.POOL]}, time=bench.T_STEPS)
        rb.push(net.monitors["Ae_spikes"].get("s"))
        if len(rb) == rb.depth: rb.pop().numpy().sum()

timed("C + AsyncReadback", C)

def D(n, _):
    cnt = n if _ else 1
    pre = WindowPrefetcher(dev, (host[i % bench.POOL] for i in range(cnt)))
    for x in pre:
        net.reset_state_variables(); net.run({"X": x}, time=bench.T_STEPS); pre.release()
        rb.push(net.monitors["Ae_spikes"].get("s"))
        if len(rb) == rb.depth: rb.pop().numpy().sum()

timed("D + WindowPrefetcher (H2D)", D)

def E(n, _):
    cnt = n if _ else 1
    for i in range(cnt):
        x = host[i % bench.POOL].to(dev, non_blocking=True)   # same stream: copy then compute, no overlap
        net.reset_state_variables(); net.run({"X": x}, time=bench.T_STEPS)
        rb.push(net.monitors["Ae_spikes"].get("s"))
        if len(rb) == rb.depth: rb.pop().numpy().sum()

timed("E serial H2D on the compute stream", E)
