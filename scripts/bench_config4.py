"""BASELINE.json config 4 (Input[1,32,32] -conv k5-> LIFNodes[16,28,28] -> Connection -> LIFNodes(10), MSTDP on both,
Bernoulli(0.1) input, B=128) on the generic window kernel, with the oracle (dense restatement, all host cores) timed
beside it on a bounded sample.  Diagnostic companion of bench.py for SURVEY.md §8a rows A11-A13; prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
import cases

B, T = 128, 100


def build(device):
    ns = cases.namespace("b200")
    torch.manual_seed(7)
    net = ns.Network(dt=1.0, batch_size=B)
    X = ns.nodes.Input(shape=[1, 32, 32], traces=True)
    H = ns.nodes.LIFNodes(shape=[16, 28, 28], traces=True)
    O = ns.nodes.LIFNodes(n=10, traces=True)
    net.add_layer(X, "X"); net.add_layer(H, "H"); net.add_layer(O, "O")
    net.add_connection(ns.topology.Conv2dConnection(X, H, kernel_size=5, update_rule=ns.learning.MSTDP, nu=1e-2,
                                                    reduction=torch.sum, wmin=-1.0, wmax=1.0), "X", "H")
    net.add_connection(ns.topology.Connection(H, O, update_rule=ns.learning.MSTDP, nu=1e-2, reduction=torch.sum,
                                              wmin=-1.0, wmax=1.0), "H", "O")
    return net.to(device) if device != "cpu" else net


g = torch.Generator().manual_seed(11)
x = torch.bernoulli(0.1 * torch.ones(T, B, 1, 32, 32), generator=g).byte()
net = build("cuda")
xd = x.cuda()
for _ in range(2):
    net.reset_state_variables(); net.run({"X": xd}, time=T, reward=1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K):
    net.reset_state_variables(); net.run({"X": xd}, time=T, reward=1.0)
e1.record(); torch.cuda.synchronize()
net.check_errors()
ms = e0.elapsed_time(e1) / K
gpu_rate = B * T / (ms * 1e-3)

from oracle.oracle import OracleBackend
ref = build("cpu")
Ts = 10
with OracleBackend(dense=1):
    ref.run({"X": x[:Ts]}, time=Ts, reward=1.0)
    t0 = time.perf_counter()
    ref.reset_state_variables(); ref.run({"X": x[:Ts]}, time=Ts, reward=1.0)
    wall = time.perf_counter() - t0
cpu_rate = B * Ts / wall
print(json.dumps({"workload": "BASELINE config 4: conv32x32 k5 -> 16x28x28 -> 10, MSTDP, B=128", "tier": "generic window kernel",
                  "gpu_sample_timesteps_per_s": gpu_rate, "gpu_ms_per_timestep": ms / T,
                  "cpu_port_sample_timesteps_per_s": cpu_rate, "cpu_cores": os.cpu_count(), "cpu_sample": f"{Ts} of {T} timesteps, dense mode"}))
