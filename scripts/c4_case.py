"""Profiling target: short windows of BASELINE config 4 (conv 32x32 k5 -> 16x28x28 -> 10, MSTDP, B=128) on the generic kernel.
    python scripts/c4_case.py [T] [windows]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bindsnet_b200.learning import MSTDP
from bindsnet_b200.network import Network, nodes, topology

T = int(sys.argv[1]) if len(sys.argv) > 1 else 30
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = 128
dev = torch.device("cuda", 0)
torch.manual_seed(7)
net = Network(dt=1.0, batch_size=B)
X = nodes.Input(shape=[1, 32, 32], traces=True)
H = nodes.LIFNodes(shape=[16, 28, 28], traces=True)
O = nodes.LIFNodes(n=10, traces=True)
net.add_layer(X, "X"); net.add_layer(H, "H"); net.add_layer(O, "O")
net.add_connection(topology.Conv2dConnection(X, H, kernel_size=5, update_rule=MSTDP, nu=1e-2, reduction=torch.sum, wmin=-1.0, wmax=1.0), "X", "H")
net.add_connection(topology.Connection(H, O, update_rule=MSTDP, nu=1e-2, reduction=torch.sum, wmin=-1.0, wmax=1.0), "H", "O")
net.to(dev)
g = torch.Generator().manual_seed(11)
x = torch.bernoulli(0.1 * torch.ones(T, B, 1, 32, 32), generator=g).byte().to(dev)
for _ in range(W):
    net.run({"X": x}, time=T, reward=1.0)
net.check_errors()
torch.cuda.synchronize()
print("ok; H spikes in the last step per sample", float(net.layers["H"].s.float().sum()) / B, "O", float(net.layers["O"].s.float().sum()) / B)
