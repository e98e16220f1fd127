"""Profiling target: short windows of a BASELINE configuration on the GENERIC kernel (tier 1).
    python scripts/generic_case.py metric|c3 [T] [windows]
(under ncu: `ncu --set full --import-source on --clock-control none -k regex:snn_generic_window -c 1 ...`)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from bindsnet_b200 import _backend

cfg = sys.argv[1] if len(sys.argv) > 1 else "metric"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
W = int(sys.argv[3]) if len(sys.argv) > 3 else 2
bench.apply_config(cfg)
dev = torch.device("cuda", 0)
net = bench.make_network(dev)
net.force_tier = 1
x = bench.synth_windows(1, seed=5, T=T, B=bench.BATCH)[0].to(dev)
for _ in range(W):
    net.run({"X": x}, time=T)      # no reset: the second window starts with charged membranes and spikes
net.check_errors()
torch.cuda.synchronize()
print("tier", _backend.last_tier, "ok; Ae spikes in the last step", int(net.layers["Ae"].s.sum()))
