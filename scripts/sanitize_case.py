"""compute-sanitizer target: one short window of the metric geometry (DiehlAndCook2015 n=1600, B=128) on a given tier,
plus a conv / MSTDP window on the generic kernel.  Usage under the sanitizer:
    compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_case.py <tier> [T]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import bench, cases, helpers
from bindsnet_b200 import _backend

tier = int(sys.argv[1]); T = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda", 0)
if tier in (2, 3):
    net = bench.make_network(dev)
    net.force_tier = tier
    x = bench.synth_windows(1, seed=5, T=T)[0].to(dev)
    for _ in range(2):
        net.reset_state_variables()
        net.run({"X": x}, time=T)
    net.check_errors()
    torch.cuda.synchronize()
    print("tier", _backend.last_tier, "ok; Ae spikes/window", int(net.layers["Ae"].s.sum()))
else:
    for name in ("conv_mstdp", "conv_postpre", "hebbian_dense", "dc2015_onespike"):
        fx = helpers.Fixture(name)
        n, inputs, kw, Tc = fx.build("cuda")
        n.force_tier = 1
        n.run(inputs={k: v.cuda() for k, v in inputs.items()}, time=min(Tc, T), one_spike_seed=cases.ONE_SPIKE_SEED, **kw)
        n.check_errors()
    torch.cuda.synchronize()
    print("generic tier ok")
