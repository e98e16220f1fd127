"""Manual check (a tool, not part of the suite: ~3.5 min): BASELINE config 3 at full size (n = 6400, B = 256, T = 70: the network of
tests/test_gpu_c3.py) through the EMULATED generic kernel, bit for bit against the oracle.  python tests/emu/c3_full.py [T]
profiles/emu_baseline_sizes_r2.txt holds the round-2 run."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("", "tests", "tests/golden", "tests/emu"): sys.path.insert(0, os.path.join(ROOT, p))
import torch, cases, helpers, emu
import test_gpu_c3 as C3
from oracle.oracle import OracleBackend
T = int(sys.argv[1]) if len(sys.argv) > 1 else C3.T
C3.T = T
outs = []
for be, tier in ((emu.EmuBackend, 1), (OracleBackend, 0)):
    t0 = time.time()
    net, inputs = C3._build("cpu")
    net.force_tier = tier
    helpers.add_spike_monitors(net, T)
    with be() as b_:
        net.run(inputs=inputs, time=T, one_spike_seed=cases.ONE_SPIKE_SEED); assert b_.err == 0
    outs.append((helpers.snapshot(net), helpers.spike_counts(net, T)))
    print(be.__name__, "%.0f s" % (time.time() - t0), flush=True)
print("Ae spikes", int(outs[1][1]["L/Ae/count"].sum()))
helpers.assert_bit_identical(outs[0][0], outs[1][0], "c3 state"); helpers.assert_bit_identical(outs[0][1], outs[1][1], "c3 counts")
print("bit-identical: BASELINE config 3 (n = 6400, B = 256, T = %d) through the emulated generic kernel" % T)
