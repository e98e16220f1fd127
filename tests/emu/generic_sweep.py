"""Exploratory sweep (a tool, not part of the test suite): the random networks of tests/test_oracle_fuzz_vs_reference.py
for a range of seeds — node kinds x learning rules x reductions x options x batch sizes, one or two learned
connections — through the EMULATED generic kernel under random thread schedules and grid sizes, bit for bit against the
oracle.
    python tests/emu/generic_sweep.py <first seed> <count>
SWEEP_LARGE=1 redraws the sizes from wider ranges (up to 900 inputs, 700 neurons, batch 80, short windows): several tiles, sample
chunks and row chunks per layer / connection, the any-spike flags of wide sources, batches that do not fill a warp."""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("", "tests", os.path.join("tests", "golden"), os.path.join("tests", "emu")):
    sys.path.insert(0, os.path.join(ROOT, p))
import cases, helpers, emu
import test_oracle_fuzz_vs_reference as fuzz
from oracle.oracle import OracleBackend

first, count = int(sys.argv[1]), int(sys.argv[2])
ns = cases.namespace("b200")
rng = random.Random(first)
bad = 0
for seed in range(first, first + count):
    spec = fuzz._draw(seed)
    if os.environ.get("SWEEP_LARGE"):
        spec.update(n_in=rng.choice([33, 100, 257, 640, 900]), n_hid=rng.choice([31, 64, 130, 333, 700]), n_out=rng.choice([6, 40, 150]),
                    B=rng.choice([1, 7, 32, 33, 80]), T=rng.choice([8, 14, 25]), p_in=rng.choice([0.02, 0.1, 0.3]))
        if spec["norm"] is not None: spec["norm"] = spec["norm"] * spec["n_in"] / 50.0
    sh, sms = rng.choice([None, "1", "5"]), rng.choice(["1", "2", "3", "6"])
    os.environ["SNN_EMU_SMS"] = sms
    if sh: os.environ["SNN_EMU_SHUFFLE"] = sh
    else: os.environ.pop("SNN_EMU_SHUFFLE", None)
    outs, t0 = [], time.time()
    for backend in (emu.EmuBackend, OracleBackend):
        net, x = fuzz._build(ns, spec)
        net.force_tier = 1
        helpers.add_spike_monitors(net, spec["T"])
        with backend() as be:
            net.run(inputs={"X": x}, time=spec["T"])
            assert be.err == 0
        outs.append((helpers.snapshot(net), helpers.spike_counts(net, spec["T"])))
    try:
        helpers.assert_bit_identical(outs[0][0], outs[1][0], "state"); helpers.assert_bit_identical(outs[0][1], outs[1][1], "counts")
        status = "ok"
    except AssertionError as e:
        status, bad = "MISMATCH " + str(e)[:120], bad + 1
    print(f"{seed:4d} {spec['kind'][:10]:10s} {spec['rule'][:8]:8s} B={spec['B']} T={spec['T']:3d} n={spec['n_in']}/{spec['n_hid']} second={spec['second']} sms={sms} sh={sh} rate={outs[1][1]['L/Y/count'].sum() / (spec['T'] * spec['B'] * spec['n_hid']):.3f} "
          f"{time.time() - t0:5.1f}s {status}", flush=True)
print("bad:", bad)
