"""Exploratory sweep (a tool, not part of the test suite): random DiehlAndCook2015-shaped networks — neurons, batch size,
window length, rule, reduction, traces, bounds, decay, one_spike, inhibition strength, input density, thread schedule —
through the EMULATED fused kernel, bit for bit against the oracle.
    python tests/emu/fused_sweep.py <seed> <count> [tier = 2 | 3]
Configurations the forced tier does not take (tier 3: PostPre, sum, B <= 128 ...) are reported as "skip".
profiles/emu_fused_sweep_r2.txt holds the round-2 runs (seeds 1-4, 480 configurations, 0 mismatches)."""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("", "tests", os.path.join("tests", "golden"), os.path.join("tests", "emu")):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, numpy as np
import cases, helpers, emu
import test_gpu_variants as V
from oracle.oracle import OracleBackend
ns = cases.namespace("b200")
L = ns.learning
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
TIER = int(sys.argv[3]) if len(sys.argv) > 3 else 2
bad = 0
for it in range(N):
    n = rng.choice([17, 20, 33, 48, 64, 75, 100, 130, 200, 260])
    B = rng.choice([1, 2, 3, 7, 16, 31, 33, 64, 100, 129, 160, 200])
    T = rng.choice([8, 15, 30, 45])
    rule = rng.choice([L.PostPre, L.PostPre, L.WeightDependentPostPre])
    red = rng.choice([torch.sum, torch.mean])
    kw = dict(rule=rule, reduction=red, nu=(rng.choice([1e-4, 1e-3, 2e-3]), rng.choice([0.0, 1e-2, 5e-2])), additive=rng.random() < 0.3,
              lbound=(-70.0 if rng.random() < 0.3 else None), weight_decay=(1e-3 if rng.random() < 0.2 else 0.0), one_spike=rng.random() < 0.8,
              inh=rng.choice([120.0, 17.5, 3.0]), w_seed=rng.randrange(1000))
    if os.environ.get("SWEEP_EDGE"):   # weights next to wmin / wmax: the clamp is active on many rows
        kw.update(tiny=rng.choice([0.0, 0.3, 0.6]), huge=rng.choice([0.0, 0.0, 0.2]), norm=rng.choice([78.4, None]))
    if int(os.environ.get('SWEEP_DRAW', TIER)) == 3:   # the lean option set the column-group kernel takes
        kw.update(rule=L.PostPre, reduction=torch.sum, additive=False, lbound=None, weight_decay=0.0)
        B = min(B, 128)
        n = rng.choice([n, 300, 450, 640])
    p = rng.choice([0.03, 0.08, 0.2])
    # SWEEP_WINDOWS=1: one to three consecutive windows without a reset in between (membrane, refractory counters, traces —
    # the Input layer's included — and theta carry over), learning switched off for one of them
    NW = rng.choice([1, 2, 3]) if os.environ.get("SWEEP_WINDOWS") else 1
    frozen = rng.randrange(NW + 1) if NW > 1 else -1
    shuffle = rng.choice([None, "1", "7"])
    if shuffle: os.environ["SNN_EMU_SHUFFLE"] = shuffle
    else: os.environ.pop("SNN_EMU_SHUFFLE", None)
    if os.environ.get("SWEEP_ONLY") and int(os.environ["SWEEP_ONLY"]) != it: continue   # re-run one line of a sweep
    if os.environ.get("SWEEP_ONLY"): print("   kw:", {k: (v.__name__ if callable(v) else v) for k, v in kw.items()}, "x seed", 1000 + it)
    if os.environ.get("SWEEP_SHUFFLE"): os.environ["SNN_EMU_SHUFFLE"] = os.environ["SWEEP_SHUFFLE"]
    outs = []
    t0 = time.time()
    try:
        for be, tier in ((emu.EmuBackend, TIER), (OracleBackend, 0)):
            torch.manual_seed(99)
            net = V._graph(ns, n, B, **kw)
            net.force_tier = tier
            helpers.add_spike_monitors(net, T)
            with be() as b_:
                for wi in range(NW):
                    x = cases._bernoulli_inputs(T, B, (1, 28, 28), p, 1000 + it + 7919 * wi)
                    net.train(wi != frozen)
                    net.run(inputs={"X": x}, time=T, one_spike_seed=cases.ONE_SPIKE_SEED + wi); assert b_.err == 0
                    if wi + 1 < NW: outs.append((helpers.snapshot(net), helpers.spike_counts(net, T)))
            outs.append((helpers.snapshot(net), helpers.spike_counts(net, T)))
        for wi in range(NW):
            helpers.assert_bit_identical(outs[wi][0], outs[NW + wi][0], f"state after window {wi}")
            helpers.assert_bit_identical(outs[wi][1], outs[NW + wi][1], f"counts of window {wi}")
        status = "ok"
    except AssertionError as e:
        status = "MISMATCH " + str(e)[:120]; bad += 1
    except Exception as e:
        status = "ERR " + type(e).__name__ + " " + str(e)[:100]
        if "not implemented" in str(e) or "tier" in str(e): status = "skip (" + str(e)[:60] + ")"
        else: bad += 1
    print(f"{it:3d} n={n:3d} B={B:3d} T={T:2d} {rule.__name__[:8]:8s} {red.__name__:4s} os={kw['one_spike']} sh={shuffle} p={p} w={NW}/{frozen} Ae={int(outs[-1][1]['L/Ae/count'].sum()) if outs else -1:5d} {time.time()-t0:5.1f}s {status}", flush=True)
print("bad:", bad)
