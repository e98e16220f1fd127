"""Debug helper: run golden cases on the GPU (a given tier) against the oracle and list every differing key."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import cases, helpers
from oracle import oracle
oracle.build()

def run_gpu(name, tier):
    fx = helpers.Fixture(name)
    net, inputs, kw, T = fx.build("cuda")
    net.force_tier = tier
    helpers.add_spike_monitors(net, T, device="cuda")
    kw = {k: ({l: v.cuda() for l, v in d.items()} if isinstance(d, dict) else d) for k, d in kw.items()}
    torch.cuda.synchronize(); t0 = time.time()
    net.run(inputs={k: v.cuda() for k, v in inputs.items()}, time=T, one_spike_seed=cases.ONE_SPIKE_SEED, **kw)
    net.check_errors()
    dt = time.time() - t0
    return helpers.snapshot(net), helpers.spike_counts(net, T), dt

for name in sys.argv[2:]:
    t0 = time.time(); _, s_cpu, c_cpu = helpers.run_case_oracle(name); t_or = time.time() - t0
    for rep in range(2):
        s_gpu, c_gpu, dt = run_gpu(name, int(sys.argv[1]))
        bad = []
        for d_g, d_c in ((s_gpu, s_cpu), (c_gpu, c_cpu)):
            for k in d_g:
                a, b = d_g[k], d_c[k]
                if not np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b.view(np.uint32) if b.dtype == np.float32 else b):
                    diff = np.abs(a.astype(np.float64) - b.astype(np.float64))
                    idx = np.argwhere(diff > 0)[:6].tolist()
                    bad.append(f"{k}: {int((diff > 0).sum())} entries, max {diff.max():.3e}, first {idx}")
        print(f"{name} tier {sys.argv[1]} rep {rep}: gpu {dt:.2f}s oracle {t_or:.1f}s ->", "OK" if not bad else "DIFF", flush=True)
        for x in bad: print("   ", x, flush=True)
