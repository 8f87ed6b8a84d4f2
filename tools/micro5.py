#!/usr/bin/env python3
"""592 strips x 65536 columns trace only (bisecting hand-off delays)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nw = importlib.import_module("fast-needleman-wunsch_b200")
GHZ = 1.965
nw.init(0)
rng = np.random.default_rng(1)
s1 = rng.integers(1, 5, size=65536, dtype=np.int8); s2 = rng.integers(1, 5, size=256 * 592, dtype=np.int8)
for lag2 in (1, 0):
    os.environ["NW_CUDA_LAG2"] = str(lag2)
    with nw.Plan(s1.size, s2.size, rows_per_lane=8, warps_per_cta=4) as p:
        p.upload(s1, s2); p.time(1); p.run(); p.sync()
        a, b = p.strip_times()
        lag = np.diff(a) * GHZ
        dur = (b - a) * GHZ / s1.size
        q = lambda x: " ".join(f"{v:8.0f}" for v in np.percentile(x, [0, 10, 50, 90, 100]))
        print(f"{sys.argv[1] if len(sys.argv) > 1 else ''} lag2={lag2}: total {(b[-1]-a[0])*1e-6:.3f} ms; start lag p0/10/50/90/100: {q(lag)} | cyc/col: {q(dur)} | by mod 4:",
              [int(np.median(lag[(np.arange(1, 592) % 4) == k])) for k in range(4)], flush=True)
