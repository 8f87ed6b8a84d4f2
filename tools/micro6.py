#!/usr/bin/env python3
"""Per-strip pace along the chain (592 strips x 131072 columns)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nw = importlib.import_module("fast-needleman-wunsch_b200")
GHZ = 1.965
nw.init(0)
rng = np.random.default_rng(1)
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
S = int(sys.argv[2]) if len(sys.argv) > 2 else 592
s1 = rng.integers(1, 5, size=n1, dtype=np.int8); s2 = rng.integers(1, 5, size=256 * S, dtype=np.int8)
with nw.Plan(s1.size, s2.size, rows_per_lane=8, warps_per_cta=4) as p:
    p.upload(s1, s2); p.time(1); p.run(); p.sync()
    a, b = p.strip_times()
    dur = (b - a) * GHZ / s1.size
    lag = np.diff(a) * GHZ; elag = np.diff(b) * GHZ
    print(f"total {(b[-1]-a[0])*1e-6:.3f} ms")
    print("pace strips 0..47:", " ".join(f"{x:.1f}" for x in dur[:48]))
    print("pace every 37th:", " ".join(f"{x:.1f}" for x in dur[::37]))
    print("start lag every 37th:", " ".join(f"{x:.0f}" for x in lag[::37]))
    print("end-lag minus start-lag (cycles), strips 1..24:", " ".join(f"{x:.0f}" for x in (elag - lag)[:24]))
