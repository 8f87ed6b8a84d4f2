#!/usr/bin/env python3
"""Pace of the first strips against the number of CTAs (= SMs) at work."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nw = importlib.import_module("fast-needleman-wunsch_b200")
GHZ = 1.965
nw.init(0)
rng = np.random.default_rng(1)
n1 = 262144
for warps in (4, 1):
    for ctas in (16, 37, 60, 74, 90, 111, 148):
        S = ctas * warps
        s1 = rng.integers(1, 5, size=n1, dtype=np.int8); s2 = rng.integers(1, 5, size=256 * S, dtype=np.int8)
        with nw.Plan(s1.size, s2.size, rows_per_lane=8, warps_per_cta=warps, ctas=ctas) as p:
            p.upload(s1, s2); p.run(); p.sync()
            a, b, cyc = p.strip_times(cycles=True)
            dur = (b - a) * GHZ / s1.size
            mhz = np.median(cyc / (b - a)) * 1e3
            print(f"warps/CTA={warps} ctas={ctas:3d} strips={S:3d}: pace strip0 {dur[0]:.1f} median {np.median(dur):.1f} last {dur[-1]:.1f}; SM cycles/col {np.median(cyc)/s1.size:.1f}; SM clock seen {mhz:.0f} MHz; total {(b[-1]-a[0])*1e-6:.2f} ms", flush=True)
