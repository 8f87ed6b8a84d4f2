#!/usr/bin/env python3
"""Round-2 developer probes: (1) how warps of one SM slow each other down, (2) per-strip start/end trace of a chained
fill (start-up lag, per-strip speed).   python tools/micro3.py"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nw = importlib.import_module("fast-needleman-wunsch_b200")
GHZ = 1.965
nw.init(0)

def pair(n1, n2, seed=1):
    rng = np.random.default_rng(seed)
    return rng.integers(1, 5, size=n1, dtype=np.int8), rng.integers(1, 5, size=n2, dtype=np.int8)

print("# (1) S strips of 256 rows in ONE CTA of S warps, 262144 columns: cycles per column")
for lag2 in (1, 0):
    os.environ["NW_CUDA_LAG2"] = str(lag2)
    row = []
    for S in (1, 2, 4, 8, 16):
        s1, s2 = pair(1 << 18, 256 * S)
        with nw.Plan(s1.size, s2.size, rows_per_lane=8, warps_per_cta=S, ctas=1) as p:
            p.upload(s1, s2); p.time(1); ms = p.time(2)
            a, b = p.strip_times()
            dur = (b - a) * GHZ / s1.size          # cycles per column of each strip by its own stamps
            row.append(f"S={S:2d}: {ms*1e-3*GHZ*1e9/s1.size:6.1f} (per strip {dur.min():.1f}..{dur.max():.1f})")
    print(f"lag2={lag2}: " + " | ".join(row), flush=True)

print("# (2) 592 strips x 65536 columns, one warp per scheduler on every SM: trace")
for lag2 in (1, 0):
    os.environ["NW_CUDA_LAG2"] = str(lag2)
    s1, s2 = pair(65536, 256 * 592)
    with nw.Plan(s1.size, s2.size, rows_per_lane=8, warps_per_cta=4) as p:
        p.upload(s1, s2); p.time(1); p.run(); p.sync()
        a, b = p.strip_times()
        t0 = a[0]
        lag = np.diff(a) * GHZ            # cycles between consecutive strip starts
        dur = (b - a) * GHZ / s1.size     # cycles per column of each strip
        endlag = np.diff(b) * GHZ
        q = lambda x: " ".join(f"{v:8.0f}" for v in np.percentile(x, [0, 10, 50, 90, 100]))
        print(f"lag2={lag2}: total {(b[-1]-t0)*1e-6:.3f} ms; start lag cycles p0/10/50/90/100: {q(lag)} | end lag: {q(endlag)} | "
              f"cycles/col per strip: {q(dur)}")
        print("   start lag by strip index mod 4 (0 = first warp of a CTA, its predecessor is on another SM):",
              [int(np.median(lag[(np.arange(1, 592) % 4) == k])) for k in range(4)])
        print("   cycles/col strips 0,1,2,3,100,300,591:", [round(float(dur[i]), 1) for i in (0, 1, 2, 3, 100, 300, 591)], flush=True)
