#!/usr/bin/env python3
"""Developer microbenchmarks of the strip kernel on synthetic shapes: per-column step time (one long strip set) and
per-strip start-up lag (many short strips).   python tools/micro.py"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nw = importlib.import_module("fast-needleman-wunsch_b200")

def run(n1, n2, R, warps=8, ctas=0, tag=""):
    rng = np.random.default_rng(1)
    s1 = rng.integers(1, 5, size=n1, dtype=np.int8); s2 = rng.integers(1, 5, size=n2, dtype=np.int8)
    with nw.Plan(n1, n2, rows_per_lane=R, warps_per_cta=warps, ctas=ctas) as p:
        p.upload(s1, s2); p.time(1); ms = p.time(3)
        info = p.strip_info()
        cyc = ms * 1e-3 * 1.965e9
        print(f"{tag} n1={n1} n2={n2} R={R} strips={info['nstrips']} ctas={info['ctas']}x{info['warps']}w ms={ms:.3f} "
              f"GCUPS={n1*n2/ms/1e6:.1f} cycles/col={cyc/n1:.1f} cycles/(col+32*strips)={cyc/(n1+32*info['nstrips']):.1f}", flush=True)
        return ms

nw.init(0)
for R in (4, 8, 16):
    run(1 << 20, 32 * R, R, warps=1, tag="1 warp alone       ")
    run(1 << 20, 32 * R * 4, R, warps=4, tag="4 warps, 1/SMSP    ")
    run(1 << 20, 32 * R * 8, R, warps=8, tag="8 warps, 2/SMSP    ")
    run(1 << 20, 32 * R * 16, R, warps=16, tag="16 warps, 4/SMSP   ")
    # lag: many strips, few columns: time ~ (n1 + strips*d) * t_col
    for n1 in (2048, 8192):
        run(n1, 32 * R * 1000, R, warps=8, tag="lag 1000 strips    ")
