#!/usr/bin/env python3
"""Column strips over the GPUs of ONE process (nw_plan_connect): per-part strip trace of one 64gb fill.
    python tools/mgpu_trace.py [ngpus] [lag2]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
nw = importlib.import_module("fast-needleman-wunsch_b200")
from conftest import BDNA
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
if len(sys.argv) > 2: os.environ["NW_CUDA_LAG2"] = sys.argv[2]
os.environ["NW_CUDA_GPUS"] = str(G)
for d in range(G): nw.init(d)
s1 = np.fromfile(os.path.join(BDNA, "64gb-1.bdna"), dtype=np.int8); s2 = np.fromfile(os.path.join(BDNA, "64gb-2.bdna"), dtype=np.int8)
plans = [nw.Plan(s1.size, s2.size, device=d, part=d, nparts=G, rows_per_lane=8) for d in range(G)]
for a, b in zip(plans, plans[1:]): a.connect(b)
for p in plans: p.upload(s1, s2)
for rep in range(3):
    for p in plans: p.run()
    for p in plans: p.sync()
print("score", plans[-1].score(), "device ms per part:", [round(p.last_ms(), 3) for p in plans])
for d, p in enumerate(plans):
    a, b, cyc = p.strip_times(cycles=True)
    print(f"part {d}: ncols {p.ncols}; first start..last end {(b[-1]-a[0])*1e-6:.3f} ms; median start lag {np.median(np.diff(a))*1.965:.0f} cyc; "
          f"pace {np.median(cyc)/p.ncols:.1f} SM cyc/col; strip durations ms p0/50/100 {np.percentile((b-a)*1e-6,[0,50,100]).round(3)}")
