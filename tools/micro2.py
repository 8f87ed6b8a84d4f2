#!/usr/bin/env python3
"""Developer microbenchmarks, round 2: per-step cost and per-strip start-up lag of the boundary-mode strip kernels
(lag-2 kernel of nw_lag2.cuh against the one-column-skew kernel of nw_packed.cuh), then the fixture pairs.
    python tools/micro2.py [quick]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
nw = importlib.import_module("fast-needleman-wunsch_b200")
GHZ = 1.965

def pair(n1, n2, seed=1):
    rng = np.random.default_rng(seed)
    return rng.integers(1, 5, size=n1, dtype=np.int8), rng.integers(1, 5, size=n2, dtype=np.int8)

def t_plan(s1, s2, R, warps=0, ctas=0, mode=0, iters=3):
    with nw.Plan(s1.size, s2.size, mode=mode, rows_per_lane=R, warps_per_cta=warps, ctas=ctas) as p:
        p.upload(s1, s2); p.time(1); ms = p.time(iters)
        return ms, p.strip_info() if mode != 2 else {}

def step_and_lag(R, lag2):
    os.environ["NW_CUDA_LAG2"] = str(lag2)
    SH = 32 * R
    # one warp alone, one strip: cycles per column step
    s1, s2 = pair(1 << 19, SH)
    ms1, _ = t_plan(s1, s2, R, warps=1, ctas=1)
    c1 = ms1 * 1e-3 * GHZ * 1e9 / s1.size
    # one warp per scheduler on every SM
    s1, s2 = pair(1 << 18, SH * 592)
    ms4, info = t_plan(s1, s2, R, warps=4)
    # lag: 592 strips, two widths -> t = (n1 + S * lag) * c
    ts = []
    for n1 in (4096, 32768):
        a, b = pair(n1, SH * 592, seed=2)
        ms, _ = t_plan(a, b, R, warps=4, iters=5)
        ts.append(ms * 1e-3 * GHZ * 1e9)
    c = (ts[1] - ts[0]) / (32768 - 4096)
    lag = (ts[0] / c - 4096) / 592
    print(f"lag2={lag2} R={R:2d}: 1 warp alone {c1:6.1f} cyc/col | 592 strips x {1<<18} cols: {ms4:7.3f} ms = "
          f"{(1<<18)*SH*592/ms4/1e6:7.0f} GCUPS | fitted {c:5.1f} cyc/col, lag {lag:6.1f} cols/strip ({lag*c:6.0f} cycles)", flush=True)

nw.init(0)
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
for R in ((8,) if quick else (2, 4, 8, 16)):
    for lag2 in (1, 0):
        step_and_lag(R, lag2)

from conftest import BDNA
def fixture(name):
    sep = "-" if name.endswith("gb") else ""
    return (np.fromfile(os.path.join(BDNA, f"{name}{sep}1.bdna"), dtype=np.int8),
            np.fromfile(os.path.join(BDNA, f"{name}{sep}2.bdna"), dtype=np.int8))
for name in ("64gb", "big", "mid", "2gb"):
    s1, s2 = fixture(name)
    for lag2 in (1, 0):
        os.environ["NW_CUDA_LAG2"] = str(lag2)
        out = []
        for R in (0, 2, 4, 8, 16):
            ms, info = t_plan(s1, s2, R)
            out.append(f"R={info['rows_per_lane']:2d}{'*' if R == 0 else ' '} {ms:6.3f} ms")
        ms, _ = t_plan(s1, s2, 0, mode=2)
        print(f"{name:5s} lag2={lag2}: " + " | ".join(out) + f" | score-mode {ms:6.3f} ms  ({s1.size*s2.size/ms/1e6:.0f} GCUPS)", flush=True)
