#!/usr/bin/env python3
"""Per-strip trace of one boundary-mode fill of a fixture pair: start-up lag between consecutive strips and pace
(cycles per column) of every strip, from the kernel's own stamps.   python tools/trace_pair.py [pair]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
nw = importlib.import_module("fast-needleman-wunsch_b200")
from conftest import BDNA, GOLDEN
GHZ = 1.965
name = sys.argv[1] if len(sys.argv) > 1 else "64gb"
sep = "-" if name.endswith("gb") else ""
s1 = np.fromfile(os.path.join(BDNA, f"{name}{sep}1.bdna"), dtype=np.int8)
s2 = np.fromfile(os.path.join(BDNA, f"{name}{sep}2.bdna"), dtype=np.int8)
nw.init(0)
with nw.Plan(s1.size, s2.size) as p:
    p.upload(s1, s2); p.time(2); ms = p.time(1)
    p.run(); p.sync()
    a, b = p.strip_times()
    n = a.size
    lag = np.diff(a) * GHZ
    dur = (b - a) * GHZ / s1.size
    q = lambda x: " ".join(f"{v:7.1f}" for v in np.percentile(x, [0, 10, 50, 90, 100]))
    print(f"{name} lib={os.path.basename(nw.lib_path)} {ms:.3f} ms score {p.score()} (golden {GOLDEN['fixtures'][name]['score']}) strips {n}")
    print(f"  start lag cycles p0/10/50/90/100: {q(lag)}  sum {lag.sum()/1e6:.2f} Mcyc")
    print(f"  cycles/col per strip           : {q(dur)}  strip0 {dur[0]:.1f} last {dur[-1]:.1f}")
    print(f"  total {(b[-1]-a[0])*GHZ/1e6:.2f} Mcyc = first start -> last end", flush=True)
