#!/usr/bin/env python3
"""One table-free alignment of a fixture pair, for ncu / timing:  python tools/align_case.py [pair] [repeats]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
nw = importlib.import_module("fast-needleman-wunsch_b200")
from conftest import BDNA, GOLDEN
name = sys.argv[1] if len(sys.argv) > 1 else "mid"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sep = "-" if name.endswith("gb") else ""
s1 = np.fromfile(os.path.join(BDNA, f"{name}{sep}1.bdna"), dtype=np.int8)
s2 = np.fromfile(os.path.join(BDNA, f"{name}{sep}2.bdna"), dtype=np.int8)
nw.init(0)
for _ in range(reps):
    t0 = time.perf_counter()
    a1, a2, sc = nw.align(s1, s2)
    dt = time.perf_counter() - t0
    ok = sc == GOLDEN["fixtures"][name]["score"] and np.array_equal(a1[a1 != 0], s1) and np.array_equal(a2[a2 != 0], s2)
    print(f"{name}: align {dt*1e3:.1f} ms wall, score {sc}, {a1.size} columns, {'ok' if ok else 'MISMATCH'}", flush=True)
