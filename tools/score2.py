#!/usr/bin/env python3
"""Score-only mode on one GPU against one half per GPU (NW_MODE_SCORE, part 0 of 2):  python tools/score2.py [pairs...]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
nw = importlib.import_module("fast-needleman-wunsch_b200")
from conftest import BDNA, GOLDEN
nw.init(0)
two = nw.device_count() >= 2
if two:
    nw.init(1)
for name in (sys.argv[1:] or ["64gb", "big", "mid", "2gb"]):
    sep = "-" if name.endswith("gb") else ""
    s1 = np.fromfile(os.path.join(BDNA, f"{name}{sep}1.bdna"), dtype=np.int8)
    s2 = np.fromfile(os.path.join(BDNA, f"{name}{sep}2.bdna"), dtype=np.int8)
    out = []
    for nparts in (1, 2) if two else (1,):
        with nw.Plan(s1.size, s2.size, mode=nw.NW_MODE_SCORE, part=0, nparts=nparts) as p:
            p.upload(s1, s2); p.time(2)
            ms = min(p.time(1) for _ in range(5))
            ok = p.score() == GOLDEN["fixtures"][name]["score"]
            out.append(f"{nparts} GPU: {ms:6.3f} ms {s1.size*s2.size/ms/1e6:6.0f} GCUPS {'ok' if ok else 'MISMATCH'}")
    print(f"{name:5s} score-only  " + " | ".join(out), flush=True)
