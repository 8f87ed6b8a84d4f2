#!/usr/bin/env python3
"""A/B of one environment knob on the fixture pairs (boundary mode and score mode), scores checked against golden.json.
    python tools/ab.py NW_CUDA_WS 0 1 [pairs...]"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
nw = importlib.import_module("fast-needleman-wunsch_b200")
from conftest import BDNA, GOLDEN

def fixture(name):
    sep = "-" if name.endswith("gb") else ""
    return (np.fromfile(os.path.join(BDNA, f"{name}{sep}1.bdna"), dtype=np.int8),
            np.fromfile(os.path.join(BDNA, f"{name}{sep}2.bdna"), dtype=np.int8))

knob, vals = sys.argv[1], sys.argv[2:4]
pairs = sys.argv[4:] or ["64gb", "big", "mid", "2gb"]
nw.init(0)
for name in pairs:
    s1, s2 = fixture(name)
    want = GOLDEN["fixtures"][name]["score"]
    for v in vals:
        os.environ[knob] = v
        out = []
        for mode in (0, 2):
            with nw.Plan(s1.size, s2.size, mode=mode) as p:
                p.upload(s1, s2); p.time(2)
                ms = min(p.time(1) for _ in range(5))
                sc = p.score()
                out.append(f"mode {mode}: {ms:6.3f} ms {s1.size*s2.size/ms/1e6:6.0f} GCUPS score {sc}{'' if want is None or sc == want else ' MISMATCH want %d' % want}")
        print(f"{name:5s} {knob}={v}: " + " | ".join(out), flush=True)
