#!/usr/bin/env python3
"""64gb pair: rows per lane x warps per CTA x resident warps per SM, both packed boundary kernels."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
nw = importlib.import_module("fast-needleman-wunsch_b200")
from conftest import BDNA
nw.init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "64gb"
sep = "-" if name.endswith("gb") else ""
s1 = np.fromfile(os.path.join(BDNA, f"{name}{sep}1.bdna"), dtype=np.int8); s2 = np.fromfile(os.path.join(BDNA, f"{name}{sep}2.bdna"), dtype=np.int8)
for lag2 in (1, 0):
    os.environ["NW_CUDA_LAG2"] = str(lag2)
    for R, warps, wpsm in ((8, 4, 4), (4, 4, 8), (4, 8, 8), (2, 8, 16), (2, 16, 16), (4, 4, 4), (8, 8, 8)):
        os.environ["NW_CUDA_WARPS_PER_SM"] = str(wpsm)
        with nw.Plan(s1.size, s2.size, rows_per_lane=R, warps_per_cta=warps) as p:
            p.upload(s1, s2); p.time(1); ms = p.time(3)
            a, b = p.strip_times()
            info = p.strip_info()
            lag = np.median(np.diff(a)) * 1.965
            dur = np.median(b - a) * 1.965 / s1.size
            print(f"{name} lag2={lag2} R={R:2d} warps/CTA={warps:2d} warps/SM={wpsm:2d} ctas={info['ctas']:3d} strips={info['nstrips']:4d}: {ms:6.3f} ms "
                  f"{s1.size*s2.size/ms/1e6:6.0f} GCUPS | median start lag {lag:6.0f} cyc, {dur:5.1f} cyc/col per strip", flush=True)
