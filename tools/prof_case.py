#!/usr/bin/env python3
"""One boundary-mode fill of a synthetic pair, for ncu:  python tools/prof_case.py n1 n2 R warps ctas [iters]
(kernel choice through the environment: NW_CUDA_LAG2=0/1, NW_CUDA_NO_PACKED=1, ...)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nw = importlib.import_module("fast-needleman-wunsch_b200")
n1, n2, R, warps, ctas = (int(x) for x in sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 2
rng = np.random.default_rng(1)
s1 = rng.integers(1, 5, size=n1, dtype=np.int8); s2 = rng.integers(1, 5, size=n2, dtype=np.int8)
nw.init(0)
with nw.Plan(n1, n2, rows_per_lane=R, warps_per_cta=warps, ctas=ctas) as p:
    p.upload(s1, s2)
    ms = p.time(iters)
    print(f"n1={n1} n2={n2} {p.strip_info()} lag2={os.environ.get('NW_CUDA_LAG2', '1')} ms={ms:.3f} "
          f"cycles/col={ms * 1e-3 * 1.965e9 / n1:.1f} score={p.score()}")
