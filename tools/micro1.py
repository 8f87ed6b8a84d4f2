import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nw = importlib.import_module("fast-needleman-wunsch_b200")
nw.init(0)
rng = np.random.default_rng(1)
for R, warps in [(4, 1), (8, 1), (8, 4), (16, 1)]:
    n1, n2 = 1 << 19, 32 * R * warps
    s1 = rng.integers(1, 5, size=n1, dtype=np.int8); s2 = rng.integers(1, 5, size=n2, dtype=np.int8)
    with nw.Plan(n1, n2, rows_per_lane=R, warps_per_cta=warps) as p:
        p.upload(s1, s2); p.time(1); ms = p.time(3)
        print(f"{os.environ.get('NW_CUDA_LIB','default')[-16:]} R={R} warps={warps} cycles/col={ms*1e-3*1.965e9/n1:.1f}", flush=True)
