import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nw = importlib.import_module("fast-needleman-wunsch_b200")
nw.init(0)
rng = np.random.default_rng(1)
n1 = 262144
out = []
for warps, ctas in ((1, 16), (1, 148), (4, 37)):
    S = ctas * warps
    s1 = rng.integers(1, 5, size=n1, dtype=np.int8); s2 = rng.integers(1, 5, size=256 * S, dtype=np.int8)
    with nw.Plan(s1.size, s2.size, rows_per_lane=8, warps_per_cta=warps, ctas=ctas) as p:
        p.upload(s1, s2); p.run(); p.sync()
        a, b, cyc = p.strip_times(cycles=True)
        pace = cyc / n1
        out.append(f"{ctas}x{warps}: strip0 {pace[0]:.1f} strip1 {pace[1]:.1f} median {np.median(pace):.1f} last {pace[-1]:.1f}")
print(sys.argv[1], "SM cycles/col:", " | ".join(out), flush=True)
