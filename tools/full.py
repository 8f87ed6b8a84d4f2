import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nw = importlib.import_module("fast-needleman-wunsch_b200")
nw.init(0)
B = os.path.join(ROOT, "oracle", "_ref", "bdna")
s1 = np.fromfile(f"{B}/2gb-1.bdna", dtype=np.int8); s2 = np.fromfile(f"{B}/2gb-2.bdna", dtype=np.int8)
cells = s1.size * s2.size
for R in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,2,4,8").split(",")]:
    for w in (4,):
        with nw.Plan(s1.size, s2.size, mode=nw.NW_MODE_FULL, rows_per_lane=R, warps_per_cta=w) as p:
            p.upload(s1, s2); p.time(1); ms = p.time(3)
            print(f"2gb FULL R={R} warps={w} {p.strip_info()} ms={ms:.3f} GCUPS={cells/ms/1e6:.1f} write GB/s={cells*4/ms/1e6:.0f}", flush=True)
t = np.empty((s2.size + 1, s1.size + 1), dtype=np.int32)
t0 = time.perf_counter(); nw.needlemanWunsch(s1, s2, t); dt = time.perf_counter() - t0
print(f"nw_cuda_fill_ex full, pageable host table: {dt*1e3:.1f} ms  ({cells/dt/1e9:.2f} GCUPS)  checksum {int(t.sum(dtype=np.int64))}")
