"""Experiment: score of a pair by meeting in the middle -- forward fill of the top half and forward fill of the REVERSED
sequences for the bottom half run concurrently (two plans, two streams); score = max_j F[m][j] + B[m][j]."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nw = importlib.import_module("fast-needleman-wunsch_b200")
nw.init(0)
B = os.path.join(ROOT, "oracle", "_ref", "bdna")
name = sys.argv[1] if len(sys.argv) > 1 else "64gb"
a, b = (f"{B}/{name}-1.bdna", f"{B}/{name}-2.bdna") if name.endswith("gb") else (f"{B}/{name}1.bdna", f"{B}/{name}2.bdna")
s1 = np.fromfile(a, dtype=np.int8); s2 = np.fromfile(b, dtype=np.int8)
n1, n2 = s1.size, s2.size
m = n2 // 2
top = nw.Plan(n1, m); bot = nw.Plan(n1, n2 - m)
top.upload(s1, s2[:m].copy()); bot.upload(s1[::-1].copy(), s2[m:][::-1].copy())
for rep in range(3):
    top.sync(); bot.sync()
    t0 = time.perf_counter()
    for _ in range(5):
        top.run(); bot.run()
    top.sync(); bot.sync()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: both halves concurrently {dt*1e3:.3f} ms per pair -> {n1*n2/dt/1e9:.1f} GCUPS   (top alone {top.last_ms():.3f} ms, bottom alone-ish {bot.last_ms():.3f} ms)")
F = top.last_row().astype(np.int64); Bk = bot.last_row().astype(np.int64)[::-1]
print("score", int((F + Bk).max()), top.strip_info(), bot.strip_info())
with nw.Plan(n1, n2) as p:
    p.upload(s1, s2); p.time(1); print("single chain:", p.time(3), "ms, score", p.score())
