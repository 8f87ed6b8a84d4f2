"""Small run of every kernel (for compute-sanitizer): boundary (packed, packed two-column, 32-bit), full-table (two-pass
and one-pass), batch (packed and 32-bit), traceback, column strips on one device."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nw = importlib.import_module("fast-needleman-wunsch_b200")
nw.init(0)
rng = np.random.default_rng(3)
s1 = rng.integers(1, 5, size=700, dtype=np.int8); s2 = rng.integers(1, 5, size=520, dtype=np.int8)
ref = None
for env in ({}, {"NW_CUDA_K2": "1"}, {"NW_CUDA_NO_PACKED": "1"}):
    os.environ.pop("NW_CUDA_K2", None); os.environ.pop("NW_CUDA_NO_PACKED", None)
    os.environ.update(env)
    row, col, sc = nw.boundaries(s1, s2)
    t = nw.needlemanWunsch(s1, s2)
    assert t[-1, -1] == sc and np.array_equal(row, t[-1]) and np.array_equal(col, t[:, -1])
    if ref is None: ref = t
    assert np.array_equal(ref, t)
    S1 = rng.integers(1, 5, size=(40, 300), dtype=np.int8); S2 = rng.integers(1, 5, size=(40, 700), dtype=np.int8)
    b = nw.batch_scores(S1, S2)
    print(env, sc, b[:3])
os.environ.pop("NW_CUDA_NO_PACKED", None)
with nw.Plan(700, 520, mode=nw.NW_MODE_FULL) as p:
    p.upload(s1, s2); p.run(); a1, a2 = p.traceback(); print("traceback", a1.size)
plans = [nw.Plan(700, 520, part=k, nparts=2, rows_per_lane=4) for k in range(2)]
plans[0].connect(plans[1])
for p in plans: p.upload(s1, s2)
for p in plans: p.run(); p.sync()
assert plans[1].score() == ref[-1, -1]
print("ok")
