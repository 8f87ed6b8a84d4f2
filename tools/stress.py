"""Randomised stress of every mode against the oracle (developer tool; the pytest suite holds the fixed cases)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Oracle
nw = importlib.import_module("fast-needleman-wunsch_b200")
nw.init(0)
orc = Oracle()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 12345)
t_end = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 120)
n = 0
while time.time() < t_end:
    kind = rng.integers(0, 6)
    hi = int(rng.choice([2, 3, 5, 5, 5, 5, 9, 100]))
    if kind <= 3:
        n1, n2 = int(rng.integers(0, 6000)), int(rng.integers(0, 6000))
        if rng.random() < 0.2: n1 = int(rng.integers(0, 70000))
        if rng.random() < 0.2: n2 = int(rng.integers(0, 30000))
        s1 = rng.integers(1, hi, size=n1, dtype=np.int8); s2 = rng.integers(1, hi, size=n2, dtype=np.int8)
        if rng.random() < 0.3 and min(n1, n2) > 10:        # long common stretch: fastest growth of G
            k = int(rng.integers(1, min(n1, n2))); s2[:k] = s1[:k]
        R = int(rng.choice([0, 0, 1, 2, 4, 8, 16]))
        os.environ["NW_CUDA_K2"] = str(int(rng.integers(0, 2)))
        if hi > 5 and R == 16: R = 8
        row, col, sc, _ = orc.boundaries(s1, s2)
        if kind == 0:
            with nw.Plan(n1, n2, rows_per_lane=R) as p:
                p.upload(s1, s2); p.run()
                assert np.array_equal(p.last_row(), row) and np.array_equal(p.last_col(), col) and p.score() == sc, ("boundary", n1, n2, hi, R)
        elif kind == 1:
            with nw.Plan(n1, n2, mode=nw.NW_MODE_SCORE, rows_per_lane=R) as p:
                p.upload(s1, s2); p.run(); assert p.score() == sc, ("score", n1, n2, hi, R)
        elif kind == 2 and (n1 + 1) * (n2 + 1) < 60_000_000:
            os.environ["NW_CUDA_TILE_BLOCKS"] = str(int(rng.choice([2, 5, 32])))
            with nw.Plan(n1, n2, mode=nw.NW_MODE_FULL, rows_per_lane=R) as p:
                p.upload(s1, s2); p.run()
                assert np.array_equal(p.table_to_host(), orc.fill(s1, s2)), ("full", n1, n2, hi, R)
        elif kind == 3 and n1 >= 64:
            P = int(rng.choice([2, 3, 5]))
            plans = [nw.Plan(n1, n2, part=k, nparts=P, rows_per_lane=R if R else 4) for k in range(P)]
            for a, b in zip(plans, plans[1:]): a.connect(b)
            for p in plans: p.upload(s1, s2)
            for p in plans: p.run(); p.sync()
            assert plans[-1].score() == sc and np.array_equal(plans[-1].last_col(), col), ("strips", n1, n2, hi, R, P)
            for p in plans: p.close()
    else:
        npairs, l1, l2 = int(rng.integers(1, 400)), int(rng.integers(0, 2500)), int(rng.integers(0, 2500))
        S1 = rng.integers(1, hi, size=(npairs, l1), dtype=np.int8); S2 = rng.integers(1, hi, size=(npairs, l2), dtype=np.int8)
        assert np.array_equal(nw.batch_scores(S1, S2), orc.batch_scores(S1, S2)), ("batch", npairs, l1, l2, hi)
    n += 1
print(f"stress ok: {n} random cases")
