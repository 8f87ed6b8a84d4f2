"""Small run of every kernel family, for the bounds-check build (`make check`; compute-sanitizer is closed on the pool):
    NW_CUDA_LIB=build/libnw_check.so python tools/sanity_small.py
Boundary mode (lag-2, one-column skew, 32-bit, generic alphabet), full-table mode (two-pass and one-pass, streamed delivery),
score mode (horizontal cut and staircase), scoring parameters, Smith-Waterman, batches (packed and 32-bit), column strips on
one device, both tracebacks.  Results are compared with each other (the parity tests compare them with the oracle); an index
outside an allocation traps the kernel and fails the call."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nw = importlib.import_module("fast-needleman-wunsch_b200")
nw.init(0)
rng = np.random.default_rng(3)
KNOBS = ("NW_CUDA_LAG2", "NW_CUDA_NO_PACKED", "NW_CUDA_GENERIC", "NW_CUDA_NO_STAIR", "NW_CUDA_FORCE_STAIR", "NW_CUDA_NO_STREAMED",
         "NW_CUDA_TILE_BLOCKS", "NW_CUDA_ALIGN_TILE")


def knobs(**kw):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in kw.items()})


for (n1, n2) in ((700, 520), (3000, 1100), (1, 1), (33, 2100)):
    s1 = rng.integers(1, 5, size=n1, dtype=np.int8); s2 = rng.integers(1, 5, size=n2, dtype=np.int8)
    ref = None
    for env in ({}, {"NW_CUDA_LAG2": 0}, {"NW_CUDA_NO_PACKED": 1}, {"NW_CUDA_GENERIC": 1}, {"NW_CUDA_TILE_BLOCKS": 2}):
        knobs(**env)
        for R in (0, 2, 8):
            with nw.Plan(n1, n2, rows_per_lane=R) as p:
                p.upload(s1, s2); p.run(); row, col, sc = p.last_row(), p.last_col(), p.score()
                a1, a2 = p.traceback()
            with nw.Plan(n1, n2, mode=nw.NW_MODE_FULL, rows_per_lane=R) as p:
                p.upload(s1, s2); p.run(); t = p.table_to_host(); b1, b2 = p.traceback()
            assert t[-1, -1] == sc and np.array_equal(row, t[-1]) and np.array_equal(col, t[:, -1])
            assert np.array_equal(a1, b1) and np.array_equal(a2, b2)
            if ref is None:
                ref = t
            assert np.array_equal(ref, t), (env, R)
        t2 = nw.needlemanWunsch(s1, s2)                    # one-shot: streamed delivery where it applies
        assert np.array_equal(ref, t2)
    for env in ({"NW_CUDA_NO_STAIR": 1}, {"NW_CUDA_FORCE_STAIR": 1}, {}):
        knobs(**env)
        for R in (0, 2, 16):
            with nw.Plan(n1, n2, mode=nw.NW_MODE_SCORE, rows_per_lane=R) as p:
                p.upload(s1, s2); p.run()
                assert p.score() == ref[-1, -1], (env, R)
    knobs(NW_CUDA_ALIGN_TILE=300)
    a1, a2, sc = nw.align(s1, s2)
    assert sc == ref[-1, -1] and np.array_equal(a1[a1 != 0], s1) and np.array_equal(a2[a2 != 0], s2)
    knobs()
    for scoring in ((2, -1, -2), (5, -4, -3), (20, -7, -4)):
        t = nw.needlemanWunsch(s1, s2, scoring=scoring)
        assert nw.score(s1, s2, scoring=scoring) == t[-1, -1]
    for scoring in ((2, -1, -2, 1), (3, -3, -2, 1)):
        t = nw.needlemanWunsch(s1, s2, scoring=scoring)
        assert nw.best(s1, s2, scoring)[0] == t.max()
        with nw.Plan(n1, n2, mode=nw.NW_MODE_FULL, scoring=scoring) as p:
            p.upload(s1, s2); p.run(); p.traceback()
    print(f"{n1} x {n2}: ok, score {ref[-1, -1]}", flush=True)

for (n, l1, l2) in ((40, 300, 700), (9, 2500, 40), (64, 1000, 1000)):
    S1 = rng.integers(1, 5, size=(n, l1), dtype=np.int8); S2 = rng.integers(1, 5, size=(n, l2), dtype=np.int8)
    knobs()
    b = nw.batch_scores(S1, S2)
    knobs(NW_CUDA_NO_PACKED=1)
    assert np.array_equal(b, nw.batch_scores(S1, S2))
    knobs()
    nw.batch_scores(S1, S2, scoring=(2, -1, -2))
print("batches: ok")

s1 = rng.integers(1, 5, size=3003, dtype=np.int8); s2 = rng.integers(1, 5, size=1700, dtype=np.int8)
want = nw.score(s1, s2)
for mode in (nw.NW_MODE_BOUNDARY, nw.NW_MODE_FULL):
    for env in ({}, {"NW_CUDA_LAG2": 0}, {"NW_CUDA_NO_PACKED": 1}):
        knobs(**env)
        plans = [nw.Plan(s1.size, s2.size, mode=mode, part=k, nparts=3, rows_per_lane=4) for k in range(3)]
        for a, b in zip(plans, plans[1:]):
            a.connect(b)
        for p in plans:
            p.upload(s1, s2)
        for rep in range(2):
            for p in plans:
                p.run(); p.sync()
        assert plans[-1].score() == want
        if mode == nw.NW_MODE_BOUNDARY:
            a1, a2 = nw.plans_traceback(plans)
            assert np.array_equal(a1[a1 != 0], s1)
        for p in plans:
            p.close()
knobs()
print("column strips: ok")
print("ok")
