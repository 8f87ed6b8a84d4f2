"""GPU: scoring parameters and Smith-Waterman through the C ABI.

Global alignment with other MATCH / MISMATCH / GAP values is checked against tests/golden/golden_scoring.json (the
reference's serial.cpp built with edited macros) and against the oracle; local alignment against the oracle
(parity unpinned: the reference has no Smith-Waterman).  Bit-exact everywhere."""
import numpy as np
import pytest

from conftest import GOLDEN_SCORING, load_pair, synth_pair

pytestmark = pytest.mark.gpu
TRIPLES = [(e["match"], e["mismatch"], e["gap"]) for e in GOLDEN_SCORING["triples"]]


@pytest.fixture(params=["auto", "lag1", "int32", "generic"])
def kernel_kind(request, monkeypatch):
    if request.param == "lag1":
        monkeypatch.setenv("NW_CUDA_LAG2", "0")
    elif request.param == "int32":
        monkeypatch.setenv("NW_CUDA_NO_PACKED", "1")
    elif request.param == "generic":
        monkeypatch.setenv("NW_CUDA_GENERIC", "1")
    return request.param


@pytest.mark.parametrize("entry", GOLDEN_SCORING["triples"], ids=lambda e: f"{e['match']}_{e['mismatch']}_{e['gap']}")
def test_full_table_against_reference_vectors(gpu, oracle, entry, kernel_kind):
    sc = (entry["match"], entry["mismatch"], entry["gap"])
    for name, g in entry["synthetic"].items():
        s1, s2 = synth_pair(g["seed"], g["n1"], g["n2"], g["alphabet_hi"])
        t = gpu.needlemanWunsch(s1, s2, scoring=sc)
        for k, v in oracle.table_facts(t).items():
            assert g[k] == v, (name, k)
        assert gpu.score(s1, s2, scoring=sc) == g["score"]                      # score mode (meet in the middle)
        assert gpu.best(s1, s2, sc) == (g["score"], s2.size, s1.size)
    for name, g in entry["fixtures"].items():
        s1, s2 = load_pair(name)
        for k, v in oracle.table_facts(gpu.needlemanWunsch(s1, s2, scoring=sc)).items():
            assert g[k] == v, (name, k)


@pytest.mark.parametrize("sc", TRIPLES + [(7, 2, -6), (0, 0, -1), (1, 1, -1), (300, -200, -150), (2, -1, 0)])
@pytest.mark.parametrize("R", [0, 2, 8])
def test_boundaries_and_checkpoints(gpu, oracle, sc, R, kernel_kind):
    s1, s2 = synth_pair(77, 2300, 1900, 5)
    t = oracle.fill_ex(s1, s2, sc)
    with gpu.Plan(s1.size, s2.size, rows_per_lane=R, scoring=sc) as p:
        p.upload(s1, s2)
        p.run()
        assert p.score() == t[-1, -1]
        assert np.array_equal(p.last_row(), t[-1]) and np.array_equal(p.last_col(), t[:, -1])
        info = p.strip_info()
        for k in range(info["nstrips"]):
            i = s2.size - (info["nstrips"] - 1 - k) * info["strip_rows"]
            assert np.array_equal(p.strip_row(k), t[i]), k
    with gpu.Plan(s1.size, s2.size, mode=gpu.NW_MODE_SCORE, scoring=sc) as p:
        p.upload(s1, s2)
        p.run()
        assert p.score() == t[-1, -1]


def test_long_rows_rebase_with_large_weights(gpu, oracle):
    # w = match - 2 gap = 16: the largest weight the packed kernels take; 40 000 columns cross many re-basing points
    sc = (6, -2, -5)
    rng = np.random.default_rng(5)
    s1 = rng.integers(1, 5, size=40000, dtype=np.int8)
    s2 = np.concatenate([s1[100:700], rng.integers(1, 5, size=200, dtype=np.int8)]).astype(np.int8)
    want = oracle.score_ex(s1, s2, sc)[0]
    for mode in (gpu.NW_MODE_BOUNDARY, gpu.NW_MODE_SCORE):
        with gpu.Plan(s1.size, s2.size, mode=mode, scoring=sc) as p:
            p.upload(s1, s2)
            p.run()
            assert p.score() == want


@pytest.mark.parametrize("P", [2, 3])
@pytest.mark.parametrize("mode", ["boundary", "full"])
@pytest.mark.parametrize("sc", [(2, -1, -2), (3, 1, -1)])
def test_column_strips_with_scoring(gpu, oracle, P, mode, sc, kernel_kind):
    s1, s2 = synth_pair(52, 2503, 1300, 5)
    t = oracle.fill_ex(s1, s2, sc)
    m = gpu.NW_MODE_FULL if mode == "full" else gpu.NW_MODE_BOUNDARY
    plans = [gpu.Plan(s1.size, s2.size, mode=m, part=p, nparts=P, rows_per_lane=4, scoring=sc) for p in range(P)]
    try:
        for a, b in zip(plans, plans[1:]):
            a.connect(b)
        for p in plans:
            p.upload(s1, s2)
        for rep in range(2):
            for p in plans:
                p.run()
                p.sync()
        out = np.zeros_like(t)
        for p in plans:
            assert np.array_equal(p.last_col(), t[:, p.jstart + p.ncols])
            assert np.array_equal(p.last_row(), t[-1, p.jstart:p.jstart + p.ncols + 1])
            if mode == "full":
                p.table_to_host(out)
        assert plans[-1].score() == t[-1, -1]
        if mode == "full":
            assert np.array_equal(out, t)
    finally:
        for p in plans:
            p.close()


@pytest.mark.parametrize("sc", [(2, -1, -2), (5, -4, -3), (1, -3, -1), (3, 1, -1)])
def test_traceback_with_scoring(gpu, oracle, sc):
    s1, s2 = synth_pair(61, 900, 1100, 5)
    with gpu.Plan(s1.size, s2.size, mode=gpu.NW_MODE_FULL, scoring=sc) as p:
        p.upload(s1, s2)
        p.run()
        a1, a2 = p.traceback()
    b1, b2 = oracle.traceback_ex(s1, s2, sc)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2)


@pytest.mark.parametrize("sc", [(2, -1, -2), (5, -4, -3), (20, -7, -4), (300, -200, -150)])
@pytest.mark.parametrize("shape", [(200, 100, 90), (40, 1000, 1000), (5, 1500, 2100), (3, 0, 5)])
def test_batch_with_scoring(gpu, oracle, sc, shape):
    n, l1, l2 = shape
    rng = np.random.default_rng(l1 + l2)
    S1 = rng.integers(1, 5, size=(n, l1), dtype=np.int8)
    S2 = rng.integers(1, 5, size=(n, l2), dtype=np.int8)
    want = np.array([oracle.score_ex(S1[k], S2[k], sc)[0] for k in range(n)], dtype=np.int32)
    assert np.array_equal(gpu.batch_scores(S1, S2, scoring=sc), want)


# ---- Smith-Waterman ---------------------------------------------------------------------------------------------------------
def test_local_textbook_vector(gpu):
    code = {"A": 1, "T": 2, "G": 3, "C": 4}
    s1 = np.array([code[c] for c in "TGTTACGG"], dtype=np.int8)
    s2 = np.array([code[c] for c in "GGTTGACTA"], dtype=np.int8)
    assert gpu.best(s1, s2, (3, -3, -2, 1))[0] == 13


@pytest.mark.parametrize("sc", [(1, 0, -1, 1), (2, -1, -2, 1), (3, -3, -2, 1), (5, -4, -3, 1), (1, -1, 0, 1), (2, 1, -1, 1)])
@pytest.mark.parametrize("shape", [(0, 0), (0, 7), (7, 0), (1, 1), (31, 33), (33, 31), (64, 64), (300, 1000), (1000, 300),
                                   (2500, 2100), (4097, 129), (129, 4097)])
def test_local_best_and_table(gpu, oracle, sc, shape):
    n1, n2 = shape
    for hi in (5, 90) if n1 * n2 < 10000 else (5,):
        s1, s2 = synth_pair(200 + n1 + 3 * n2, n1, n2, hi)
        want = oracle.score_ex(s1, s2, sc)
        assert gpu.best(s1, s2, sc) == want
        t = gpu.needlemanWunsch(s1, s2, scoring=sc)
        assert np.array_equal(t, oracle.fill_ex(s1, s2, sc))


@pytest.mark.parametrize("R", [1, 2, 4, 8])
def test_local_plan_rows_per_lane_and_checkpoints(gpu, oracle, R):
    sc = (2, -1, -2, 1)
    rng = np.random.default_rng(3)
    s1 = rng.integers(1, 5, size=3000, dtype=np.int8)
    s2 = np.concatenate([rng.integers(1, 5, size=900, dtype=np.int8), s1[1000:1400], rng.integers(1, 5, size=500, dtype=np.int8)]).astype(np.int8)
    t = oracle.fill_ex(s1, s2, sc)
    with gpu.Plan(s1.size, s2.size, rows_per_lane=R, scoring=sc) as p:
        p.upload(s1, s2)
        for _ in range(2):
            p.run()
            assert p.best() == oracle.score_ex(s1, s2, sc)
        assert p.score() == t.max() >= 2 * 400
        info = p.strip_info()
        for k in range(info["nstrips"]):
            i = s2.size - (info["nstrips"] - 1 - k) * info["strip_rows"]
            assert np.array_equal(p.strip_row(k), t[i]), k
        with pytest.raises(gpu.NwCudaError):
            p.last_row()


@pytest.mark.parametrize("sc", [(2, -1, -2, 1), (3, -3, -2, 1), (5, -4, -3, 1), (1, -1, 0, 1)])
@pytest.mark.parametrize("shape", [(0, 5), (1, 1), (33, 31), (300, 1000), (2500, 2100), (129, 4097)])
def test_local_traceback(gpu, oracle, sc, shape):
    n1, n2 = shape
    s1, s2 = synth_pair(700 + n1 + 3 * n2, n1, n2, 5)
    with gpu.Plan(n1, n2, mode=gpu.NW_MODE_FULL, scoring=sc) as p:
        p.upload(s1, s2)
        p.run()
        a1, a2 = p.traceback()
        best, i, j = p.best()
    b1, b2 = oracle.traceback_local(s1, s2, sc)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2)
    assert np.array_equal(a1[a1 != 0], s1[j - int((a1 != 0).sum()):j]) and np.array_equal(a2[a2 != 0], s2[i - int((a2 != 0).sum()):i])


def test_local_tie_rule(gpu, oracle):
    # many equal maxima: identical short motifs scattered over both sequences
    rng = np.random.default_rng(8)
    motif = np.array([1, 2, 3, 4, 1, 1, 2], dtype=np.int8)
    s1 = np.concatenate([np.concatenate([motif, rng.integers(5, 9, size=11, dtype=np.int8)]) for _ in range(30)]).astype(np.int8)
    s2 = np.concatenate([np.concatenate([rng.integers(9, 13, size=7, dtype=np.int8), motif]) for _ in range(40)]).astype(np.int8)
    sc = (1, -5, -5, 1)
    want = oracle.score_ex(s1, s2, sc)
    assert want[0] == motif.size
    assert gpu.best(s1, s2, sc) == want


def test_scoring_errors(gpu):
    s = np.ones(10, np.int8)
    with pytest.raises(gpu.NwCudaError):
        gpu.best(s, s, (1, 0, 1, 1))                 # local alignment with a positive gap
    with pytest.raises(gpu.NwCudaError):
        gpu.best(s, s, (1, 0, -1, 2))                # unknown `local`
    with pytest.raises(gpu.NwCudaError):
        gpu.score(np.ones(40000, np.int8), np.ones(40000, np.int8), scoring=(100000, 0, -1))     # overflows int32
    bad = gpu.Scoring(1, 0, -1, 0)
    bad.reserved[0] = 1
    with pytest.raises(gpu.NwCudaError):
        gpu.score(s, s, scoring=bad)
    with pytest.raises(gpu.NwCudaError):
        gpu.Plan(100, 100, part=0, nparts=2, scoring=(1, 0, -1, 1))      # local alignment is single-device
