"""GPU: alignment output without the table: the path is recovered tile by tile from the checkpoint rows
(strip boundary rows) and checkpoint columns (part halos) that a boundary-mode fill leaves in HBM.  The oracle walks the
full CPU table with the same rule (diagonal, then up, then left), so the alignments must be identical byte for byte."""
import numpy as np
import pytest

from conftest import GOLDEN, load_pair, synth_pair

pytestmark = pytest.mark.gpu


def column_score(a1, a2, sc=(1, 0, -1)):
    m, x, g = sc
    return int(np.where((a1 == 0) | (a2 == 0), g, np.where(a1 == a2, m, x)).sum())


@pytest.fixture(params=["lag2", "packed16", "int32"])
def kernel_kind(request, monkeypatch):
    if request.param == "int32":
        monkeypatch.setenv("NW_CUDA_NO_PACKED", "1")
    else:
        monkeypatch.setenv("NW_CUDA_LAG2", "1" if request.param == "lag2" else "0")
    return request.param


@pytest.mark.parametrize("name", ["small", "t", "debug", "smid"])
@pytest.mark.parametrize("tile", [64, 1000, 4096])
def test_align_fixtures(gpu, oracle, monkeypatch, name, tile):
    monkeypatch.setenv("NW_CUDA_ALIGN_TILE", str(tile))
    s1, s2 = load_pair(name)
    a1, a2, sc = gpu.align(s1, s2)
    b1, b2 = oracle.traceback(s1, s2)
    assert sc == GOLDEN["fixtures"][name]["score"]
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2)


@pytest.mark.parametrize("shape", [(0, 0), (0, 9), (9, 0), (1, 1), (1, 70), (70, 1), (63, 64), (64, 63), (65, 129), (700, 90),
                                   (90, 700), (1025, 1023), (5000, 3000), (2999, 6007)])
@pytest.mark.parametrize("tile", [64, 333, 4096])
def test_align_shapes(gpu, oracle, monkeypatch, shape, tile):
    monkeypatch.setenv("NW_CUDA_ALIGN_TILE", str(tile))
    n1, n2 = shape
    for hi in (5, 3):
        s1, s2 = synth_pair(900 + n1 + 3 * n2 + hi, n1, n2, hi)
        a1, a2, sc = gpu.align(s1, s2)
        b1, b2 = oracle.traceback(s1, s2)
        assert np.array_equal(a1, b1) and np.array_equal(a2, b2)
        assert sc == oracle.score(s1, s2) == column_score(a1, a2)


@pytest.mark.parametrize("R", [0, 1, 2, 4, 8, 16])
def test_boundary_plan_traceback(gpu, oracle, R, kernel_kind):
    if R == 16 and kernel_kind == "int32":
        pytest.skip("16 rows per lane exists only in the packed kernels")
    if R == 1 and kernel_kind != "int32":
        pytest.skip("1 row per lane exists only in the 32-bit kernels")
    rng = np.random.default_rng(41)
    s1 = rng.integers(1, 5, size=2600, dtype=np.int8)
    s2 = np.concatenate([s1[:900], rng.integers(1, 5, size=300, dtype=np.int8), s1[1500:2300]]).astype(np.int8)   # a long gap
    with gpu.Plan(s1.size, s2.size, rows_per_lane=R) as p:
        p.upload(s1, s2)
        p.run()
        a1, a2 = p.traceback()
        p.run()                                   # a second epoch
        c1, c2 = p.traceback()
    b1, b2 = oracle.traceback(s1, s2)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2) and np.array_equal(c1, b1) and np.array_equal(c2, b2)


@pytest.mark.parametrize("P", [2, 3, 8])
def test_pipeline_parts_traceback(gpu, oracle, P, kernel_kind):
    s1, s2 = synth_pair(53, 3003, 1700, 5)
    plans = [gpu.Plan(s1.size, s2.size, part=p, nparts=P, rows_per_lane=4) for p in range(P)]
    try:
        for a, b in zip(plans, plans[1:]):
            a.connect(b)
        for p in plans:
            p.upload(s1, s2)
        for rep in range(2):
            for p in plans:
                p.run()
                p.sync()
            a1, a2 = gpu.plans_traceback(plans)
            b1, b2 = oracle.traceback(s1, s2)
            assert np.array_equal(a1, b1) and np.array_equal(a2, b2)
        with pytest.raises(gpu.NwCudaError):
            gpu.plans_traceback(plans[1:])        # not the whole pipeline
    finally:
        for p in plans:
            p.close()


@pytest.mark.parametrize("sc", [(2, -1, -2), (5, -4, -3), (1, -3, -1), (3, 1, -1), (20, -7, -4)])
def test_align_with_scoring(gpu, oracle, monkeypatch, sc):
    monkeypatch.setenv("NW_CUDA_ALIGN_TILE", "500")
    s1, s2 = synth_pair(62, 1900, 2100, 5)
    a1, a2, score = gpu.align(s1, s2, scoring=sc)
    b1, b2 = oracle.traceback_ex(s1, s2, sc)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2)
    assert score == oracle.score_ex(s1, s2, sc)[0] == column_score(a1, a2, sc)


def test_align_generic_alphabet(gpu, oracle, monkeypatch):
    monkeypatch.setenv("NW_CUDA_ALIGN_TILE", "700")
    rng = np.random.default_rng(6)
    s1 = rng.integers(-128, 128, size=2100).astype(np.int8)
    s2 = np.concatenate([s1[200:1500], rng.integers(-128, 128, size=400).astype(np.int8)]).astype(np.int8)
    a1, a2, sc = gpu.align(s1, s2)
    b1, b2 = oracle.traceback(s1, s2)
    # (a byte value of 0 is also the gap code, like in the reference's encoding; compare the paths, not the letters)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2) and sc == oracle.score(s1, s2)


@pytest.mark.parametrize("name", ["2gb", "64gb"])
def test_align_large_pairs_properties(gpu, name):
    # tables of 2 GB and 64 GB that are never materialised: the alignment must spell both sequences and score the golden
    s1, s2 = load_pair(name)
    a1, a2, sc = gpu.align(s1, s2)
    assert sc == GOLDEN["fixtures"][name]["score"] == column_score(a1, a2)
    assert np.array_equal(a1[a1 != 0], s1) and np.array_equal(a2[a2 != 0], s2)
    assert not ((a1 == 0) & (a2 == 0)).any()
