"""CPU: pin the oracle (oracle/nw_oracle.c) to the reference -- golden vectors generated from the reference's own
serial.cpp (tests/golden/make_golden.py) and, where oracle/_ref was built, the reference library itself."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_pair, synth_pair

SMALL_TABLE = np.array([
    [0, -1, -2, -3, -4, -5, -6], [-1, 0, -1, -2, -3, -4, -5], [-2, 0, 0, 0, -1, -2, -3], [-3, -1, 0, 1, 1, 0, -1],
    [-4, -2, 0, 0, 1, 2, 1], [-5, -3, -1, 1, 1, 1, 2], [-6, -4, -2, 0, 2, 1, 1], [-7, -5, -3, -1, 1, 2, 1],
    [-8, -6, -4, -2, 0, 1, 2], [-9, -7, -5, -3, -1, 1, 1], [-10, -8, -6, -4, -2, 0, 2]], dtype=np.int32)


def test_small_table_literal(oracle):
    # the `small` pair written out (s1 = 1 3 1 1 3 4 across, s2 = 2 1 1 3 1 1 1 1 3 4 down)
    s1 = np.array([1, 3, 1, 1, 3, 4], dtype=np.int8)
    s2 = np.array([2, 1, 1, 3, 1, 1, 1, 1, 3, 4], dtype=np.int8)
    assert np.array_equal(oracle.fill(s1, s2), SMALL_TABLE)
    assert oracle.score(s1, s2) == 2


@pytest.mark.parametrize("name", ["small", "t", "debug", "smid", "2gb", "4gb", "mid"])
def test_fixture_scores(oracle, name):
    s1, s2 = load_pair(name)
    g = GOLDEN["fixtures"][name]
    assert (s1.size, s2.size) == (g["n1"], g["n2"])
    assert oracle.score(s1, s2) == g["score"]


@pytest.mark.parametrize("name", ["small", "t", "debug", "smid"])
def test_fixture_tables(oracle, name):
    s1, s2 = load_pair(name)
    t = oracle.fill(s1, s2)
    assert oracle.table_facts(t) == GOLDEN["tables"][name]
    row, col, sc, _ = oracle.boundaries(s1, s2)
    assert np.array_equal(row, t[-1]) and np.array_equal(col, t[:, -1]) and sc == t[-1, -1]


@pytest.mark.parametrize("name", sorted(GOLDEN["synthetic"]))
def test_synthetic_tables(oracle, name):
    g = GOLDEN["synthetic"][name]
    s1, s2 = synth_pair(g["seed"], g["n1"], g["n2"], g["alphabet_hi"])
    facts = oracle.table_facts(oracle.fill(s1, s2))
    for k, v in facts.items():
        assert g[k] == v, k


def test_against_reference_library(oracle):
    path = os.path.join(ROOT, "oracle", "_ref", "libnw_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libnw_ref.so not built (no reference tree)")
    ref = C.CDLL(path)
    ref.nw_ref_serial_fill.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    ref.nw_ref_serial_fill.restype = None
    rng = np.random.default_rng(7)
    for n1, n2 in [(1, 1), (5, 9), (64, 64), (257, 100), (100, 257), (1500, 1234)]:
        s1 = rng.integers(1, 5, size=n1, dtype=np.int8)
        s2 = rng.integers(1, 5, size=n2, dtype=np.int8)
        t = np.empty((n2 + 1, n1 + 1), dtype=np.int32)
        ref.nw_ref_serial_fill(s1.ctypes.data, n1, s2.ctypes.data, n2, t.ctypes.data)
        assert np.array_equal(oracle.fill(s1, s2), t)


@pytest.mark.parametrize("P", [2, 3, 4, 8])
def test_strip_chain_equals_whole(oracle, P):
    # the mpi-vert decomposition (mpi-vert-driver.cpp:35-36) chained left to right reproduces the serial table
    s1, s2 = synth_pair(21, 1003, 517, 5)
    t = oracle.fill(s1, s2)
    halo, covered = None, 0
    for p in range(P):
        start, ncols = oracle.strip_partition(s1.size, P, p)
        assert start == (covered - 1 if p else 0)
        right, last = oracle.strip(s1, s2, P, p, halo)
        assert np.array_equal(right, t[:, start + ncols - 1])
        halo, covered = right, start + ncols
    assert covered == s1.size + 1 and last == t[-1, -1]


def test_checkpoint_rows(oracle):
    s1, s2 = synth_pair(22, 300, 1000, 5)
    t = oracle.fill(s1, s2)
    _, _, _, rows = oracle.boundaries(s1, s2, row_stride=128)
    assert rows.shape[0] == 7
    for k in range(7):
        assert np.array_equal(rows[k], t[(k + 1) * 128])


def test_batch(oracle):
    rng = np.random.default_rng(20240607)
    S1 = rng.integers(1, 5, size=(50, 100), dtype=np.int8)
    S2 = rng.integers(1, 5, size=(50, 90), dtype=np.int8)
    sc = oracle.batch_scores(S1, S2)
    for p in (0, 17, 49):
        assert sc[p] == oracle.fill(S1[p], S2[p])[-1, -1]


def _check_alignment(s1, s2, a1, a2, score):
    assert a1.size == a2.size
    assert not ((a1 == 0) & (a2 == 0)).any()                       # never gap against gap
    assert np.array_equal(a1[a1 != 0], s1) and np.array_equal(a2[a2 != 0], s2)
    gaps = int((a1 == 0).sum() + (a2 == 0).sum())
    matches = int(((a1 == a2) & (a1 != 0)).sum())
    assert matches - gaps == score                                 # MATCH +1, MISMATCH 0, GAP -1


@pytest.mark.parametrize("shape", [(0, 0), (0, 4), (4, 0), (1, 1), (6, 10), (300, 280), (1000, 40)])
def test_traceback_is_an_optimal_alignment(oracle, shape):
    n1, n2 = shape
    s1, s2 = synth_pair(5 + n1 + n2, n1, n2, 5)
    a1, a2 = oracle.traceback(s1, s2)
    _check_alignment(s1, s2, a1, a2, oracle.score(s1, s2))


def test_traceback_small_fixture_literal(oracle):
    # the `small` pair: score 2
    s1 = np.array([1, 3, 1, 1, 3, 4], dtype=np.int8)
    s2 = np.array([2, 1, 1, 3, 1, 1, 1, 1, 3, 4], dtype=np.int8)
    a1, a2 = oracle.traceback(s1, s2)
    _check_alignment(s1, s2, a1, a2, 2)
    assert a1.size == 10
