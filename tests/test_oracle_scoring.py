"""CPU: the oracle's general-scoring and local-alignment restatements (oracle/nw_oracle.c, *_ex).

Global alignment with other MATCH / MISMATCH / GAP values is PINNED: tests/golden/golden_scoring.json was produced by the
reference's own serial.cpp compiled against needleman-wunsch.hpp with only those three #defines edited
(tests/golden/make_golden_scoring.py).  Local alignment (Smith-Waterman) has no reference implementation: properties and one
textbook vector only ("parity unpinned")."""
import numpy as np
import pytest

from conftest import GOLDEN_SCORING, load_pair, synth_pair

TRIPLES = [(e["match"], e["mismatch"], e["gap"]) for e in GOLDEN_SCORING["triples"]]


@pytest.mark.parametrize("entry", GOLDEN_SCORING["triples"], ids=lambda e: f"{e['match']}_{e['mismatch']}_{e['gap']}")
def test_fill_ex_matches_the_reference_built_with_edited_macros(oracle, entry):
    sc = (entry["match"], entry["mismatch"], entry["gap"])
    for name, g in entry["synthetic"].items():
        s1, s2 = synth_pair(g["seed"], g["n1"], g["n2"], g["alphabet_hi"])
        t = oracle.fill_ex(s1, s2, sc)
        for k, v in oracle.table_facts(t).items():
            assert g[k] == v, (name, k)
        assert oracle.score_ex(s1, s2, sc) == (g["score"], s2.size, s1.size)
    for name, g in entry["fixtures"].items():
        s1, s2 = load_pair(name)
        for k, v in oracle.table_facts(oracle.fill_ex(s1, s2, sc)).items():
            assert g[k] == v, (name, k)


def test_default_triple_is_the_plain_oracle(oracle):
    s1, s2 = synth_pair(3, 400, 333, 5)
    assert np.array_equal(oracle.fill_ex(s1, s2, (1, 0, -1)), oracle.fill(s1, s2))
    a1, a2 = oracle.traceback_ex(s1, s2, (1, 0, -1))
    b1, b2 = oracle.traceback(s1, s2)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2)


@pytest.mark.parametrize("sc", [(2, -1, -2), (5, -4, -3), (1, -3, -1)])
def test_traceback_ex_reproduces_the_score(oracle, sc):
    s1, s2 = synth_pair(9, 300, 280, 5)
    a1, a2 = oracle.traceback_ex(s1, s2, sc)
    assert np.array_equal(a1[a1 != 0], s1) and np.array_equal(a2[a2 != 0], s2)
    m, x, g = sc
    col = np.where((a1 == 0) | (a2 == 0), g, np.where(a1 == a2, m, x))
    assert int(col.sum()) == oracle.score_ex(s1, s2, sc)[0]


def test_local_textbook_vector(oracle):
    # Smith-Waterman's standard worked example: TGTTACGG / GGTTGACTA, +3 / -3, linear gap -2 -> best local score 13
    code = {"A": 1, "T": 2, "G": 3, "C": 4}      # the reference's printer maps 1..4 to ATGC (src/common/helper.cpp:27-34)
    s1 = np.array([code[c] for c in "TGTTACGG"], dtype=np.int8)
    s2 = np.array([code[c] for c in "GGTTGACTA"], dtype=np.int8)
    sc, i, j = oracle.score_ex(s1, s2, (3, -3, -2, 1))
    t = oracle.fill_ex(s1, s2, (3, -3, -2, 1))
    assert sc == 13 == t.max() and t[i, j] == 13


@pytest.mark.parametrize("sc", [(1, 0, -1, 1), (2, -1, -2, 1), (3, -3, -2, 1), (5, -4, -3, 1), (1, -1, 0, 1)])
def test_local_properties(oracle, sc):
    rng = np.random.default_rng(17)
    for _ in range(6):
        n1, n2 = int(rng.integers(0, 300)), int(rng.integers(0, 300))
        s1, s2 = synth_pair(int(rng.integers(1 << 30)), n1, n2, int(rng.choice([2, 5, 60])))
        t = oracle.fill_ex(s1, s2, sc)
        best, i, j = oracle.score_ex(s1, s2, sc)
        assert t.min() >= 0 and (t[0] == 0).all() and (t[:, 0] == 0).all()
        assert best == t.max() and t[i, j] == best
        if best > 0:        # smallest column first, then smallest row
            jj = int(np.argmax((t == best).any(axis=0)))
            ii = int(np.argmax(t[:, jj] == best))
            assert (i, j) == (ii, jj)
        else:
            assert (i, j) == (0, 0)
        # a local alignment is at least as good as the global one, and as any clamped global alignment of prefixes
        assert best >= max(0, int(oracle.fill_ex(s1, s2, sc[:3]).max()))


@pytest.mark.parametrize("sc", [(2, -1, -2, 1), (3, -3, -2, 1), (1, -1, 0, 1)])
def test_local_traceback_properties(oracle, sc):
    rng = np.random.default_rng(23)
    m, x, g = sc[:3]
    for _ in range(5):
        s1, s2 = synth_pair(int(rng.integers(1 << 30)), int(rng.integers(1, 250)), int(rng.integers(1, 250)), 5)
        best, i, j = oracle.score_ex(s1, s2, sc)
        a1, a2 = oracle.traceback_local(s1, s2, sc)
        col = np.where((a1 == 0) | (a2 == 0), g, np.where(a1 == a2, m, x))
        assert int(col.sum()) == best                           # the aligned segments score the best cell's value
        n1, n2 = int((a1 != 0).sum()), int((a2 != 0).sum())     # ... and are substrings ending at the best cell
        assert np.array_equal(a1[a1 != 0], s1[j - n1:j]) and np.array_equal(a2[a2 != 0], s2[i - n2:i])
