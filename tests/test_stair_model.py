"""CPU: the cut behind staircase score mode, checked on full oracle tables.

NW_MODE_SCORE fills one part of the table forwards and the rest backwards and takes the maximum of F + B over a vertex cut
that every monotone path crosses.  The product uses the row n2/2 or, when there are schedulers for twice the strips, a
4-connected anti-diagonal staircase aligned to the strip grid:
    { (r_s+1, j) : x_s+1 <= j <= x_s }  and  { (i, x_s) : r_s < i < r_s+1 },   s = 0 .. S-1,  x_0 = n1,  x_s = n1 (S - s) / S,  x_S = 0
with strip s = table rows (r_s, r_s+1], r_s = max(s * SH - pad, 0), pad = S * SH - n2 (padding rows on top, as the kernels
place them).  This test restates exactly what nw_stair_combine_kernel evaluates, with F and B taken from the oracle's full
tables instead of the GPU's boundary rows and columns, and compares with the oracle's score -- for the reference scoring and
for other triples."""
import numpy as np
import pytest

from conftest import synth_pair


def stair_score(oracle, s1, s2, SH, scoring=(1, 0, -1)):
    n1, n2 = s1.size, s2.size
    F = oracle.fill_ex(s1, s2, scoring)                           # F[i][j]: best score (0,0) -> (i,j)
    Brev = oracle.fill_ex(s1[::-1].copy(), s2[::-1].copy(), scoring)
    B = Brev[::-1, ::-1]                                          # B[i][j]: best score (i,j) -> (n2,n1)
    S = max(1, -(-n2 // SH))
    pad = S * SH - n2
    x = [n1 if s == 0 else n1 * (S - s) // S for s in range(S)] + [0]
    best = F[0, n1] + B[0, n1]                                    # the corner vertex (all gaps first)
    for s in range(S):
        r_lo, r_hi = max(s * SH - pad, 0), (s + 1) * SH - pad
        for j in range(x[s + 1], x[s] + 1):
            best = max(best, F[r_hi, j] + B[r_hi, j])
        for i in range(r_lo + 1, r_hi):
            best = max(best, F[i, x[s]] + B[i, x[s]])
    return int(best), int(F[n2, n1])


@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (3, 7), (40, 33), (33, 40), (130, 70), (70, 130), (257, 64), (64, 257)])
@pytest.mark.parametrize("SH", [1, 4, 32, 64])
def test_staircase_cut_gives_the_score(oracle, shape, SH):
    n1, n2 = shape
    for seed in range(3):
        for hi in (3, 5):
            s1, s2 = synth_pair(1000 * seed + n1 + 3 * n2 + hi, n1, n2, hi)
            got, want = stair_score(oracle, s1, s2, SH)
            assert got == want == oracle.score(s1, s2)


@pytest.mark.parametrize("scoring", [(2, -1, -2), (5, -4, -3), (1, -3, -1), (3, 1, -1)])
def test_staircase_cut_with_other_scoring(oracle, scoring):
    s1, s2 = synth_pair(77, 150, 90, 5)
    got, want = stair_score(oracle, s1, s2, 16, scoring)
    assert got == want == oracle.score_ex(s1, s2, scoring)[0]


def test_a_cut_with_a_hole_would_not(oracle):
    # sanity of the test itself: dropping the vertical segments leaves an 8-connected cut that diagonal moves slip through
    rng = np.random.default_rng(5)
    missed = 0
    for _ in range(40):
        n = int(rng.integers(20, 60))
        s1 = rng.integers(1, 3, size=n, dtype=np.int8)
        s2 = s1.copy()                                            # the optimal path is the main diagonal
        F = oracle.fill(s1, s2)
        B = oracle.fill(s1[::-1].copy(), s2[::-1].copy())[::-1, ::-1]
        SH, S = 8, -(-n // 8)
        pad = S * SH - n
        x = [n if s == 0 else n * (S - s) // S for s in range(S)] + [0]
        best = -10 ** 9
        for s in range(S):
            r_hi = (s + 1) * SH - pad
            for j in range(x[s + 1] + 1, x[s]):                   # horizontal segments WITHOUT their end points
                best = max(best, F[r_hi, j] + B[r_hi, j])
        missed += best < F[n, n]
    assert missed > 0
