"""GPU: the CUDA path (through the C ABI of include/nw_cuda.h) against the oracle and the golden vectors.
Bit-exact: everything here is int32 / byte work."""
import numpy as np
import pytest

from conftest import GOLDEN, load_pair, synth_pair

pytestmark = pytest.mark.gpu


def facts_match(oracle, t, g):
    f = oracle.table_facts(t)
    for k, v in f.items():
        assert g[k] == v, k


# ---- full-table mode ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["small", "t", "debug", "smid"])
def test_fixture_full_table(gpu, oracle, name):
    s1, s2 = load_pair(name)
    t = gpu.needlemanWunsch(s1, s2)
    facts_match(oracle, t, GOLDEN["tables"][name])
    if t.size < 200_000_000:
        assert np.array_equal(t, oracle.fill(s1, s2))


@pytest.mark.parametrize("name", sorted(GOLDEN["synthetic"]))
@pytest.mark.parametrize("R", [0, 1, 2, 4, 8, 16])
def test_synthetic_full_table(gpu, oracle, name, R):
    g = GOLDEN["synthetic"][name]
    s1, s2 = synth_pair(g["seed"], g["n1"], g["n2"], g["alphabet_hi"])
    with gpu.Plan(s1.size, s2.size, mode=gpu.NW_MODE_FULL, rows_per_lane=R) as p:
        p.upload(s1, s2)
        p.run()
        t = p.table_to_host()
        facts_match(oracle, t, g)
        assert np.array_equal(t, oracle.fill(s1, s2))
        assert p.score() == g["score"]
        assert np.array_equal(p.last_row(), t[-1]) and np.array_equal(p.last_col(), t[:, -1])


@pytest.mark.parametrize("tile_blocks", [2, 3, 32])
@pytest.mark.parametrize("R", [2, 8])
def test_full_table_tiles_and_snapshots(gpu, oracle, monkeypatch, tile_blocks, R):
    # packed full-table mode is two passes: pass 1 snapshots the warp state every `tile_blocks` blocks, pass 2 replays
    # every tile independently.  Small tiles on a wide table exercise many snapshots, re-basing across tiles, the ragged
    # last tile and the edge (predicated) blocks.
    monkeypatch.setenv("NW_CUDA_TILE_BLOCKS", str(tile_blocks))
    rng = np.random.default_rng(12)
    s1 = rng.integers(1, 5, size=9000, dtype=np.int8)
    s2 = np.concatenate([s1[:700], rng.integers(1, 5, size=333, dtype=np.int8)]).astype(np.int8)
    with gpu.Plan(s1.size, s2.size, mode=gpu.NW_MODE_FULL, rows_per_lane=R) as p:
        p.upload(s1, s2)
        p.run()
        assert np.array_equal(p.table_to_host(), oracle.fill(s1, s2))
        p.run()                                   # a second epoch over the same buffers
        assert np.array_equal(p.table_to_host(), oracle.fill(s1, s2))


def test_2gb_full_table_golden(gpu, oracle):
    # BASELINE.json configs[2], full-table mode: 22 117 x 22 542 int32 = 1.99 GB, checked by sum/min/max/FNV
    s1, s2 = load_pair("2gb")
    t = gpu.needlemanWunsch(s1, s2)
    facts_match(oracle, t, GOLDEN["tables"]["2gb"])


# ---- boundary-only mode ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(GOLDEN["fixtures"]))      # all 38 bdna pairs of the reference
def test_fixture_scores(gpu, name):
    s1, s2 = load_pair(name)
    assert gpu.score(s1, s2) == GOLDEN["fixtures"][name]["score"]


def test_boundary_mode_writes_only_the_score(gpu):
    s1, s2 = load_pair("debug")
    t = np.full((s2.size + 1, s1.size + 1), 12345, dtype=np.int32)
    gpu.needlemanWunsch(s1, s2, t, mode=gpu.NW_MODE_BOUNDARY)
    assert t[-1, -1] == GOLDEN["fixtures"]["debug"]["score"]
    t[-1, -1] = 12345
    assert (t == 12345).all()


@pytest.mark.parametrize("name", ["smid", "2gb"])
def test_boundaries_match_golden_hashes(gpu, oracle, name, kernel_kind):
    s1, s2 = load_pair(name)
    row, col, sc = gpu.boundaries(s1, s2)
    g = GOLDEN["tables"][name]
    assert oracle.fnv(row) == g["fnv_lastrow"] and oracle.fnv(col) == g["fnv_lastcol"] and sc == g["score"]


@pytest.fixture(params=["lag2", "packed16", "int32"])
def kernel_kind(request, monkeypatch):
    """Boundary mode has three kernels: packed s16x2 with virtual lanes two columns apart (nw_lag2.cuh, the default), one
    column apart (nw_packed.cuh; also pass 1 of full-table mode) and 32-bit (nw_kernels.cuh; generic alphabets)."""
    if request.param == "int32":
        monkeypatch.setenv("NW_CUDA_NO_PACKED", "1")
    else:
        monkeypatch.setenv("NW_CUDA_LAG2", "1" if request.param == "lag2" else "0")
    return request.param


@pytest.mark.parametrize("R", [1, 2, 4, 8, 16])
def test_checkpoint_rows(gpu, oracle, R, kernel_kind):
    if R == 16 and kernel_kind == "int32":
        pytest.skip("16 rows per lane exists only in the packed kernel")
    s1, s2 = synth_pair(31, 1500, 2000, 5)
    with gpu.Plan(s1.size, s2.size, rows_per_lane=R) as p:
        p.upload(s1, s2)
        p.run()
        info = p.strip_info()
        assert info["rows_per_lane"] == R and info["strip_rows"] == 32 * R
        t = oracle.fill(s1, s2)
        for k in range(info["nstrips"]):
            i = s2.size - (info["nstrips"] - 1 - k) * info["strip_rows"]
            assert np.array_equal(p.strip_row(k), t[i]), (k, i)


@pytest.mark.parametrize("shape", [(0, 0), (0, 5), (5, 0), (1, 1), (1, 40), (40, 1), (31, 33), (33, 31), (64, 64),
                                   (65, 255), (255, 65), (2, 3000), (3000, 2), (777, 1025)])
def test_edge_shapes(gpu, oracle, shape, kernel_kind):
    n1, n2 = shape
    s1, s2 = synth_pair(100 + n1 + 7 * n2, n1, n2, 5)
    t = oracle.fill(s1, s2)
    assert np.array_equal(gpu.needlemanWunsch(s1, s2), t)
    row, col, sc = gpu.boundaries(s1, s2)
    assert np.array_equal(row, t[-1]) and np.array_equal(col, t[:, -1]) and sc == t[-1, -1]


def test_many_random_shapes(gpu, oracle, kernel_kind):
    rng = np.random.default_rng(5)
    for _ in range(40):
        n1, n2 = int(rng.integers(1, 700)), int(rng.integers(1, 700))
        hi = int(rng.choice([2, 3, 5, 9, 120]))
        s1 = rng.integers(1, hi, size=n1, dtype=np.int8)
        s2 = rng.integers(1, hi, size=n2, dtype=np.int8)
        t = oracle.fill(s1, s2)
        assert np.array_equal(gpu.needlemanWunsch(s1, s2), t), (n1, n2, hi)
        row, col, sc = gpu.boundaries(s1, s2)
        assert np.array_equal(row, t[-1]) and np.array_equal(col, t[:, -1]) and sc == t[-1, -1], (n1, n2, hi)


@pytest.mark.parametrize("kind", ["lag2", "lag1"])
@pytest.mark.parametrize("R", [0, 2, 4, 8, 16])
def test_boundaries_long_rows_rebase(gpu, oracle, R, kind, monkeypatch):
    monkeypatch.setenv("NW_CUDA_LAG2", "1" if kind == "lag2" else "0")
    # wide tables make the packed kernel re-base its 16-bit lanes many times; identical prefixes make G grow fastest
    rng = np.random.default_rng(9)
    s1 = rng.integers(1, 5, size=40000, dtype=np.int8)
    s2 = np.concatenate([s1[:1500], rng.integers(1, 5, size=700, dtype=np.int8)]).astype(np.int8)
    with gpu.Plan(s1.size, s2.size, rows_per_lane=R) as p:
        p.upload(s1, s2)
        p.run()
        row, col, _, _ = oracle.boundaries(s1, s2)
        assert np.array_equal(p.last_row(), row) and np.array_equal(p.last_col(), col)
    # tall and narrow: many strips, few columns
    with gpu.Plan(s2.size, s1.size, rows_per_lane=R) as p:
        p.upload(s2, s1)
        p.run()
        row, col, _, _ = oracle.boundaries(s2, s1)
        assert np.array_equal(p.last_row(), row) and np.array_equal(p.last_col(), col)


def test_generic_alphabet_and_negative_bytes(gpu, oracle):
    # any byte values must compare as bytes (serial.cpp:23); more than four distinct values takes the generic path
    rng = np.random.default_rng(6)
    s1 = rng.integers(-128, 128, size=900, dtype=np.int8)
    s2 = rng.integers(-128, 128, size=1100, dtype=np.int8)
    s2[:400] = s1[:400]
    assert np.array_equal(gpu.needlemanWunsch(s1, s2), oracle.fill(s1, s2))
    a = np.array([0, -1, 7, 0, 0, -1, 7, 7] * 50, dtype=np.int8)      # four-letter path with unusual codes
    b = np.array([7, 0, -1, -1, 0, 7] * 70, dtype=np.int8)
    assert np.array_equal(gpu.needlemanWunsch(a, b), oracle.fill(a, b))


def test_repeated_runs_and_reupload(gpu, oracle):
    s1, s2 = synth_pair(41, 2000, 1800, 5)
    u1, u2 = synth_pair(42, 2000, 1800, 5)
    with gpu.Plan(2000, 1800) as p:
        p.upload(s1, s2)
        for _ in range(3):
            p.run()
        assert p.score() == oracle.score(s1, s2)
        p.upload(u1, u2)
        p.run()
        assert p.score() == oracle.score(u1, u2)
        assert p.time(3) > 0


# ---- column strips (mpi-vert decomposition) on ONE device, parts run one after the other -------------------------------
@pytest.mark.parametrize("P", [2, 3, 8])
@pytest.mark.parametrize("mode", ["boundary", "full"])
def test_column_strips_sequential(gpu, oracle, P, mode, kernel_kind):
    s1, s2 = synth_pair(51, 3003, 1700, 5)
    t = oracle.fill(s1, s2)
    m = gpu.NW_MODE_FULL if mode == "full" else gpu.NW_MODE_BOUNDARY
    plans = [gpu.Plan(s1.size, s2.size, mode=m, part=p, nparts=P, rows_per_lane=4) for p in range(P)]
    try:
        for a, b in zip(plans, plans[1:]):
            a.connect(b)
        for p in plans:
            p.upload(s1, s2)
        for rep in range(3):                       # exercises the double-buffered mailboxes and the ack words
            for p in plans:
                p.run()
                p.sync()
        out = np.zeros_like(t)
        for p in plans:
            assert p.jstart == oracle.strip_partition(s1.size, P, p.part)[0]
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


# ---- batch of independent pairs ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(300, 100, 90), (64, 1000, 1000), (10, 1500, 2100), (7, 0, 5), (7, 5, 0), (0, 10, 10),
                                   (33, 129, 31)])
def test_batch_scores(gpu, oracle, shape):
    npairs, len1, len2 = shape
    rng = np.random.default_rng(20240607)
    S1 = rng.integers(1, 5, size=(npairs, len1), dtype=np.int8)
    S2 = rng.integers(1, 5, size=(npairs, len2), dtype=np.int8)
    assert np.array_equal(gpu.batch_scores(S1, S2), oracle.batch_scores(S1, S2))


@pytest.mark.parametrize("shape", [(5000, 100, 90), (4100, 1000, 1000), (3, 50, 60), (4097, 31, 700)])
@pytest.mark.parametrize("nchunks", [0, 1, 3, 8])
def test_batch_run_host_chunked(gpu, oracle, shape, nchunks):
    # nw_batch_run_host: H2D of chunk c+1 overlaps the kernel of chunk c; same scores as the plain path and the oracle
    n, l1, l2 = shape
    rng = np.random.default_rng(n + l1)
    S1 = rng.integers(1, 5, size=(n, l1), dtype=np.int8)
    S2 = rng.integers(1, 5, size=(n, l2), dtype=np.int8)
    b = gpu.Batch(n, l1, l2)
    try:
        got = b.run_host(S1, S2, nchunks=nchunks)
        again = b.run_host(S1, S2, nchunks=nchunks)
    finally:
        b.close()
    want = gpu.batch_scores(S1[:64], S2[:64])
    assert np.array_equal(got, again) and np.array_equal(got[:64], want)
    idx = rng.choice(n, size=min(n, 40), replace=False)
    assert np.array_equal(got[idx], oracle.batch_scores(S1[idx], S2[idx]))


def test_batch_run_host_late_letter_falls_back(gpu, oracle):
    # a fifth letter that appears only in the last chunk: the kernel chosen from chunk 0 does not fit, the batch is redone
    rng = np.random.default_rng(77)
    n, L = 4200, 120
    S1 = rng.integers(1, 5, size=(n, L), dtype=np.int8)
    S2 = rng.integers(1, 5, size=(n, L), dtype=np.int8)
    S1[-3:, ::7] = 9
    S2[-2:, ::5] = 9
    b = gpu.Batch(n, L, L)
    try:
        got = b.run_host(S1, S2, nchunks=6)
    finally:
        b.close()
    idx = np.r_[0:20, n - 20:n]
    assert np.array_equal(got[idx], oracle.batch_scores(S1[idx], S2[idx]))
    assert np.array_equal(gpu.batch_scores(S1, S2), got)          # the one-shot call takes the same path


def test_batch_generic_alphabet(gpu, oracle):
    rng = np.random.default_rng(8)
    S1 = rng.integers(-100, 100, size=(40, 300), dtype=np.int8)
    S2 = S1.copy()
    S2[:, ::7] = 5
    assert np.array_equal(gpu.batch_scores(S1, S2[:, :280].copy()), oracle.batch_scores(S1, S2[:, :280].copy()))


def test_batch_headline_shape_sample(gpu, oracle):
    # BASELINE.json configs[4] generator, first 20 000 pairs on the GPU, 600 of them checked on the CPU
    rng = np.random.default_rng(20240607)
    N = 20000
    S1 = rng.integers(1, 5, size=(N, 1000), dtype=np.int8)
    S2 = rng.integers(1, 5, size=(N, 1000), dtype=np.int8)
    got = gpu.batch_scores(S1, S2)
    idx = np.concatenate([np.arange(200), np.arange(N - 200, N), rng.choice(N, 200, replace=False)])
    assert np.array_equal(got[idx], oracle.batch_scores(np.ascontiguousarray(S1[idx]), np.ascontiguousarray(S2[idx])))
    # property at full size: identical sequences score len, and the score is symmetric in (s1, s2)
    assert (gpu.batch_scores(S1[:500], S1[:500]) == 1000).all()
    assert np.array_equal(gpu.batch_scores(S2[:500], S1[:500]), got[:500])


# ---- size-independent properties at full BASELINE sizes ------------------------------------------------------------------
def test_properties_mid(gpu):
    s1, s2 = load_pair("mid")
    sc = gpu.score(s1, s2)
    assert sc == GOLDEN["fixtures"]["mid"]["score"]
    assert gpu.score(s2, s1) == sc                      # transposition symmetry of the recurrence
    assert gpu.score(s1, s1) == s1.size                 # identity: every diagonal step is a match
    assert gpu.score(s1[::-1].copy(), s2[::-1].copy()) == sc   # reversal symmetry


def test_driver_binary_prints_reference_format(gpu):
    # the reference's UNCHANGED driver.cpp + helper.cpp around our entry point (built where the reference tree exists)
    import os
    import re
    import subprocess
    from conftest import ROOT, pair_paths
    exe = os.path.join(ROOT, "fast-needleman-wunsch_b200", "bin", "cuda.e")
    if not os.path.exists(exe):
        pytest.skip("cuda.e not built (no reference tree at build time)")
    a, b = pair_paths("smid")
    for mode in ("full", "boundary"):
        out = subprocess.run([exe, a, b], capture_output=True, text=True, env=dict(os.environ, NW_CUDA_MODE=mode))
        assert out.returncode == 0, out.stderr
        assert re.fullmatch(r"\d+\nScore: 5839\n", out.stdout), out.stdout
    out = subprocess.run([exe, a], capture_output=True, text=True)
    assert out.returncode == 1 and "incorrect number of arguments" in out.stdout
    out = subprocess.run([exe, a, "/nonexistent.bdna"], capture_output=True, text=True)
    assert out.returncode == 1 and "ERROR: no such file /nonexistent.bdna" in out.stdout


# ---- more than one GPU (skipped on a 1-GPU box; run with gpurun --gpus 2) -------------------------------------------------
def _need_gpus(gpu, n):
    if gpu.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("ngpus", [2, 4])
def test_multi_gpu_in_process_pipeline(gpu, oracle, ngpus):
    # nw_cuda_fill_ex with ngpus > 1: column strips on devices 0..n-1 of ONE process, boundary column stored straight
    # into the right neighbour's mailbox over NVLink, all kernels in flight at once
    _need_gpus(gpu, ngpus)
    s1, s2 = synth_pair(61, 20011, 9000, 5)
    t = oracle.fill(s1, s2)
    got = gpu.needlemanWunsch(s1, s2, mode=gpu.NW_MODE_FULL, ngpus=ngpus)
    assert np.array_equal(got, t)
    tb = np.zeros_like(t)
    gpu.needlemanWunsch(s1, s2, tb, mode=gpu.NW_MODE_BOUNDARY, ngpus=ngpus)
    assert tb[-1, -1] == t[-1, -1]


def test_multi_gpu_64gb_score(gpu):
    _need_gpus(gpu, 2)
    s1, s2 = load_pair("64gb")
    n = gpu.device_count()
    for g in sorted({2, min(n, 4), min(n, 8)}):
        t = np.zeros(1, dtype=np.int32)   # boundary mode writes only the last cell: hand it a 1-cell view trick is not
        # allowed (the ABI indexes (n1+1)*(n2+1)-1), so use the plan API instead
        plans = [gpu.Plan(s1.size, s2.size, device=d, part=d, nparts=g, rows_per_lane=8) for d in range(g)]
        try:
            for a, b in zip(plans, plans[1:]):
                a.connect(b)
            for p in plans:
                p.upload(s1, s2)
            for rep in range(3):
                for p in plans:
                    p.run()
            for p in plans:
                p.sync()
            assert plans[-1].score() == GOLDEN["fixtures"]["64gb"]["score"]
        finally:
            for p in plans:
                p.close()


@pytest.mark.parametrize("shape", [(1, 1), (40, 3), (3, 40), (700, 701), (5000, 2999), (2999, 5000), (20011, 9000)])
def test_score_mode_two_gpus(gpu, oracle, shape):
    # NW_MODE_SCORE, part 0 of 2: the forward half on device 0, the reversed half on device 1, cut along a staircase
    _need_gpus(gpu, 2)
    n1, n2 = shape
    s1, s2 = synth_pair(300 + n1 + n2, n1, n2, 5)
    want = oracle.score(s1, s2)
    with gpu.Plan(n1, n2, mode=gpu.NW_MODE_SCORE, part=0, nparts=2) as p:
        p.upload(s1, s2)
        for _ in range(2):
            p.run()
            assert p.score() == want
        assert p.time(2) > 0 and p.score() == want
    g1, g2 = synth_pair(5, n1, n2, 90)                    # a generic alphabet: horizontal cut, still one half per GPU
    with gpu.Plan(n1, n2, mode=gpu.NW_MODE_SCORE, part=0, nparts=2) as p:
        p.upload(g1, g2)
        p.run()
        assert p.score() == oracle.score(g1, g2)


@pytest.mark.parametrize("name", ["smid", "2gb", "mid", "big", "64gb"])
def test_score_mode_two_gpus_fixtures(gpu, name):
    _need_gpus(gpu, 2)
    s1, s2 = load_pair(name)
    with gpu.Plan(s1.size, s2.size, mode=gpu.NW_MODE_SCORE, part=0, nparts=2) as p:
        p.upload(s1, s2)
        p.run()
        assert p.score() == GOLDEN["fixtures"][name]["score"]
    if name == "smid":
        t = np.zeros((s2.size + 1) * (s1.size + 1), dtype=np.int32)
        gpu.needlemanWunsch(s1, s2, t, mode=gpu.NW_MODE_BOUNDARY, ngpus=2)      # the plug-in call on two GPUs
        assert t[-1] == GOLDEN["fixtures"][name]["score"]


@pytest.mark.parametrize("shape", [(1, 1), (33, 31), (700, 701), (5000, 2999), (2999, 5000), (9000, 20011), (20011, 9000)])
def test_score_mode_staircase_and_horizontal_cut_agree(gpu, oracle, monkeypatch, shape):
    # one GPU: the staircase is chosen while twice the strips still fit the schedulers; both cuts must give the oracle's score
    n1, n2 = shape
    s1, s2 = synth_pair(400 + n1 + n2, n1, n2, 5)
    want = oracle.score(s1, s2)
    for no_stair, force in (("0", "1"), ("1", "0"), ("0", "0")):
        monkeypatch.setenv("NW_CUDA_NO_STAIR", no_stair)
        monkeypatch.setenv("NW_CUDA_FORCE_STAIR", force)
        for R in (0, 2, 16):
            with gpu.Plan(n1, n2, mode=gpu.NW_MODE_SCORE, rows_per_lane=R) as p:
                p.upload(s1, s2)
                p.run()
                assert p.score() == want, (no_stair, force, R)


def test_score_mode_forced_staircase_64gb(gpu, monkeypatch):
    monkeypatch.setenv("NW_CUDA_FORCE_STAIR", "1")
    s1, s2 = load_pair("64gb")
    with gpu.Plan(s1.size, s2.size, mode=gpu.NW_MODE_SCORE) as p:
        p.upload(s1, s2)
        p.run()
        assert p.score() == GOLDEN["fixtures"]["64gb"]["score"]


# ---- API behaviour ------------------------------------------------------------------------------------------------------
def test_error_behaviour(gpu):
    s = np.ones(100, dtype=np.int8)
    with pytest.raises(gpu.NwCudaError):
        gpu.Plan(100, 100, mode=7)
    with pytest.raises(gpu.NwCudaError):
        gpu.Plan(100, 100, part=2, nparts=2)
    with pytest.raises(gpu.NwCudaError):
        gpu.Plan(5, 100, part=0, nparts=8)              # too narrow for 8 column strips
    with pytest.raises(gpu.NwCudaError):
        gpu.Plan(100, 100, rows_per_lane=3)
    with gpu.Plan(100, 100) as p:
        with pytest.raises(gpu.NwCudaError):
            p.run()                                      # no sequences uploaded
        p.upload(s, s)
        with pytest.raises(gpu.NwCudaError):
            p.score()                                    # no fill has been run
        with pytest.raises(gpu.NwCudaError):
            p.table_to_host()                            # not a full-table plan
        p.run()
        assert p.score() == 100
    assert b"" != gpu.lib().nw_cuda_last_error()


def test_device_resident_inputs(gpu, oracle):
    # sequences already in HBM (what bench.py's `value` times): nw_plan_upload_device / nw_batch_upload_device
    torch = pytest.importorskip("torch")
    s1, s2 = synth_pair(71, 5000, 4000, 5)
    d1, d2 = torch.from_numpy(s1).cuda(), torch.from_numpy(s2).cuda()
    with gpu.Plan(s1.size, s2.size) as p:
        p.upload_device(d1.data_ptr(), d2.data_ptr())
        p.run()
        assert p.score() == oracle.score(s1, s2)
    g1, g2 = synth_pair(72, 3000, 2500, 120)            # generic alphabet detected on the device
    e1, e2 = torch.from_numpy(g1).cuda(), torch.from_numpy(g2).cuda()
    with gpu.Plan(g1.size, g2.size, mode=gpu.NW_MODE_FULL) as p:
        p.upload_device(e1.data_ptr(), e2.data_ptr())
        p.run()
        assert np.array_equal(p.table_to_host(), oracle.fill(g1, g2))
    rng = np.random.default_rng(3)
    S1 = rng.integers(1, 5, size=(500, 333), dtype=np.int8)
    S2 = rng.integers(1, 5, size=(500, 1100), dtype=np.int8)
    b = gpu.Batch(500, 333, 1100)
    t1, t2 = torch.from_numpy(S1).cuda(), torch.from_numpy(S2).cuda()
    b.upload_device(t1.data_ptr(), t2.data_ptr())
    b.run()
    assert np.array_equal(b.scores(), oracle.batch_scores(S1, S2))
    b.close()


def test_dpx_peak_is_plausible(gpu):
    g, mhz = gpu.dpx_peak(0)
    sms = gpu.device_info(0)["sm_count"]
    lanes_per_clk_per_sm = g * 1e3 / (sms * mhz)
    assert 40 < lanes_per_clk_per_sm < 80, (g, mhz)      # B200: 64 int32 lanes per clock per SM


@pytest.mark.parametrize("shape", [(200, 31, 1000), (50, 1000, 31), (20, 513, 1025), (9, 2500, 40), (3, 7000, 6000)])
def test_batch_more_shapes(gpu, oracle, shape, kernel_kind):
    npairs, len1, len2 = shape
    rng = np.random.default_rng(77)
    S1 = rng.integers(1, 5, size=(npairs, len1), dtype=np.int8)
    S2 = rng.integers(1, 5, size=(npairs, len2), dtype=np.int8)
    S2[0, :min(len1, len2)] = S1[0, :min(len1, len2)]      # one near-identical pair: the largest scores
    assert np.array_equal(gpu.batch_scores(S1, S2), oracle.batch_scores(S1, S2))


# ---- traceback on the materialised table ---------------------------------------------------------------
@pytest.mark.parametrize("name", ["small", "t", "debug", "smid"])
def test_traceback_fixtures(gpu, oracle, name):
    s1, s2 = load_pair(name)
    with gpu.Plan(s1.size, s2.size, mode=gpu.NW_MODE_FULL) as p:
        p.upload(s1, s2)
        p.run()
        a1, a2 = p.traceback()
    b1, b2 = oracle.traceback(s1, s2)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2)
    gaps = int((a1 == 0).sum() + (a2 == 0).sum())
    assert int(((a1 == a2) & (a1 != 0)).sum()) - gaps == GOLDEN["fixtures"][name]["score"]
    assert set(gpu.printSequence(a1[:50])) <= set("-ATGC")


@pytest.mark.parametrize("shape", [(0, 0), (0, 9), (9, 0), (1, 1), (63, 64), (64, 63), (65, 129), (700, 90), (90, 700),
                                   (2000, 2100)])
def test_traceback_shapes(gpu, oracle, shape, kernel_kind):
    n1, n2 = shape
    s1, s2 = synth_pair(900 + n1 + 3 * n2, n1, n2, 5)
    with gpu.Plan(n1, n2, mode=gpu.NW_MODE_FULL) as p:
        p.upload(s1, s2)
        p.run()
        a1, a2 = p.traceback()
    b1, b2 = oracle.traceback(s1, s2)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2)


def test_traceback_2gb_properties(gpu):
    # full BASELINE size: the alignment reproduces both sequences and its column score is the golden score
    s1, s2 = load_pair("2gb")
    with gpu.Plan(s1.size, s2.size, mode=gpu.NW_MODE_FULL) as p:
        p.upload(s1, s2)
        p.run()
        a1, a2 = p.traceback()
        with pytest.raises(gpu.NwCudaError):
            gpu.Plan(10, 10).traceback()
    assert np.array_equal(a1[a1 != 0], s1) and np.array_equal(a2[a2 != 0], s2)
    gaps = int((a1 == 0).sum() + (a2 == 0).sum())
    assert int(((a1 == a2) & (a1 != 0)).sum()) - gaps == GOLDEN["fixtures"]["2gb"]["score"]


@pytest.mark.parametrize("shape", [(5003, 3001), (2999, 6007)])
def test_streamed_table_delivery(gpu, oracle, monkeypatch, shape):
    # one-shot full-table calls with a host table never hold the whole table on the device: pass 2 runs band by band
    # into a two-band ring while the previous band travels to the host.  Small bands force several of them here.
    monkeypatch.setenv("NW_CUDA_BAND_MB", "16")
    n1, n2 = shape
    s1, s2 = synth_pair(55 + n1, n1, n2, 5)
    t = np.full((n2 + 1, n1 + 1), -7, dtype=np.int32)
    gpu.needlemanWunsch(s1, s2, t)
    assert np.array_equal(t, oracle.fill(s1, s2))
    gpu.needlemanWunsch(s1, s2, t)                        # cached plan, second epoch
    assert np.array_equal(t, oracle.fill(s1, s2))


# ---- score-only mode: meeting in the middle -------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(0, 0), (0, 5), (5, 0), (1, 1), (1, 2), (40, 3), (3, 40), (700, 701), (5000, 2999),
                                   (2999, 5000)])
def test_score_mode_shapes(gpu, oracle, shape, kernel_kind):
    n1, n2 = shape
    s1, s2 = synth_pair(300 + n1 + 5 * n2, n1, n2, 5)
    with gpu.Plan(n1, n2, mode=gpu.NW_MODE_SCORE) as p:
        p.upload(s1, s2)
        p.run()
        assert p.score() == oracle.score(s1, s2)
        p.run(); p.run()
        assert p.score() == oracle.score(s1, s2)
        assert p.time(2) > 0
        with pytest.raises(gpu.NwCudaError):
            p.last_row()
    assert gpu.score(s1, s2) == oracle.score(s1, s2)          # the one-shot call uses the same mode


@pytest.mark.parametrize("name", ["smid", "2gb", "mid", "big", "64gb"])
def test_score_mode_fixtures(gpu, name):
    s1, s2 = load_pair(name)
    with gpu.Plan(s1.size, s2.size, mode=gpu.NW_MODE_SCORE) as p:
        p.upload(s1, s2)
        p.run()
        assert p.score() == GOLDEN["fixtures"][name]["score"]


def test_score_mode_generic_alphabet(gpu, oracle):
    rng = np.random.default_rng(17)
    s1 = rng.integers(-128, 128, size=1500, dtype=np.int8)
    s2 = np.concatenate([s1[:600], rng.integers(-128, 128, size=900, dtype=np.int8)]).astype(np.int8)
    with gpu.Plan(s1.size, s2.size, mode=gpu.NW_MODE_SCORE) as p:
        p.upload(s1, s2)
        p.run()
        assert p.score() == oracle.score(s1, s2)


def test_score_mode_rectangular(gpu, oracle):
    # score mode sweeps along the shorter sequence (the score is symmetric); both orientations, extreme aspect ratios
    for n1, n2 in [(30000, 300), (300, 30000), (9000, 1), (1, 9000), (5000, 0), (0, 5000)]:
        s1, s2 = synth_pair(n1 + 2 * n2, n1, n2, 5)
        assert gpu.score(s1, s2) == oracle.score(s1, s2), (n1, n2)


# ---- round 2: robustness of the plug-in path -------------------------------------------------------------------------------
def test_table_rows_wider_than_the_staging_buffer(gpu, oracle, monkeypatch):
    # a host row wider than the pinned staging buffer used to overrun it.  1 MB staging, rows of 1.2 MB.
    monkeypatch.setenv("NW_CUDA_STAGE_MB", "1")
    n1, n2 = 300_000, 40
    s1, s2 = synth_pair(77, n1, n2, 5)
    t = np.full((n2 + 1, n1 + 1), -9, dtype=np.int32)          # pageable
    gpu.needlemanWunsch(s1, s2, t)
    assert np.array_equal(t, oracle.fill(s1, s2))
    with gpu.Plan(n1, n2, mode=gpu.NW_MODE_FULL, part=1, nparts=2) as _:
        pass                                                   # (creation only: column parts deliver 2-D slices too)
    plans = [gpu.Plan(n1, n2, mode=gpu.NW_MODE_FULL, part=k, nparts=2, rows_per_lane=4) for k in range(2)]
    try:
        plans[0].connect(plans[1])
        out = np.zeros_like(t)
        for p in plans:
            p.upload(s1, s2)
            p.run()
            p.sync()
            p.table_to_host(out)
        assert np.array_equal(out, t)
    finally:
        for p in plans:
            p.close()
    monkeypatch.setenv("NW_CUDA_STAGE_MB", "32")
    gpu.needlemanWunsch(s1[:100], s2, np.empty((n2 + 1, 101), dtype=np.int32))      # back to the default buffers


def test_failing_fill_prints_to_stderr_and_exits_2(gpu):
    # no device -> the reference's driver around our entry point prints no score and exits with status 2 (no CPU fallback)
    import os
    import subprocess
    from conftest import ROOT, pair_paths
    exe = os.path.join(ROOT, "fast-needleman-wunsch_b200", "bin", "cuda.e")
    if not os.path.exists(exe):
        pytest.skip("cuda.e not built (no reference tree at build time)")
    a, b = pair_paths("smid")
    out = subprocess.run([exe, a, b], capture_output=True, text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 2, (out.returncode, out.stdout, out.stderr)
    assert "Score" not in out.stdout and out.stderr.startswith("cuda:")
    out = subprocess.run([exe, a, b], capture_output=True, text=True, env=dict(os.environ, NW_CUDA_MODE="sideways"))
    assert out.returncode == 2 and "Score" not in out.stdout and "NW_CUDA_MODE" in out.stderr


def test_a_part_without_its_left_neighbour_times_out_instead_of_hanging(gpu, oracle, monkeypatch):
    # wait loops are bounded.  Part 1 of 2 run alone never receives its halo: the fill gives up after
    # NW_CUDA_SPIN_TIMEOUT_MS, the call fails, and the library keeps working afterwards.
    monkeypatch.setenv("NW_CUDA_SPIN_TIMEOUT_MS", "300")
    s1, s2 = synth_pair(5, 4000, 3000, 5)
    with gpu.Plan(s1.size, s2.size, part=1, nparts=2, rows_per_lane=8) as p:
        p.upload(s1, s2)
        p.run()
        with pytest.raises(gpu.NwCudaError, match="aborted"):
            p.sync()
    monkeypatch.setenv("NW_CUDA_SPIN_TIMEOUT_MS", "20000")
    assert gpu.score(s1, s2) == oracle.score(s1, s2)


def test_one_shot_cache_is_dropped_after_a_failure(gpu, oracle):
    # a failed one-shot call must not leave half-advanced cached plans behind
    s1, s2 = synth_pair(6, 3000, 2500, 5)
    t = np.empty((s2.size + 1, s1.size + 1), dtype=np.int32)
    gpu.needlemanWunsch(s1, s2, t)
    with pytest.raises(gpu.NwCudaError):
        gpu.needlemanWunsch(s1, s2, t, ngpus=64)               # more GPUs than there are
    gpu.needlemanWunsch(s1, s2, t)
    assert np.array_equal(t, oracle.fill(s1, s2))


def test_strip_trace(gpu):
    s1, s2 = synth_pair(8, 20000, 2048, 5)
    with gpu.Plan(s1.size, s2.size, rows_per_lane=8) as p:
        p.upload(s1, s2)
        p.run()
        p.sync()
        a, b, cyc = p.strip_times(cycles=True)
        assert len(a) == p.strip_info()["nstrips"] == 8
        assert (b > a).all() and (np.diff(a) > 0).all() and (cyc > 0).all()      # strips start one after the other
        mhz = cyc / (b - a) * 1e3
        assert ((mhz > 500) & (mhz < 3000)).all()


def test_random_stress(gpu, oracle, monkeypatch):
    # tools/stress.py as a test: ~200 seeded random cases over every mode against the oracle
    rng = np.random.default_rng(20260101)
    for case in range(200):
        kind = int(rng.integers(0, 6))
        hi = int(rng.choice([2, 3, 5, 5, 5, 5, 9, 100]))
        monkeypatch.setenv("NW_CUDA_LAG2", str(int(rng.integers(0, 2))))
        if kind <= 3:
            n1, n2 = int(rng.integers(0, 3000)), int(rng.integers(0, 3000))
            if rng.random() < 0.15:
                n1 = int(rng.integers(0, 40000))
            if rng.random() < 0.15:
                n2 = int(rng.integers(0, 15000))
            s1 = rng.integers(1, hi, size=n1, dtype=np.int8)
            s2 = rng.integers(1, hi, size=n2, dtype=np.int8)
            if rng.random() < 0.3 and min(n1, n2) > 10:        # long common stretch: fastest growth of G
                k = int(rng.integers(1, min(n1, n2)))
                s2[:k] = s1[:k]
            R = int(rng.choice([0, 0, 1, 2, 4, 8, 16]))
            if hi > 5 and R == 16:
                R = 8
            row, col, sc, _ = oracle.boundaries(s1, s2)
            tag = (case, kind, n1, n2, hi, R)
            if kind == 0:
                with gpu.Plan(n1, n2, rows_per_lane=R) as p:
                    p.upload(s1, s2)
                    p.run()
                    assert np.array_equal(p.last_row(), row) and np.array_equal(p.last_col(), col) and p.score() == sc, tag
            elif kind == 1:
                with gpu.Plan(n1, n2, mode=gpu.NW_MODE_SCORE, rows_per_lane=R) as p:
                    p.upload(s1, s2)
                    p.run()
                    assert p.score() == sc, tag
            elif kind == 2 and (n1 + 1) * (n2 + 1) < 30_000_000:
                monkeypatch.setenv("NW_CUDA_TILE_BLOCKS", str(int(rng.choice([2, 5, 32]))))
                with gpu.Plan(n1, n2, mode=gpu.NW_MODE_FULL, rows_per_lane=R) as p:
                    p.upload(s1, s2)
                    p.run()
                    assert np.array_equal(p.table_to_host(), oracle.fill(s1, s2)), tag
            elif kind == 3 and n1 >= 64:
                P = int(rng.choice([2, 3, 5]))
                plans = [gpu.Plan(n1, n2, part=k, nparts=P, rows_per_lane=R if R else 4) for k in range(P)]
                try:
                    for a, b in zip(plans, plans[1:]):
                        a.connect(b)
                    for p in plans:
                        p.upload(s1, s2)
                    for p in plans:
                        p.run()
                        p.sync()
                    assert plans[-1].score() == sc and np.array_equal(plans[-1].last_col(), col), tag + (P,)
                finally:
                    for p in plans:
                        p.close()
        else:
            npairs, l1, l2 = int(rng.integers(1, 200)), int(rng.integers(0, 1500)), int(rng.integers(0, 1500))
            S1 = rng.integers(1, hi, size=(npairs, l1), dtype=np.int8)
            S2 = rng.integers(1, hi, size=(npairs, l2), dtype=np.int8)
            assert np.array_equal(gpu.batch_scores(S1, S2), oracle.batch_scores(S1, S2)), (case, npairs, l1, l2, hi)


@pytest.mark.parametrize("what", ["synthetic", "64gb"])
def test_multi_process_ipc_pipeline(gpu, what):
    # one process per GPU, CUDA-IPC halo mailboxes: the path `bench.py --gpus N` and the driver's SCALE run use
    _need_gpus(gpu, 2)
    import os
    import subprocess
    import sys
    from conftest import ROOT
    n = min(gpu.device_count(), 8)
    for world in sorted({2, n}):
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                              "--master-addr", "127.0.0.1", "--master-port", "29541",
                              os.path.join(ROOT, "tests", "mgpu_worker.py"), what],
                             capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "mgpu_worker ok" in out.stdout, (world, out.stdout[-1500:], out.stderr[-3000:])
