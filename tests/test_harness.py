"""CPU: the sweep harness (tools/harness.py) writes the reference's TSV layouts and reads them back with the column logic
of the reference's data/graph.py:33-50.  The driver binaries are replaced by stub programs that print what a driver
prints (src/common/driver.cpp:33-35): "<ms>\\nScore: <n>\\n"."""
import os
import stat
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tools", "harness.py")
sys.path.insert(0, os.path.join(ROOT, "tools"))
import harness      # noqa: E402


@pytest.fixture()
def stubs(tmp_path):
    """cuda.e / serial.e stubs: ms = file size of the first sequence x a factor that depends on the knobs in the environment."""
    exe_dir, bdna = tmp_path / "bin", tmp_path / "bdna"
    exe_dir.mkdir()
    bdna.mkdir()
    for g in (2, 4, 6):
        (bdna / f"{g}gb-1.bdna").write_bytes(b"\x01" * g)
        (bdna / f"{g}gb-2.bdna").write_bytes(b"\x02" * g)
    for name, base in (("cuda.e", 10), ("serial.e", 1000)):
        p = exe_dir / name
        p.write_text("#!/bin/sh\n"
                     "n=$(wc -c < \"$1\")\n"
                     f"ms=$(( n * {base} / ${{NW_CUDA_GPUS:-1}} + ${{NW_CUDA_R:-0}} ))\n"
                     "[ \"$NW_CUDA_MODE\" = full ] && ms=$(( ms * 3 ))\n"
                     "printf '%s\\nScore: 7\\n' \"$ms\"\n")
        p.chmod(p.stat().st_mode | stat.S_IEXEC)
    return str(exe_dir), str(bdna)


def run(args, cwd):
    out = subprocess.run([sys.executable, HARNESS] + args, capture_output=True, text=True, cwd=cwd)
    assert out.returncode == 0, out.stderr
    return out


def test_sizes_table(stubs, tmp_path):
    exe_dir, bdna = stubs
    run(["--runs", "2", "--bdna", bdna, "--exe-dir", exe_dir, "sizes", "--min", "2", "--max", "6", "--step", "2",
         "--programs", "cuda,cuda-full,serial", "--max-full-gb", "4", "-o", "sizes.tsv"], tmp_path)
    title, x, y = harness.parse_tsv(tmp_path / "sizes.tsv")
    assert title == "benchmarking cuda, cuda-full, serial" and x == [2, 4, 6]
    assert y == {"cuda": [20, 40, 60], "cuda-full": [60, 120, 0], "serial": [2000, 4000, 6000]}
    # the speed-up transform of data/graph.py:45-50 works on it
    assert [y["serial"][i] / y["cuda"][i] for i in range(3)] == [100.0, 100.0, 100.0]
    # like the reference's scripts, an existing result file is never overwritten
    out = subprocess.run([sys.executable, HARNESS, "--bdna", bdna, "--exe-dir", exe_dir, "sizes", "-o", "sizes.tsv"],
                         capture_output=True, text=True, cwd=tmp_path)
    assert out.returncode != 0 and "already exists" in out.stderr


def test_gpu_count_table(stubs, tmp_path):
    exe_dir, bdna = stubs
    run(["--runs", "1", "--bdna", bdna, "--exe-dir", exe_dir, "gpus", "--min", "2", "--max", "4", "--step", "2",
         "--gpu-counts", "1,2", "--mode", "boundary", "-o", "threads.tsv"], tmp_path)
    _, x, y = harness.parse_tsv(tmp_path / "threads.tsv")
    assert x == [2, 4] and y == {"serial": [2000, 4000], "1": [20, 40], "2": [10, 20]}      # layout of benchmark-threads.sh


def test_tuning_files(stubs, tmp_path):
    exe_dir, bdna = stubs
    run(["--runs", "1", "--bdna", bdna, "--exe-dir", exe_dir, "tune", "--sizes", "2,6", "--knob", "rows_per_lane",
         "--values", "4,8,16", "--mode", "full", "--prefix", "rtune"], tmp_path)
    assert harness.parse_tune_tsv(tmp_path / "rtune2.tsv") == ("2gb", [4, 8, 16], [72, 84, 108])
    assert harness.parse_tune_tsv(tmp_path / "rtune6.tsv") == ("6gb", [4, 8, 16], [192, 204, 228])
    lines = (tmp_path / "rtune2.tsv").read_text().splitlines()
    assert lines[1].startswith("bufsize\t") and lines[2].startswith("time\t")                 # src/buf-tune.sh:28-49


def test_reads_the_reference_s_own_tables():
    ref = "/root/reference/data/multi.tsv"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present")
    _, x, y = harness.parse_tsv(ref)
    assert x[0] == 2 and x[-1] == 64 and y["serial"][0] == 1380 and y["hybrid"][-1] == 2996
