"""CPU: the C-ABI library loads, exports every symbol include/nw_cuda.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nw_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nw_(?:cuda|plans?|batch)_\w+)\s*\(", src)))


def test_header_declares_expected_surface(nw):
    syms = declared_symbols()
    assert "nw_cuda_fill" in syms and "nw_plan_run" in syms and "nw_batch_run" in syms
    assert sorted(nw.nwcuda.EXPORTS) == syms


def test_library_exports_every_declared_symbol(nw):
    assert os.path.exists(nw.lib_path), "libnw_cuda.so has not been built (make lib)"
    L = C.CDLL(nw.lib_path)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/nw_cuda.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", nw.lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (nw_\w+)", out))
    assert set(declared_symbols()) <= exported


def test_version_and_argument_errors(nw):
    assert b"sm_100a" in nw.lib().nw_cuda_version()
    with pytest.raises(TypeError):
        nw.score(np.zeros(4, dtype=np.int32), np.zeros(4, dtype=np.int8))
    with pytest.raises(ValueError):
        nw.needlemanWunsch(np.ones(4, np.int8), np.ones(4, np.int8), np.zeros(3, np.int32))


def test_no_cpu_fallback_without_device(nw):
    # in the build container there is no GPU: every compute entry point must fail with an error, never compute
    try:
        n = nw.device_count()
    except nw.NwCudaError:
        n = 0
    if n > 0:
        pytest.skip("a CUDA device is visible")
    s = np.ones(8, dtype=np.int8)
    with pytest.raises(nw.NwCudaError):
        nw.score(s, s)
    with pytest.raises(nw.NwCudaError):
        nw.needlemanWunsch(s, s)
    with pytest.raises(nw.NwCudaError):
        nw.batch_scores(s.reshape(2, 4), s.reshape(2, 4))
    # the widened entry points: scoring parameters, Smith-Waterman, alignment without a table, plans
    for call in (lambda: nw.score(s, s, scoring=(2, -1, -2)), lambda: nw.best(s, s, (2, -1, -2, 1)),
                 lambda: nw.needlemanWunsch(s, s, scoring=(5, -4, -3)), lambda: nw.align(s, s),
                 lambda: nw.batch_scores(s.reshape(2, 4), s.reshape(2, 4), scoring=(2, -1, -2)),
                 lambda: nw.Plan(8, 8), lambda: nw.Plan(8, 8, mode=nw.NW_MODE_SCORE, nparts=2),
                 lambda: nw.Plan(8, 8, scoring=(1, 0, -1, 1)), lambda: nw.dpx_peak(0)):
        with pytest.raises(nw.NwCudaError):
            call()


def test_scoring_arguments_are_validated_before_any_device_work(nw):
    # argument errors carry their own code and text, with or without a device
    s = np.ones(8, dtype=np.int8)
    for bad in ((1, 0, 1, 1), (1, 0, -1, 7)):
        with pytest.raises(nw.NwCudaError, match="-2"):
            nw.best(s, s, bad)
    sc = nw.Scoring(1, 0, -1, 0)
    sc.reserved[2] = 5
    with pytest.raises(nw.NwCudaError, match="reserved"):
        nw.score(s, s, scoring=sc)
    with pytest.raises(nw.NwCudaError, match="overflow"):
        nw.Plan(1 << 20, 1 << 20, scoring=(5000, 0, -1))


def test_product_never_touches_the_oracle(nw):
    # the product sources and the shipped library must not reference oracle/ in any way
    pkg = os.path.join(ROOT, "fast-needleman-wunsch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".py", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "nw_oracle_" not in txt and "libnw_ref" not in txt, f
    out = subprocess.run(["nm", "-D", nw.lib_path], capture_output=True, text=True, check=True).stdout
    assert "nw_oracle" not in out and "nw_ref" not in out


def test_strip_partition_matches_reference_formula(nw, oracle):
    for n1 in (7, 100, 1003, 126440):
        for P in (1, 2, 3, 4, 8):
            if (n1 + 1) // P < 2:
                continue
            for p in range(P):
                assert nw.strip_partition(n1, P, p) == oracle.strip_partition(n1, P, p)
