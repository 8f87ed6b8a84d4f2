import ctypes as C
import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
BDNA = os.path.join(ROOT, "oracle", "_ref", "bdna")
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
GOLDEN_SCORING = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_scoring.json")))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pair_paths(name):
    if name.endswith("gb"):
        return os.path.join(BDNA, f"{name}-1.bdna"), os.path.join(BDNA, f"{name}-2.bdna")
    return os.path.join(BDNA, f"{name}1.bdna"), os.path.join(BDNA, f"{name}2.bdna")


def load_pair(name):
    a, b = pair_paths(name)
    if not (os.path.exists(a) and os.path.exists(b)):
        pytest.skip(f"bdna fixture {name} not staged (run `make -C oracle ref` where /root/reference exists)")
    return np.fromfile(a, dtype=np.int8), np.fromfile(b, dtype=np.int8)


def synth_pair(seed, n1, n2, hi):
    """Same generator as tests/golden/make_golden.py."""
    rng = np.random.default_rng(seed)
    s1 = rng.integers(1, hi, size=n1, dtype=np.int8)
    s2 = rng.integers(1, hi, size=n2, dtype=np.int8)
    return s1, s2


class Oracle:
    """ctypes view of oracle/liboracle.so (the CPU restatement) -- the CHECKER, used by tests only."""

    def __init__(self):
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
        L = C.CDLL(path)
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.nw_oracle_fill.argtypes = [vp, i32, vp, i32, vp]
        L.nw_oracle_fill.restype = None
        L.nw_oracle_boundaries.argtypes = [vp, i32, vp, i32, vp, vp, i32, vp]
        L.nw_oracle_boundaries.restype = i32
        L.nw_oracle_score.argtypes = [vp, i32, vp, i32]
        L.nw_oracle_score.restype = i32
        L.nw_oracle_batch_scores.argtypes = [vp, vp, i64, i32, i32, vp]
        L.nw_oracle_batch_scores.restype = None
        L.nw_oracle_strip_partition.argtypes = [i32, i32, i32, C.POINTER(i64), C.POINTER(i64)]
        L.nw_oracle_strip_partition.restype = None
        L.nw_oracle_strip.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp]
        L.nw_oracle_strip.restype = i32
        L.nw_oracle_traceback.argtypes = [vp, i32, vp, i32, vp, vp]
        L.nw_oracle_traceback.restype = i32
        L.nw_oracle_fill_ex.argtypes = [vp, i32, vp, i32, i32, i32, i32, i32, vp]
        L.nw_oracle_fill_ex.restype = None
        L.nw_oracle_score_ex.argtypes = [vp, i32, vp, i32, i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
        L.nw_oracle_score_ex.restype = i32
        L.nw_oracle_traceback_ex.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, vp]
        L.nw_oracle_traceback_ex.restype = i32
        L.nw_oracle_traceback_local.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, vp]
        L.nw_oracle_traceback_local.restype = i32
        L.nw_oracle_fnv1a64.argtypes = [vp, i64]
        L.nw_oracle_fnv1a64.restype = C.c_uint64
        L.nw_oracle_fnv1a64_col.argtypes = [vp, i64, i64]
        L.nw_oracle_fnv1a64_col.restype = C.c_uint64
        self.L = L

    @staticmethod
    def _p(a):
        return a.ctypes.data if a.size else None

    def fill(self, s1, s2):
        t = np.empty((s2.size + 1, s1.size + 1), dtype=np.int32)
        self.L.nw_oracle_fill(self._p(s1), s1.size, self._p(s2), s2.size, t.ctypes.data)
        return t

    def score(self, s1, s2):
        return int(self.L.nw_oracle_score(self._p(s1), s1.size, self._p(s2), s2.size))

    def boundaries(self, s1, s2, row_stride=0):
        row = np.empty(s1.size + 1, dtype=np.int32)
        col = np.empty(s2.size + 1, dtype=np.int32)
        nk = (s2.size // row_stride) if row_stride else 0
        rows = np.empty((max(nk, 1), s1.size + 1), dtype=np.int32)
        sc = self.L.nw_oracle_boundaries(self._p(s1), s1.size, self._p(s2), s2.size, row.ctypes.data, col.ctypes.data,
                                         row_stride, rows.ctypes.data if nk else None)
        return row, col, int(sc), rows[:nk]

    def batch_scores(self, S1, S2):
        out = np.empty(S1.shape[0], dtype=np.int32)
        self.L.nw_oracle_batch_scores(self._p(S1), self._p(S2), S1.shape[0], S1.shape[1], S2.shape[1],
                                      out.ctypes.data if out.size else None)
        return out

    def strip_partition(self, n1, P, p):
        s, n = C.c_int64(), C.c_int64()
        self.L.nw_oracle_strip_partition(n1, P, p, C.byref(s), C.byref(n))
        return s.value, n.value

    def strip(self, s1, s2, P, p, halo):
        right = np.empty(s2.size + 1, dtype=np.int32)
        last = self.L.nw_oracle_strip(self._p(s1), s1.size, self._p(s2), s2.size, P, p,
                                      halo.ctypes.data if halo is not None else None, right.ctypes.data)
        return right, int(last)

    def traceback(self, s1, s2):
        a1 = np.empty(s1.size + s2.size + 1, dtype=np.int8)
        a2 = np.empty(s1.size + s2.size + 1, dtype=np.int8)
        n = self.L.nw_oracle_traceback(self._p(s1), s1.size, self._p(s2), s2.size, a1.ctypes.data, a2.ctypes.data)
        return a1[:n].copy(), a2[:n].copy()

    def fill_ex(self, s1, s2, scoring):
        """scoring = (match, mismatch, gap[, local])"""
        m, x, g, local = (list(scoring) + [0])[:4]
        t = np.empty((s2.size + 1, s1.size + 1), dtype=np.int32)
        self.L.nw_oracle_fill_ex(self._p(s1), s1.size, self._p(s2), s2.size, m, x, g, local, t.ctypes.data)
        return t

    def score_ex(self, s1, s2, scoring):
        """(score, end_i, end_j)"""
        m, x, g, local = (list(scoring) + [0])[:4]
        i, j = C.c_int32(), C.c_int32()
        sc = self.L.nw_oracle_score_ex(self._p(s1), s1.size, self._p(s2), s2.size, m, x, g, local, C.byref(i), C.byref(j))
        return int(sc), i.value, j.value

    def traceback_ex(self, s1, s2, scoring):
        m, x, g = scoring[:3]
        a1 = np.empty(s1.size + s2.size + 1, dtype=np.int8)
        a2 = np.empty(s1.size + s2.size + 1, dtype=np.int8)
        n = self.L.nw_oracle_traceback_ex(self._p(s1), s1.size, self._p(s2), s2.size, m, x, g, a1.ctypes.data, a2.ctypes.data)
        return a1[:n].copy(), a2[:n].copy()

    def traceback_local(self, s1, s2, scoring):
        m, x, g = scoring[:3]
        a1 = np.empty(s1.size + s2.size + 1, dtype=np.int8)
        a2 = np.empty(s1.size + s2.size + 1, dtype=np.int8)
        n = self.L.nw_oracle_traceback_local(self._p(s1), s1.size, self._p(s2), s2.size, m, x, g, a1.ctypes.data, a2.ctypes.data)
        return a1[:n].copy(), a2[:n].copy()

    def fnv(self, a):
        a = np.ascontiguousarray(a)
        return f"{self.L.nw_oracle_fnv1a64(a.ctypes.data, a.nbytes):016x}"

    def table_facts(self, t):
        nrows, ncols = t.shape
        return {
            "rows": int(nrows), "cols": int(ncols), "score": int(t[-1, -1]),
            "sum": int(t.sum(dtype=np.int64)), "min": int(t.min()), "max": int(t.max()),
            "fnv_table": self.fnv(t), "fnv_lastrow": self.fnv(t[-1]), "fnv_lastcol": self.fnv(t[:, -1].copy()),
        }


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    return GOLDEN


@pytest.fixture(scope="session")
def nw():
    return importlib.import_module("fast-needleman-wunsch_b200")


@pytest.fixture(scope="session")
def gpu(nw):
    """The product library on a real device.  Fails (does not skip) when the extension is missing on a GPU box."""
    if not os.path.exists(nw.lib_path):
        pytest.fail(f"{nw.lib_path} is missing: the CUDA extension was not built")
    n = nw.device_count()
    assert n >= 1, "no CUDA device visible"
    nw.init(0)
    return nw
