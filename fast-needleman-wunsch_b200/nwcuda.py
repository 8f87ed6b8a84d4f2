"""ctypes binding of libnw_cuda.so -- host-side mirror of the reference's interface for the NW fill.

Names follow the reference: `readSequence` (src/common/helper.cpp:3-25) and `needlemanWunsch(s1, s2, t)`
(src/serial/serial.cpp:4).  Everything computes on the GPU through the C ABI of include/nw_cuda.h; there is no CPU
path here, and a missing library or device raises instead of falling back.
"""
import ctypes as C
import os

import numpy as np

NW_MODE_BOUNDARY = 0
NW_MODE_FULL = 1
NW_MODE_SCORE = 2      # score only, meeting in the middle (two half-length dependency chains)

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.environ.get("NW_CUDA_LIB", os.path.join(_HERE, "libnw_cuda.so"))   # override: A/B builds in development

EXPORTS = [
    "nw_cuda_version", "nw_cuda_last_error", "nw_cuda_device_count", "nw_cuda_init", "nw_cuda_device_info",
    "nw_cuda_fill", "nw_cuda_fill_ex", "nw_cuda_score", "nw_cuda_boundaries", "nw_cuda_batch_scores",
    "nw_plan_create", "nw_plan_destroy", "nw_plan_upload", "nw_plan_upload_device", "nw_plan_connect",
    "nw_plan_export_mailbox", "nw_plan_import_mailbox", "nw_plan_run", "nw_plan_sync", "nw_plan_time",
    "nw_plan_timer_start", "nw_plan_timer_stop", "nw_plan_last_ms", "nw_plan_launches_per_run", "nw_plan_score", "nw_plan_last_row", "nw_plan_last_col",
    "nw_plan_table_to_host", "nw_plan_table_device", "nw_plan_traceback", "nw_plan_strip_info", "nw_plan_strip_row", "nw_plan_strip_times",
    "nw_batch_create", "nw_batch_destroy", "nw_batch_upload", "nw_batch_upload_device", "nw_batch_run",
    "nw_batch_sync", "nw_batch_time", "nw_batch_scores", "nw_cuda_dpx_peak",
    "nw_cuda_fill_scored", "nw_cuda_score_scored", "nw_cuda_batch_scores_scored", "nw_plan_create_scored", "nw_plan_best",
    "nw_batch_set_scoring", "nw_plans_traceback", "nw_cuda_align", "nw_batch_run_host",
]


class NwCudaError(RuntimeError):
    pass


class Tuning(C.Structure):
    _fields_ = [("rows_per_lane", C.c_int), ("warps_per_cta", C.c_int), ("ctas", C.c_int), ("reserved", C.c_int * 5)]


class Scoring(C.Structure):
    """nw_scoring: match / mismatch / gap (the reference's macros, src/common/needleman-wunsch.hpp:11-13, are 1, 0, -1);
    local=1 selects Smith-Waterman."""
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap", C.c_int32), ("local", C.c_int32),
                ("reserved", C.c_int32 * 4)]


def _scoring(scoring):
    """None | Scoring | (match, mismatch, gap[, local]) -> pointer argument"""
    if scoring is None:
        return None
    if not isinstance(scoring, Scoring):
        scoring = Scoring(*[int(x) for x in scoring])
    return C.byref(scoring)


_lib = None


def lib():
    """The loaded library.  Raises NwCudaError when it has not been built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path):
            raise NwCudaError(f"{lib_path} is missing: run `make lib` (or __graft_entry__.build()) first")
        L = C.CDLL(lib_path)
        L.nw_cuda_version.restype = C.c_char_p
        L.nw_cuda_last_error.restype = C.c_char_p
        vp, i32, i64, ip = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_int)
        sigs = {
            "nw_cuda_init": [C.c_int],
            "nw_cuda_device_info": [C.c_int, C.c_char_p, C.c_int, ip, ip],
            "nw_cuda_fill": [vp, i32, vp, i32, vp],
            "nw_cuda_fill_ex": [vp, i32, vp, i32, vp, C.c_int, C.c_int],
            "nw_cuda_score": [vp, i32, vp, i32, vp],
            "nw_cuda_boundaries": [vp, i32, vp, i32, vp, vp, vp],
            "nw_cuda_batch_scores": [vp, vp, i64, i32, i32, vp, C.c_int],
            "nw_plan_create": [C.POINTER(vp), C.c_int, i32, i32, C.c_int, C.c_int, C.c_int, C.POINTER(Tuning)],
            "nw_plan_destroy": [vp], "nw_plan_upload": [vp, vp, vp], "nw_plan_upload_device": [vp, vp, vp],
            "nw_plan_connect": [vp, vp], "nw_plan_export_mailbox": [vp, vp],
            "nw_plan_import_mailbox": [vp, vp, C.c_int],
            "nw_plan_run": [vp], "nw_plan_sync": [vp], "nw_plan_time": [vp, C.c_int, C.POINTER(C.c_float)],
            "nw_plan_timer_start": [vp], "nw_plan_timer_stop": [vp, C.POINTER(C.c_float)],
            "nw_plan_last_ms": [vp, C.POINTER(C.c_float)], "nw_plan_launches_per_run": [vp, ip],
            "nw_plan_score": [vp, vp], "nw_plan_last_row": [vp, vp], "nw_plan_last_col": [vp, vp],
            "nw_plan_table_to_host": [vp, vp], "nw_plan_table_device": [vp, C.POINTER(vp), C.POINTER(i64)],
            "nw_plan_traceback": [vp, vp, vp, ip],
            "nw_plan_strip_info": [vp, ip, ip, ip, ip, ip], "nw_plan_strip_row": [vp, C.c_int, vp],
            "nw_plan_strip_times": [vp, vp, vp, vp],
            "nw_batch_create": [C.POINTER(vp), C.c_int, i64, i32, i32], "nw_batch_destroy": [vp],
            "nw_batch_upload": [vp, vp, vp], "nw_batch_upload_device": [vp, vp, vp], "nw_batch_run": [vp],
            "nw_batch_sync": [vp], "nw_batch_time": [vp, C.c_int, C.POINTER(C.c_float)],
            "nw_batch_scores": [vp, vp],
            "nw_cuda_dpx_peak": [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)],
            "nw_cuda_fill_scored": [vp, i32, vp, i32, vp, C.c_int, C.c_int, C.POINTER(Scoring)],
            "nw_cuda_score_scored": [vp, i32, vp, i32, C.POINTER(Scoring), vp, vp, vp],
            "nw_cuda_batch_scores_scored": [vp, vp, i64, i32, i32, C.POINTER(Scoring), vp, C.c_int],
            "nw_plan_create_scored": [C.POINTER(vp), C.c_int, i32, i32, C.c_int, C.c_int, C.c_int, C.POINTER(Tuning),
                                      C.POINTER(Scoring)],
            "nw_plan_best": [vp, vp, vp, vp],
            "nw_batch_set_scoring": [vp, C.POINTER(Scoring)],
            "nw_batch_run_host": [vp, vp, vp, vp, C.c_int],
            "nw_plans_traceback": [C.POINTER(vp), C.c_int, vp, vp, ip],
            "nw_cuda_align": [vp, i32, vp, i32, C.POINTER(Scoring), vp, vp, ip, ip],
        }
        for name, args in sigs.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = C.c_int
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise NwCudaError(f"nw_cuda error {rc}: {lib().nw_cuda_last_error().decode(errors='replace')}")


def _seq(a):
    a = np.ascontiguousarray(a)
    if a.dtype not in (np.int8, np.uint8):
        raise TypeError("sequences are 1-byte-per-base arrays (int8), like the reference's dnaArray.dna")
    return a


def _ptr(a):
    return a.ctypes.data if a.size else None


def readSequence(fileName):
    """bdna loader: raw bytes, one base per byte, no header (reference: src/common/helper.cpp:3-25).
    Like the reference, a file that cannot be opened is an error carrying the file name."""
    try:
        return np.fromfile(fileName, dtype=np.int8)
    except (FileNotFoundError, OSError) as e:
        raise FileNotFoundError(fileName) from e


def printSequence(a):
    """The reference's printer as a string: 0..4 -> "-ATGC" (src/common/helper.cpp:27-34)."""
    return "".join("-ATGC"[int(x)] for x in a)


def device_count():
    n = lib().nw_cuda_device_count()
    if n < 0:
        _ck(n)
    return n


def init(device=0):
    _ck(lib().nw_cuda_init(device))


def device_info(device=0):
    name = C.create_string_buffer(256)
    sms, mhz = C.c_int(), C.c_int()
    _ck(lib().nw_cuda_device_info(device, name, 256, C.byref(sms), C.byref(mhz)))
    return {"name": name.value.decode(), "sm_count": sms.value, "sm_clock_mhz": mhz.value}


def needlemanWunsch(s1, s2, t=None, mode=NW_MODE_FULL, ngpus=1, scoring=None):
    """Fill the table like the reference's needlemanWunsch(dnaArray s1, dnaArray s2, int* t)
    (src/serial/serial.cpp:4-36): s1 across (columns), s2 down (rows), t row-major int32 (n2+1) x (n1+1).
    In boundary mode only t[-1, -1] (the score, driver.cpp:35) is written.  Returns t."""
    s1, s2 = _seq(s1), _seq(s2)
    if t is None:
        t = np.empty((s2.size + 1, s1.size + 1), dtype=np.int32)
    if t.dtype != np.int32 or not t.flags.c_contiguous or t.size != (s1.size + 1) * (s2.size + 1):
        raise ValueError("t must be a C-contiguous int32 array of (n2+1)*(n1+1) elements")
    if scoring is None:
        _ck(lib().nw_cuda_fill_ex(_ptr(s1), s1.size, _ptr(s2), s2.size, t.ctypes.data, mode, ngpus))
    else:
        _ck(lib().nw_cuda_fill_scored(_ptr(s1), s1.size, _ptr(s2), s2.size, t.ctypes.data, mode, ngpus, _scoring(scoring)))
    return t


def score(s1, s2, scoring=None):
    s1, s2 = _seq(s1), _seq(s2)
    out = C.c_int32()
    if scoring is None:
        _ck(lib().nw_cuda_score(_ptr(s1), s1.size, _ptr(s2), s2.size, C.byref(out)))
    else:
        _ck(lib().nw_cuda_score_scored(_ptr(s1), s1.size, _ptr(s2), s2.size, _scoring(scoring), C.byref(out), None, None))
    return out.value


def best(s1, s2, scoring):
    """(score, end_i, end_j): for local alignment (scoring.local = 1) the best cell of the Smith-Waterman table (smallest
    column, then smallest row among equals; (0, 0) when the score is 0); for global alignment (H[n2][n1], n2, n1)."""
    s1, s2 = _seq(s1), _seq(s2)
    sc, i, j = C.c_int32(), C.c_int32(), C.c_int32()
    _ck(lib().nw_cuda_score_scored(_ptr(s1), s1.size, _ptr(s2), s2.size, _scoring(scoring), C.byref(sc), C.byref(i), C.byref(j)))
    return sc.value, i.value, j.value


def boundaries(s1, s2):
    """(last_row H[n2][0..n1], last_col H[0..n2][n1], score) without materialising the table."""
    s1, s2 = _seq(s1), _seq(s2)
    row = np.empty(s1.size + 1, dtype=np.int32)
    col = np.empty(s2.size + 1, dtype=np.int32)
    out = C.c_int32()
    _ck(lib().nw_cuda_boundaries(_ptr(s1), s1.size, _ptr(s2), s2.size, row.ctypes.data, col.ctypes.data, C.byref(out)))
    return row, col, out.value


def align(s1, s2, scoring=None):
    """(a1, a2, score): the gapped s1 and s2 (gap = 0, README.md:8) of the optimal global alignment, computed without a
    table (checkpoint rows and columns + tile replay, nw_cuda_align)."""
    s1, s2 = _seq(s1), _seq(s2)
    cap = s1.size + s2.size + 1
    a1, a2 = np.empty(cap, dtype=np.int8), np.empty(cap, dtype=np.int8)
    n, sc = C.c_int(), C.c_int()
    _ck(lib().nw_cuda_align(_ptr(s1), s1.size, _ptr(s2), s2.size, _scoring(scoring), a1.ctypes.data, a2.ctypes.data,
                            C.byref(n), C.byref(sc)))
    return a1[:n.value].copy(), a2[:n.value].copy(), sc.value


def plans_traceback(plans):
    """Traceback over the connected parts of a pipeline on one device (nw_plans_traceback)."""
    cap = plans[0].n1 + plans[0].n2 + 1
    a1, a2 = np.empty(cap, dtype=np.int8), np.empty(cap, dtype=np.int8)
    n = C.c_int()
    arr = (C.c_void_p * len(plans))(*[p._h for p in plans])
    _ck(lib().nw_plans_traceback(arr, len(plans), a1.ctypes.data, a2.ctypes.data, C.byref(n)))
    return a1[:n.value].copy(), a2[:n.value].copy()


def batch_scores(S1, S2, device=0, scoring=None):
    S1, S2 = _seq(S1), _seq(S2)
    if S1.ndim != 2 or S2.ndim != 2 or S1.shape[0] != S2.shape[0]:
        raise ValueError("S1 is npairs x len1, S2 is npairs x len2")
    out = np.empty(S1.shape[0], dtype=np.int32)
    _ck(lib().nw_cuda_batch_scores_scored(_ptr(S1), _ptr(S2), S1.shape[0], S1.shape[1], S2.shape[1], _scoring(scoring),
                                          out.ctypes.data if out.size else None, device))
    return out


def dpx_peak(device=0):
    g, mhz = C.c_double(), C.c_double()
    _ck(lib().nw_cuda_dpx_peak(device, C.byref(g), C.byref(mhz)))
    return g.value, mhz.value


def strip_partition(n1, nparts, part):
    """Column-strip partition of mpi-vert (src/mpi/mpi-vert-driver.cpp:35-36, mpi-vert.cpp:17): returns
    (start, ncols_owned) in table columns; column `start` of parts > 0 is the halo."""
    q = (n1 + 1) // nparts
    start = q * part - (1 if part > 0 else 0)
    owned = q + (1 if part > 0 else 0) + ((n1 + 1) % nparts if part == nparts - 1 else 0)
    return start, owned


class Plan:
    """Device-resident fill state (nw_plan_* of include/nw_cuda.h)."""

    def __init__(self, n1, n2, mode=NW_MODE_BOUNDARY, device=0, part=0, nparts=1, rows_per_lane=0, warps_per_cta=0,
                 ctas=0, scoring=None):
        self.n1, self.n2, self.mode, self.device, self.part, self.nparts = n1, n2, mode, device, part, nparts
        self._h = C.c_void_p()
        tune = Tuning(rows_per_lane, warps_per_cta, ctas)
        if scoring is None:
            _ck(lib().nw_plan_create(C.byref(self._h), device, n1, n2, mode, part, nparts, C.byref(tune)))
        else:
            _ck(lib().nw_plan_create_scored(C.byref(self._h), device, n1, n2, mode, part, nparts, C.byref(tune),
                                            _scoring(scoring)))
        start, owned = strip_partition(n1, nparts, part)
        self.jstart, self.ncols = start, owned - 1

    def close(self):
        if self._h:
            lib().nw_plan_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload(self, s1, s2):
        s1, s2 = _seq(s1), _seq(s2)
        if s1.size != self.n1 or s2.size != self.n2:
            raise ValueError("sequence lengths differ from the plan's")
        _ck(lib().nw_plan_upload(self._h, _ptr(s1), _ptr(s2)))

    def upload_device(self, d_s1_ptr, d_s2_ptr):
        _ck(lib().nw_plan_upload_device(self._h, d_s1_ptr, d_s2_ptr))

    def connect(self, right):
        _ck(lib().nw_plan_connect(self._h, right._h))

    def export_mailbox(self):
        buf = C.create_string_buffer(64)
        _ck(lib().nw_plan_export_mailbox(self._h, buf))
        return buf.raw

    def import_mailbox(self, handle, consumer_device):
        _ck(lib().nw_plan_import_mailbox(self._h, handle, consumer_device))

    def run(self):
        _ck(lib().nw_plan_run(self._h))

    def sync(self):
        _ck(lib().nw_plan_sync(self._h))

    def time(self, iters=1):
        ms = C.c_float()
        _ck(lib().nw_plan_time(self._h, iters, C.byref(ms)))
        return ms.value

    def timer_start(self):
        _ck(lib().nw_plan_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        _ck(lib().nw_plan_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def last_ms(self):
        ms = C.c_float()
        _ck(lib().nw_plan_last_ms(self._h, C.byref(ms)))
        return ms.value

    def launches_per_run(self):
        n = C.c_int()
        _ck(lib().nw_plan_launches_per_run(self._h, C.byref(n)))
        return n.value

    def score(self):
        out = C.c_int32()
        _ck(lib().nw_plan_score(self._h, C.byref(out)))
        return out.value

    def best(self):
        sc, i, j = C.c_int32(), C.c_int32(), C.c_int32()
        _ck(lib().nw_plan_best(self._h, C.byref(sc), C.byref(i), C.byref(j)))
        return sc.value, i.value, j.value

    def last_row(self):
        out = np.empty(self.ncols + 1, dtype=np.int32)
        _ck(lib().nw_plan_last_row(self._h, out.ctypes.data))
        return out

    def last_col(self):
        out = np.empty(self.n2 + 1, dtype=np.int32)
        _ck(lib().nw_plan_last_col(self._h, out.ctypes.data))
        return out

    def table_to_host(self, table=None):
        if table is None:
            table = np.empty((self.n2 + 1, self.n1 + 1), dtype=np.int32)
        _ck(lib().nw_plan_table_to_host(self._h, table.ctypes.data))
        return table

    def table_device(self):
        ptr, pitch = C.c_void_p(), C.c_int64()
        _ck(lib().nw_plan_table_device(self._h, C.byref(ptr), C.byref(pitch)))
        return ptr.value, pitch.value

    def traceback(self):
        """(a1, a2): the gapped s1 and s2 of the optimal alignment the table encodes, gap = 0 (README.md:8)."""
        cap = self.n1 + self.n2 + 1
        a1, a2 = np.empty(cap, dtype=np.int8), np.empty(cap, dtype=np.int8)
        n = C.c_int()
        _ck(lib().nw_plan_traceback(self._h, a1.ctypes.data, a2.ctypes.data, C.byref(n)))
        return a1[:n.value].copy(), a2[:n.value].copy()

    def strip_info(self):
        v = [C.c_int() for _ in range(5)]
        _ck(lib().nw_plan_strip_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(["nstrips", "strip_rows", "rows_per_lane", "warps", "ctas"], [x.value for x in v]))

    def strip_times(self, cycles=False):
        """(start_ns, end_ns[, sm_cycles]) per strip of the most recent fill (device %globaltimer / clock64)."""
        n = self.strip_info()["nstrips"]
        a, b, c = (np.zeros(max(n, 1), dtype=np.int64) for _ in range(3))
        _ck(lib().nw_plan_strip_times(self._h, a.ctypes.data, b.ctypes.data, c.ctypes.data))
        return (a[:n], b[:n], c[:n]) if cycles else (a[:n], b[:n])

    def strip_row(self, strip):
        out = np.empty(self.ncols + 1, dtype=np.int32)
        _ck(lib().nw_plan_strip_row(self._h, strip, out.ctypes.data))
        return out


class Batch:
    """Batch of independent pairs on one device (nw_batch_* of include/nw_cuda.h)."""

    def __init__(self, npairs, len1, len2, device=0, scoring=None):
        self.npairs, self.len1, self.len2, self.device = npairs, len1, len2, device
        self._h = C.c_void_p()
        _ck(lib().nw_batch_create(C.byref(self._h), device, npairs, len1, len2))
        if scoring is not None:
            _ck(lib().nw_batch_set_scoring(self._h, _scoring(scoring)))

    def close(self):
        if self._h:
            lib().nw_batch_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def upload(self, S1, S2):
        S1, S2 = _seq(S1), _seq(S2)
        if S1.shape != (self.npairs, self.len1) or S2.shape != (self.npairs, self.len2):
            raise ValueError("batch shape differs from the plan's")
        _ck(lib().nw_batch_upload(self._h, _ptr(S1), _ptr(S2)))

    def upload_device(self, d_S1_ptr, d_S2_ptr):
        _ck(lib().nw_batch_upload_device(self._h, d_S1_ptr, d_S2_ptr))

    def run_host(self, S1, S2, out=None, nchunks=0):
        """Host arrays in, host scores out, chunked so that copies and kernels overlap (nw_batch_run_host)."""
        S1, S2 = _seq(S1), _seq(S2)
        if S1.shape != (self.npairs, self.len1) or S2.shape != (self.npairs, self.len2):
            raise ValueError("batch shape differs from the plan's")
        if out is None:
            out = np.empty(self.npairs, dtype=np.int32)
        _ck(lib().nw_batch_run_host(self._h, _ptr(S1), _ptr(S2), out.ctypes.data if out.size else None, nchunks))
        return out

    def run(self):
        _ck(lib().nw_batch_run(self._h))

    def sync(self):
        _ck(lib().nw_batch_sync(self._h))

    def time(self, iters=1):
        ms = C.c_float()
        _ck(lib().nw_batch_time(self._h, iters, C.byref(ms)))
        return ms.value

    def scores(self):
        out = np.empty(self.npairs, dtype=np.int32)
        _ck(lib().nw_batch_scores(self._h, out.ctypes.data if out.size else None))
        return out
