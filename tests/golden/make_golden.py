#!/usr/bin/env python3
"""Generate tests/golden/golden.json from the REFERENCE ITSELF (run in the build container only).

Sources of truth, in order:
  * oracle/_ref/serial.e        -- the reference's src/serial/serial.cpp + src/common/driver.cpp, compiled
                                   unmodified by oracle/Makefile; gives the `Score:` line (driver.cpp:35)
  * oracle/_ref/libnw_ref.so    -- the same serial.cpp as a library (main renamed); gives full tables, from which
                                   the sum / min / max / FNV-1a-64 known answers are taken
  * oracle/liboracle.so         -- our restatement, used ONLY for pairs whose int32 table does not fit this
                                   container's RAM (marked source="two-row restatement"), after it matched
                                   serial.e on every pair that does fit.

Usage:  python tests/golden/make_golden.py [--max-table-gb 45] [--jobs 6]
/root/reference is needed (through oracle/_ref); the GPU box never runs this script.
"""
import argparse, ctypes, json, os, re, subprocess, sys
from concurrent.futures import ProcessPoolExecutor
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
BDNA = os.path.join(REF, "bdna")

FIXTURES = ["small", "t", "debug", "smid", "mid", "big"] + [f"{k}gb" for k in range(2, 65, 2)]
FULL_TABLE = ["small", "t", "debug", "smid", "2gb"]
# (name, seed, n1, n2, alphabet_hi): bases iid uniform on 1..alphabet_hi-1, s1 drawn first then s2
SYNTH = [
    ("syn_0x0", 1, 0, 0, 5), ("syn_0x7", 2, 0, 7, 5), ("syn_7x0", 3, 7, 0, 5), ("syn_1x1", 4, 1, 1, 5),
    ("syn_31x31", 5, 31, 31, 5), ("syn_32x32", 6, 32, 32, 5), ("syn_33x33", 7, 33, 33, 5),
    ("syn_1x300", 8, 1, 300, 5), ("syn_300x1", 9, 300, 1, 5),
    ("syn_257x511", 10, 257, 511, 5), ("syn_1000x1000", 11, 1000, 1000, 5),
    ("syn_4097x129", 12, 4097, 129, 5), ("syn_129x4097", 13, 129, 4097, 5),
    ("syn_5000x3000", 14, 5000, 3000, 5), ("syn_2048x2048_ident", 15, 2048, 2048, 2),
    ("syn_3001x2999_bytes", 16, 3001, 2999, 120),
]


def pair_paths(name):
    if name.endswith("gb"):
        return os.path.join(BDNA, f"{name}-1.bdna"), os.path.join(BDNA, f"{name}-2.bdna")
    return os.path.join(BDNA, f"{name}1.bdna"), os.path.join(BDNA, f"{name}2.bdna")


def synth_pair(seed, n1, n2, hi):
    rng = np.random.default_rng(seed)
    s1 = rng.integers(1, hi, size=n1, dtype=np.int8)
    s2 = rng.integers(1, hi, size=n2, dtype=np.int8)
    return s1, s2


def run_serial(name):
    a, b = pair_paths(name)
    out = subprocess.run([os.path.join(REF, "serial.e"), a, b], capture_output=True, text=True, check=True).stdout
    return int(re.search(r"Score:\s*(-?\d+)", out).group(1))


def oracle_lib():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    lib.nw_oracle_score.restype = ctypes.c_int32
    lib.nw_oracle_score.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32]
    lib.nw_oracle_fnv1a64.restype = ctypes.c_uint64
    lib.nw_oracle_fnv1a64.argtypes = [ctypes.c_void_p, ctypes.c_int64]
    lib.nw_oracle_fnv1a64_col.restype = ctypes.c_uint64
    lib.nw_oracle_fnv1a64_col.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
    return lib


def two_row_score(name):
    a, b = pair_paths(name)
    s1 = np.fromfile(a, dtype=np.int8); s2 = np.fromfile(b, dtype=np.int8)
    lib = oracle_lib()
    return int(lib.nw_oracle_score(s1.ctypes.data, s1.size, s2.ctypes.data, s2.size))


def ref_table(s1, s2):
    ref = ctypes.CDLL(os.path.join(REF, "libnw_ref.so"))
    ref.nw_ref_serial_fill.restype = None
    ref.nw_ref_serial_fill.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    # give zero-length inputs a valid pointer
    a = np.ascontiguousarray(s1 if s1.size else np.zeros(1, np.int8))
    b = np.ascontiguousarray(s2 if s2.size else np.zeros(1, np.int8))
    t = np.empty((s2.size + 1, s1.size + 1), dtype=np.int32)
    ref.nw_ref_serial_fill(a.ctypes.data, s1.size, b.ctypes.data, s2.size, t.ctypes.data)
    return t


def table_facts(t):
    lib = oracle_lib()
    nrows, ncols = t.shape
    last_row = np.ascontiguousarray(t[-1])
    return {
        "rows": int(nrows), "cols": int(ncols), "score": int(t[-1, -1]),
        "sum": int(t.sum(dtype=np.int64)), "min": int(t.min()), "max": int(t.max()),
        "fnv_table": f"{lib.nw_oracle_fnv1a64(t.ctypes.data, t.nbytes):016x}",
        "fnv_lastrow": f"{lib.nw_oracle_fnv1a64(last_row.ctypes.data, last_row.nbytes):016x}",
        "fnv_lastcol": f"{lib.nw_oracle_fnv1a64_col(t.ctypes.data + 4 * (ncols - 1), nrows, ncols):016x}",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-table-gb", type=float, default=45.0)
    ap.add_argument("--jobs", type=int, default=6)
    args = ap.parse_args()

    golden = {"generator": "tests/golden/make_golden.py", "reference_impl": "src/serial/serial.cpp (unmodified)",
              "fixtures": {}, "tables": {}, "synthetic": {}}
    sizes = {}
    for name in FIXTURES:
        a, b = pair_paths(name)
        sizes[name] = (os.path.getsize(a), os.path.getsize(b))

    with ProcessPoolExecutor(args.jobs) as ex:
        two_row = dict(zip(FIXTURES, ex.map(two_row_score, FIXTURES)))
    for name in FIXTURES:
        n1, n2 = sizes[name]
        gb = (n1 + 1) * (n2 + 1) * 4 / 1e9
        entry = {"n1": n1, "n2": n2, "score": two_row[name], "source": "two-row restatement (oracle/nw_oracle.c)"}
        if gb <= args.max_table_gb:
            s = run_serial(name)
            assert s == two_row[name], (name, s, two_row[name])
            entry["source"] = "reference serial.e; two-row restatement agrees"
        golden["fixtures"][name] = entry
        print(name, entry, flush=True)

    for name in FULL_TABLE:
        a, b = pair_paths(name)
        t = ref_table(np.fromfile(a, dtype=np.int8), np.fromfile(b, dtype=np.int8))
        golden["tables"][name] = table_facts(t)
        assert golden["tables"][name]["score"] == golden["fixtures"][name]["score"]
        print(name, golden["tables"][name], flush=True)

    for name, seed, n1, n2, hi in SYNTH:
        s1, s2 = synth_pair(seed, n1, n2, hi)
        facts = table_facts(ref_table(s1, s2))
        facts.update({"seed": seed, "n1": n1, "n2": n2, "alphabet_hi": hi})
        golden["synthetic"][name] = facts
        print(name, facts, flush=True)

    with open(os.path.join(ROOT, "tests", "golden", "golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    print("wrote golden.json")


if __name__ == "__main__":
    sys.exit(main())
