#!/usr/bin/env python3
"""Generate tests/golden/golden_scoring.json: known answers for OTHER scoring triples, from the reference itself.

The reference's only scoring configuration is the three macros MATCH / MISMATCH / GAP of
src/common/needleman-wunsch.hpp:11-13.  For every triple below this script writes a copy of that header into a
temporary directory with ONLY those three #define lines edited, compiles the reference's unmodified
src/serial/serial.cpp (+ src/common/helper.cpp) against it exactly like oracle/Makefile builds libnw_ref.so (outputs
only, under oracle/_ref/scoring/, git-ignored), fills seeded synthetic pairs and the small fixtures with it and records
score / sum / min / max / FNV-1a-64 of the tables.  tests/test_oracle.py pins oracle/nw_oracle.c's *_ex functions to
these vectors; the GPU tests compare the CUDA path with the oracle.

Usage:  python tests/golden/make_golden_scoring.py        (build container only: needs /root/reference)
"""
import ctypes, json, os, re, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import pair_paths, synth_pair, table_facts   # noqa: E402

REFSRC = os.environ.get("NW_REF", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref", "scoring")
# (match, mismatch, gap): the reference's own triple as a control, then common DNA scorings, a triple whose mismatch is
# worse than two gaps (weight clamps to 0 in the G form), one with a large match (leaves the packed kernels) and one
# with gap -1 but other substitution scores (keeps the packed full-table path)
TRIPLES = [(1, 0, -1), (1, -1, -2), (2, -1, -2), (5, -4, -3), (1, -3, -1), (2, -3, -5), (20, -7, -4), (3, 1, -1)]
# (name, seed, n1, n2, alphabet_hi)
PAIRS = [("syn_0x5", 21, 0, 5, 5), ("syn_5x0", 22, 5, 0, 5), ("syn_1x1", 23, 1, 1, 5), ("syn_33x31", 24, 33, 31, 5),
         ("syn_257x511", 25, 257, 511, 5), ("syn_1500x900", 26, 1500, 900, 5), ("syn_700x2100", 27, 700, 2100, 5),
         ("syn_1200x1100_bytes", 28, 1200, 1100, 100), ("syn_600x600_ident", 29, 600, 600, 2)]
FIXTURE_PAIRS = ["small", "t", "debug"]


def build_variant(match, mismatch, gap):
    tag = f"m{match}_x{mismatch}_g{gap}".replace("-", "n")
    so = os.path.join(OUT, f"libnw_ref_{tag}.so")
    os.makedirs(OUT, exist_ok=True)
    common = os.path.join(REFSRC, "src", "common")
    with tempfile.TemporaryDirectory() as tmp:
        hdr = open(os.path.join(common, "needleman-wunsch.hpp")).read()
        for name, val in (("MATCH", match), ("MISMATCH", mismatch), ("GAP", gap)):
            hdr, n = re.subn(rf"^#define {name} .*$", f"#define {name} {val}", hdr, flags=re.M)
            assert n == 1, name
        open(os.path.join(tmp, "needleman-wunsch.hpp"), "w").write(hdr)
        flags = ["-Wall", "-std=c++11", "-O3", "-I", tmp, "-I", common, "-fPIC"]
        subprocess.run(["g++", *flags, "-Wno-mismatched-new-delete", "-Dmain=nw_ref_main", "-DneedlemanWunsch=nw_ref_serial_impl",
                        "-c", os.path.join(REFSRC, "src", "serial", "serial.cpp"), "-o", os.path.join(tmp, "serial.o")], check=True)
        subprocess.run(["g++", *flags, "-c", os.path.join(common, "helper.cpp"), "-o", os.path.join(tmp, "helper.o")], check=True)
        subprocess.run(["g++", *flags, "-shared", "-o", so, os.path.join(ROOT, "oracle", "ref_shim.cpp"),
                        os.path.join(tmp, "serial.o"), os.path.join(tmp, "helper.o")], check=True)
    return so


def ref_table(so, s1, s2):
    ref = ctypes.CDLL(so)
    ref.nw_ref_serial_fill.restype = None
    ref.nw_ref_serial_fill.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    a = np.ascontiguousarray(s1 if s1.size else np.zeros(1, np.int8))
    b = np.ascontiguousarray(s2 if s2.size else np.zeros(1, np.int8))
    t = np.empty((s2.size + 1, s1.size + 1), dtype=np.int32)
    ref.nw_ref_serial_fill(a.ctypes.data, s1.size, b.ctypes.data, s2.size, t.ctypes.data)
    return t


def main():
    golden = {"generator": "tests/golden/make_golden_scoring.py",
              "reference_impl": "src/serial/serial.cpp (unmodified) against needleman-wunsch.hpp with the MATCH / MISMATCH / GAP "
                                "#defines edited", "triples": []}
    for (m, x, g) in TRIPLES:
        so = build_variant(m, x, g)
        entry = {"match": m, "mismatch": x, "gap": g, "synthetic": {}, "fixtures": {}}
        for name, seed, n1, n2, hi in PAIRS:
            s1, s2 = synth_pair(seed, n1, n2, hi)
            facts = table_facts(ref_table(so, s1, s2))
            facts.update({"seed": seed, "n1": n1, "n2": n2, "alphabet_hi": hi})
            entry["synthetic"][name] = facts
        for name in FIXTURE_PAIRS:
            a, b = pair_paths(name)
            entry["fixtures"][name] = table_facts(ref_table(so, np.fromfile(a, dtype=np.int8), np.fromfile(b, dtype=np.int8)))
        golden["triples"].append(entry)
        print((m, x, g), {k: v["score"] for k, v in entry["synthetic"].items()}, flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "golden_scoring.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
