#!/usr/bin/env python3
"""Size sweep in the reference's own TSV format (mirrors src/benchmark-sizes.sh:40-62; the table it writes has the
layout of data/multi.tsv, which data/graph.py:33-50 parses: a title line, a header `program\\t2gb\\t4gb...`, then one row
of integer wall-milliseconds per program).

Every program is run through its driver binary and the FIRST stdout token -- the driver's own wall-ms
(src/common/driver.cpp:30-33) -- is what goes into the table, exactly like the zsh script's `let "a = $(./prog a b)"`.
Programs: `cuda` (this repo's bin/cuda.e, NW_CUDA_MODE=boundary), `cuda-full` (NW_CUDA_MODE=full, tables up to
--max-full-gb), and any of the reference binaries compiled into oracle/_ref/ (serial, sentinel-otf-blocked-mt, ...).

    python tools/benchmark_sizes.py --min 2 --max 64 --step 2 --runs 3 --programs cuda,cuda-full,serial -o sizes.tsv
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
BDNA = os.path.join(REF, "bdna")


def run_ms(exe, a, b, env):
    out = subprocess.run([exe, a, b], capture_output=True, text=True, env=env)
    if out.returncode != 0:
        raise RuntimeError(f"{exe} failed: {out.stdout} {out.stderr}")
    return int(out.stdout.split()[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min", type=int, default=2)
    ap.add_argument("--max", type=int, default=16)
    ap.add_argument("--step", type=int, default=2)
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--programs", default="cuda,cuda-full")
    ap.add_argument("--max-full-gb", type=int, default=16)
    ap.add_argument("-o", "--out", default="sizes.tsv")
    args = ap.parse_args()
    if os.path.exists(args.out):
        sys.exit(f"WARNING! {args.out} already exists. please rename or remove.")    # benchmark-sizes.sh:34-38
    sizes = list(range(args.min, args.max + 1, args.step))
    rows = {}
    out = open(args.out, "w")                                  # rows are written as they complete
    out.write("benchmarking " + ", ".join(args.programs.split(",")) + "\n")
    out.write("program\t" + "".join(f"{g}gb\t" for g in sizes) + "\n")
    out.flush()
    for prog in args.programs.split(","):
        env = dict(os.environ, OMP_NUM_THREADS=str(args.threads))
        if prog == "cuda":
            exe, env["NW_CUDA_MODE"] = os.path.join(ROOT, "fast-needleman-wunsch_b200", "bin", "cuda.e"), "boundary"
        elif prog == "cuda-full":
            exe, env["NW_CUDA_MODE"] = os.path.join(ROOT, "fast-needleman-wunsch_b200", "bin", "cuda.e"), "full"
        else:
            exe = os.path.join(REF, prog + ".e")
        if not os.path.exists(exe):
            sys.exit(f"missing {exe}")
        row = []
        for g in sizes:
            if prog == "cuda-full" and g > args.max_full_gb:
                row.append(0)
                continue
            a, b = os.path.join(BDNA, f"{g}gb-1.bdna"), os.path.join(BDNA, f"{g}gb-2.bdna")
            print(f"running {prog} on {g}gb...", file=sys.stderr)
            row.append(sum(run_ms(exe, a, b, env) for _ in range(args.runs)) // args.runs)
        rows[prog] = row
        out.write(prog + "\t" + "".join(f"{v}\t" for v in row) + "\n")
        out.flush()
    out.close()
    print(open(args.out).read())


if __name__ == "__main__":
    main()
