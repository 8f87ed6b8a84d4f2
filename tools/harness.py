#!/usr/bin/env python3
"""Sweep harness in the reference's own TSV formats, so that its plotting scripts read GPU and CPU runs alike.

Mirrors the reference's zsh scripts (which are not runnable here: no zsh, no MPI):
  sizes   src/benchmark-sizes.sh:40-62     rows = programs, columns = fixture sizes (2gb .. 64gb step 2)
  gpus    src/benchmark-threads.sh:63-103  rows = `serial`, then one row per GPU count (where the reference sweeps
                                           OMP_NUM_THREADS, this sweeps NW_CUDA_GPUS = column strips over devices)
  tune    src/buf-tune.sh:24-51            one file per size: `<size>gb` / `bufsize\\t v...` / `time\\t ms...`; where the reference
                                           sweeps the MPI message size, this sweeps one knob of the CUDA fill:
                                           rows_per_lane (NW_CUDA_R), tile_blocks (NW_CUDA_TILE_BLOCKS), band_mb (NW_CUDA_BAND_MB)
Every number is the FIRST stdout token of a driver binary -- the driver's own integer wall-ms (src/common/driver.cpp:30-33)
-- exactly like the scripts' `let "a = $(./prog a b)"`, averaged over --runs.  parse_tsv() reads the table back with the
column logic of data/graph.py:33-50.

Programs: `cuda` (bin/cuda.e, NW_CUDA_MODE=boundary), `cuda-full` (NW_CUDA_MODE=full), or any reference binary compiled into
oracle/_ref/ (serial, sentinel-otf-blocked-mt, idxarray-mod-mt).

    python tools/harness.py sizes --min 2 --max 64 --step 2 --runs 3 --programs cuda,cuda-full,serial -o sizes.tsv
    python tools/harness.py gpus  --min 2 --max 16 --step 2 --gpu-counts 1,2,4,8 --mode full -o gpus.tsv
    python tools/harness.py tune  --sizes 2,8 --knob rows_per_lane --values 4,8,16 --mode full --prefix rtune
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
KNOBS = {"rows_per_lane": "NW_CUDA_R", "tile_blocks": "NW_CUDA_TILE_BLOCKS", "band_mb": "NW_CUDA_BAND_MB"}


def run_ms(exe, a, b, env):
    out = subprocess.run([exe, a, b], capture_output=True, text=True, env=env)
    if out.returncode != 0:
        raise RuntimeError(f"{exe} failed ({out.returncode}): {out.stdout} {out.stderr}")
    return int(out.stdout.split()[0])


def mean_ms(exe, a, b, env, runs):
    return sum(run_ms(exe, a, b, env) for _ in range(runs)) // runs        # integer mean, like `let "s = s / $nRuns"`


def program(prog, args, extra_env=None):
    """(executable, environment) of a program name."""
    env = dict(os.environ, OMP_NUM_THREADS=str(args.threads))
    if prog in ("cuda", "cuda-full"):
        exe = os.path.join(args.exe_dir or os.path.join(ROOT, "fast-needleman-wunsch_b200", "bin"), "cuda.e")
        env["NW_CUDA_MODE"] = "full" if prog == "cuda-full" else "boundary"
    else:
        exe = os.path.join(args.exe_dir or REF, prog + ".e")
    env.update(extra_env or {})
    if not os.path.exists(exe):
        sys.exit(f"missing {exe}")
    return exe, env


def pair(args, g):
    return os.path.join(args.bdna, f"{g}gb-1.bdna"), os.path.join(args.bdna, f"{g}gb-2.bdna")


def refuse_overwrite(path):
    if os.path.exists(path):
        sys.exit(f"WARNING! {path} already exists. please rename or remove.")        # benchmark-sizes.sh:34-38


def write_table(path, title, sizes, rows_iter):
    """Title line, header `program\\t2gb\\t...`, one row per program / count; rows are written as they complete."""
    refuse_overwrite(path)
    with open(path, "w") as out:
        out.write(title + "\n")
        out.write("program\t" + "".join(f"{g}gb\t" for g in sizes) + "\n")
        out.flush()
        for name, row in rows_iter:
            out.write(str(name) + "\t" + "".join(f"{v}\t" for v in row) + "\n")
            out.flush()


def cmd_sizes(args):
    sizes = list(range(args.min, args.max + 1, args.step))
    progs = args.programs.split(",")

    def rows():
        for prog in progs:
            exe, env = program(prog, args)
            row = []
            for g in sizes:
                if prog == "cuda-full" and g > args.max_full_gb:
                    row.append(0)
                    continue
                print(f"running {prog} on {g}gb...", file=sys.stderr)
                row.append(mean_ms(exe, *pair(args, g), env, args.runs))
            yield prog, row
    write_table(args.out, "benchmarking " + ", ".join(progs), sizes, rows())


def cmd_gpus(args):
    sizes = list(range(args.min, args.max + 1, args.step))
    counts = [int(x) for x in args.gpu_counts.split(",")]
    prog = "cuda-full" if args.mode == "full" else "cuda"

    def rows():
        if not args.no_serial:
            exe, env = program("serial", args)
            print("benchmarking serial...", file=sys.stderr)
            yield "serial", [mean_ms(exe, *pair(args, g), env, args.runs) for g in sizes]
        for n in counts:
            print(f"benchmarking {n} GPUs...", file=sys.stderr)
            exe, env = program(prog, args, {"NW_CUDA_GPUS": str(n)})
            yield n, [mean_ms(exe, *pair(args, g), env, args.runs) for g in sizes]
    write_table(args.out, f"benchmarking {prog} over GPU counts " + ", ".join(map(str, counts)), sizes, rows())


def cmd_tune(args):
    if args.knob not in KNOBS:
        sys.exit(f"unknown knob {args.knob}: one of {sorted(KNOBS)}")
    values = [int(x) for x in args.values.split(",")]
    prog = "cuda-full" if args.mode == "full" else "cuda"
    for g in (int(x) for x in args.sizes.split(",")):
        path = f"{args.prefix}{g}.tsv"
        refuse_overwrite(path)
        print(f"benchmarking {g}gb", file=sys.stderr)
        with open(path, "w") as out:                               # buf-tune.sh:28-49
            out.write(f"{g}gb\n")
            out.write("bufsize\t" + "".join(f"{v}\t" for v in values) + "\n")
            out.write("time\t")
            for v in values:
                exe, env = program(prog, args, {KNOBS[args.knob]: str(v)})
                out.write(f"{mean_ms(exe, *pair(args, g), env, args.runs)}\t")
                out.flush()
            out.write("\n")


def parse_tsv(path):
    """(title, x values, {row name: values}) with the column logic of the reference's data/graph.py:33-50."""
    with open(path) as fin:
        title = fin.readline().rstrip("\n")
        x = [int(a[:-2]) for a in fin.readline().split()[1:]]        # "2gb" -> 2
        y = {}
        for line in fin:
            sp = line.split()
            if sp:
                y[sp[0]] = [int(a) for a in sp[1:]]
    return title, x, y


def parse_tune_tsv(path):
    """(size label, knob values, times) of a tuning file (layout of src/buf-tune.sh:28-49 / data/buf-tuning/*.tsv)."""
    with open(path) as fin:
        size = fin.readline().strip()
        vals = [int(a) for a in fin.readline().split()[1:]]
        times = [int(a) for a in fin.readline().split()[1:]]
    return size, vals, times


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1, help="OMP_NUM_THREADS of reference programs")
    ap.add_argument("--bdna", default=os.path.join(REF, "bdna"), help="directory of the <N>gb-1/2.bdna fixtures")
    ap.add_argument("--exe-dir", default=None, help="where the driver binaries live (default: bin/ and oracle/_ref/)")
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("sizes")
    s.add_argument("--min", type=int, default=2)
    s.add_argument("--max", type=int, default=16)
    s.add_argument("--step", type=int, default=2)
    s.add_argument("--programs", default="cuda,cuda-full")
    s.add_argument("--max-full-gb", type=int, default=16)
    s.add_argument("-o", "--out", default="sizes.tsv")
    s.set_defaults(fn=cmd_sizes)
    g = sub.add_parser("gpus")
    g.add_argument("--min", type=int, default=2)
    g.add_argument("--max", type=int, default=16)
    g.add_argument("--step", type=int, default=2)
    g.add_argument("--gpu-counts", default="1,2,4,8")
    g.add_argument("--mode", default="full", choices=["full", "boundary"])
    g.add_argument("--no-serial", action="store_true")
    g.add_argument("-o", "--out", default="threads.tsv")
    g.set_defaults(fn=cmd_gpus)
    t = sub.add_parser("tune")
    t.add_argument("--sizes", default="2")
    t.add_argument("--knob", default="rows_per_lane")
    t.add_argument("--values", default="4,8,16")
    t.add_argument("--mode", default="full", choices=["full", "boundary"])
    t.add_argument("--prefix", default="buftune")
    t.set_defaults(fn=cmd_tune)
    args = ap.parse_args()
    args.fn(args)


if __name__ == "__main__":
    main()
