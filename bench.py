#!/usr/bin/env python3
"""bench.py -- GCUPS of the single-pair Needleman-Wunsch fill (BASELINE.json metric) on 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 64gb|big|mid|2gb|2gb-full|smid|batch] [--impl reference]

A "step" is ONE complete fill of the pair's scoring table.  The default workload is the reference's 64gb-1/64gb-2
fixture pair (BASELINE.json configs[3]; 126 440 x 127 240 = 16.09 G cells) in boundary-only memory mode, which fits one
GPU, so the same job is timed at N = 1, 2, 4, 8 (strong scaling: column strips over the GPUs, NVLink handoff of the
boundary column; reference decomposition: src/mpi/mpi-vert.cpp:17-105).  For N > 1 the driver launches one process per
GPU with torchrun; torch.distributed is only plumbing (rendezvous, IPC handle exchange, barrier, max-reduce).

What the JSON line (rank 0) says, and how it is measured:
  value / ms_per_step   one FORWARD fill that keeps every strip boundary row, the last row and the last column
                        (NW_MODE_BOUNDARY plan), sequences resident in HBM.  Every step is timed on its own with CUDA
                        events on the plan's stream (nw_plan_last_ms); a step ends with a stream sync (and, for N > 1, a
                        barrier), so fills never overlap: ms_per_step is the LATENCY of one fill, the mean over the K
                        steps of the max over ranks.  (Back-to-back fills of a pipeline overlap; that throughput is
                        reported separately as pipelined_throughput.)
  score_only            kernel-only time of the score-only mode (NW_MODE_SCORE: top half forwards, bottom half
                        backwards, concurrently) -- the algorithm the plug-in call takes in boundary mode.
  e2e                   through the reference-facing one-shot C-ABI call, nw_cuda_fill_ex(host s1, host s2, host table,
                        NW_MODE_BOUNDARY, N): pageable host sequences in, H2D, fill, score written to table[size-1],
                        every step (N = 1: score-only mode inside; N > 1: column strips in ONE process are not what
                        torchrun launches, so e2e there is upload + run + score through the plan API).  e2e_cold: the
                        reference's UNCHANGED driver around the same call in a fresh process (cuda.e), i.e. what
                        src/common/driver.cpp:26-33 prints, several runs.
  roofline              the dominant kernel against the integer/DPX pipe rate MEASURED on this GPU (nw_cuda_dpx_peak),
                        3 lane-ops per cell; traffic from the committed ncu capture.
  configs               the other BASELINE.json configurations (mid, big, 2gb boundary / full table, batch), a few
                        timed fills each, with the reference's CPU numbers for the same WHOLE pair where it was run.
  cpu_baseline          the reference's own serial.cpp and OpenMP variants (compiled unmodified into oracle/_ref/),
                        timed on this box's host cores on whole fixture pairs (2gb always; mid when RAM allows).

--impl reference times the reference's fastest CPU variant with all host threads on a bounded sample (a prefix, as large
as RAM and a few minutes allow) of the same pair.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")
BDNA = os.path.join(REF, "bdna")
CUDA_E = os.path.join(ROOT, "fast-needleman-wunsch_b200", "bin", "cuda.e")
METRIC = "GCUPS single-pair NW fill"
GOLDEN_SCORES = {"64gb": 73888, "big": 58529, "mid": 29249, "2gb": 12958, "smid": 5839}
SHAPES = {"64gb": (126440, 127240), "big": (100063, 99977), "mid": (49902, 49555), "2gb": (22541, 22116),
          "smid": (10030, 9976)}


def pair_paths(name):
    sep = "-" if name.endswith("gb") else ""
    return os.path.join(BDNA, f"{name}{sep}1.bdna"), os.path.join(BDNA, f"{name}{sep}2.bdna")


def load_pair(name):
    """The reference's bdna fixture when it was staged next to the compiled reference; else seeded synthetic bases of
    the same lengths (iid uniform on 1..4, like the fixtures)."""
    a, b = pair_paths(name)
    if os.path.exists(a) and os.path.exists(b):
        return np.fromfile(a, dtype=np.int8), np.fromfile(b, dtype=np.int8), f"reference bdna fixture {name} (1 byte per base)"
    n1, n2 = SHAPES[name]
    rng = np.random.default_rng(20240607)
    return (rng.integers(1, 5, size=n1, dtype=np.int8), rng.integers(1, 5, size=n2, dtype=np.int8),
            f"synthetic iid bases, lengths of the {name} fixture")


def workload_text(name, n1, n2, full):
    return f"{name} pair ({n1} x {n2} = {n1 * n2} cells), {'full-table' if full else 'boundary-only'} mode"


def mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 2 ** 20
    except Exception:
        pass
    return 0.0


# ---------------------------------------------------------------------------------------------------------------------
# clocks during the timed region (NVML)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._t = [], set(), None, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs: the reference compiled unmodified into oracle/_ref/ (the only place bench.py executes anything of oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def run_ref_binary(exe, a, b, threads, timeout=900):
    """One run of a reference binary; (driver-printed ms, Score) as src/common/driver.cpp:33-35 prints them."""
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false")
    try:
        out = subprocess.run([os.path.join(REF, exe), a, b], capture_output=True, text=True, env=env, timeout=timeout)
    except subprocess.TimeoutExpired:
        return None, None
    if out.returncode != 0:
        return None, None
    m = re.match(r"\s*(\d+)\s*\nScore:\s*(-?\d+)", out.stdout)
    return (int(m.group(1)), int(m.group(2))) if m else (None, None)


def write_prefix_pair(s1, s2, n):
    d = tempfile.mkdtemp(prefix="nw_bench_")
    a, b = os.path.join(d, "a.bdna"), os.path.join(d, "b.bdna")
    s1[:n].tofile(a)
    s2[:n].tofile(b)
    return a, b, min(n, s1.size), min(n, s2.size)


def have_reference_binaries():
    return all(os.path.exists(os.path.join(REF, e)) for e in ("serial.e", "idxarray-mod-mt.e"))


def oracle_port_gcups(s1, s2, n):
    """Fallback when the compiled reference did not travel: the C restatement of the same loop, one core."""
    lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    lib.nw_oracle_score.restype = C.c_int32
    lib.nw_oracle_score.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    a, b = np.ascontiguousarray(s1[:n]), np.ascontiguousarray(s2[:n])
    t0 = time.perf_counter()
    lib.nw_oracle_score(a.ctypes.data, a.size, b.ctypes.data, b.size)
    dt = time.perf_counter() - t0
    return a.size * b.size / dt / 1e9, a.size, b.size


def cpu_pair(name, cores):
    """serial + the two OpenMP variants on one WHOLE fixture pair; driver-printed ms."""
    a, b = pair_paths(name)
    n1, n2 = SHAPES[name]
    cells = n1 * n2
    out = {"pair": name, "cells": cells, "golden_score": GOLDEN_SCORES[name]}
    ms, sc = run_ref_binary("serial.e", a, b, 1)
    if ms is None:
        return None
    out["serial"] = {"gcups": cells / max(ms, 1) / 1e6, "ms": ms, "threads": 1, "score_ok": sc == GOLDEN_SCORES[name]}
    for exe in ("idxarray-mod-mt.e", "sentinel-otf-blocked-mt.e"):
        if not os.path.exists(os.path.join(REF, exe)):
            continue
        best = None
        for th in sorted({min(8, cores), cores}):
            ms_t, sc_t = run_ref_binary(exe, a, b, th)
            if ms_t is not None and (best is None or ms_t < best[0]):
                best = (ms_t, th, sc_t)
        if best:
            out[exe[:-2]] = {"gcups": cells / max(best[0], 1) / 1e6, "ms": best[0], "threads": best[1],
                             "score_ok": best[2] == GOLDEN_SCORES[name]}
    return out


def cpu_baseline(s1, s2, workload):
    cores = os.cpu_count() or 1
    if not (have_reference_binaries() and os.path.exists(pair_paths("2gb")[0])):
        g, m1, m2 = oracle_port_gcups(s1, s2, 40000)
        return {"value": g, "unit": "GCUPS", "cores": 1, "kind": "port",
                "sample": f"oracle/nw_oracle.c two-row restatement on the {m1}x{m2} prefix of {workload} (compiled reference absent)"}
    pairs = {}
    p = cpu_pair("2gb", cores)
    if p:
        pairs["2gb"] = p
    # mid: 9.9 GB table per run; idxarray-mod-mt needs ~7 s per run there, so only serial + the blocked variant
    if mem_available_gb() > 24:
        a, b = pair_paths("mid")
        n1, n2 = SHAPES["mid"]
        cells = n1 * n2
        q = {"pair": "mid", "cells": cells, "golden_score": GOLDEN_SCORES["mid"]}
        ms, sc = run_ref_binary("serial.e", a, b, 1)
        if ms is not None:
            q["serial"] = {"gcups": cells / max(ms, 1) / 1e6, "ms": ms, "threads": 1, "score_ok": sc == GOLDEN_SCORES["mid"]}
        if os.path.exists(os.path.join(REF, "sentinel-otf-blocked-mt.e")):
            ms, sc = run_ref_binary("sentinel-otf-blocked-mt.e", a, b, cores)
            if ms is not None:
                q["sentinel-otf-blocked-mt"] = {"gcups": cells / max(ms, 1) / 1e6, "ms": ms, "threads": cores,
                                                "score_ok": sc == GOLDEN_SCORES["mid"]}
        pairs["mid"] = q
    s = pairs.get("2gb", {}).get("serial")
    if not s:
        return None
    return {"value": s["gcups"], "unit": "GCUPS", "cores": 1, "kind": "reference", "host_cores": cores,
            "sample": "reference serial.e (src/serial/serial.cpp, unmodified) on the WHOLE 2gb fixture pair "
                      f"(22541 x 22116), driver-printed {s['ms']} ms; per-pair detail (serial, idxarray-mod-mt, "
                      "sentinel-otf-blocked-mt; 2gb and, RAM permitting, mid) under 'pairs'; the GPU numbers for the "
                      "same pairs are in the line's 'configs'",
            "pairs": pairs}


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation, all host threads, bounded sample of the same pair."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload.replace("-full", "")
    if name == "batch":
        name = "64gb"
    full = args.workload.endswith("-full")
    s1, s2, data = load_pair(name)
    cores = os.cpu_count() or 1
    # the sample: a prefix of the pair whose int32 table fits a fraction of the free RAM and whose fill takes well under
    # a second with the fastest variant (whole pair when it is that small)
    budget_gb = max(2.0, min(12.0, mem_available_gb() / 5))
    n = int(min(max(s1.size, s2.size), (budget_gb * 2 ** 30 / 4) ** 0.5))
    a, b, m1, m2 = write_prefix_pair(s1, s2, n)
    cells = m1 * m2
    whole = m1 == s1.size and m2 == s2.size
    if have_reference_binaries():
        kind = "reference"
        cands = [("serial.e", 1)]
        for exe in ("sentinel-otf-blocked-mt.e", "idxarray-mod-mt.e"):
            if os.path.exists(os.path.join(REF, exe)):
                cands.append((exe, cores))
        # untimed: pick the fastest variant on this host (the reference does not say which is its best); on a small prefix
        ta, tb, _, _ = write_prefix_pair(s1, s2, min(n, 12000))
        trial = []
        for exe, th in cands:
            ms, _ = run_ref_binary(exe, ta, tb, th)
            if ms is not None:
                trial.append((ms, exe, th))
        trial.sort()
        _, exe, th = trial[0]

        def step():
            ms, _ = run_ref_binary(exe, a, b, th)
            return ms / 1e3
        impl = f"{exe[:-2]} ({th} thread{'s' if th > 1 else ''}; fastest of {[e[:-2] for _, e, _ in trial]} on this host)"
    else:
        kind, th = "port", 1
        lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        lib.nw_oracle_score.restype = C.c_int32
        lib.nw_oracle_score.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        x, y = np.ascontiguousarray(s1[:n]), np.ascontiguousarray(s2[:n])

        def step():
            t0 = time.perf_counter()
            lib.nw_oracle_score(x.ctypes.data, x.size, y.ctypes.data, y.size)
            return time.perf_counter() - t0
        impl = "oracle/nw_oracle.c two-row restatement (reference binaries absent)"
    for _ in range(min(args.warmup, 2)):
        step()
    secs = [step() for _ in range(args.steps)]
    total = sum(secs)
    value = cells * args.steps / total / 1e9
    sample = (f"{impl} on {'the WHOLE pair' if whole else f'the {m1}x{m2} prefix'} of {name} "
              f"({4 * (m1 + 1) * (m2 + 1) / 2 ** 30:.1f} GB table per run, as the reference's driver allocates it); "
              "time = the reference driver's own printed wall ms (src/common/driver.cpp:26-33)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": data,
            "config": {"workload": workload_text(name, s1.size, s2.size, full), "sample_cells_per_step": cells,
                       "sample_is_whole_pair": whole},
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": th, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def ncu_traffic(name, full):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of the same workload, else None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t[name + ("-full" if full else "")]["bytes"]
    except Exception:
        return None


def plugin_boundary_call(nw, s1, s2, ngpus=1):
    """nw_cuda_fill_ex(s1, s2, table, NW_MODE_BOUNDARY, ngpus) exactly as the reference-side binding calls it
    (csrc/cuda.cpp), without owning the (n1+1)(n2+1)-int table the driver allocates: boundary mode writes only
    table[size-1] (what driver.cpp:35 reads), so `table` is positioned such that this one element is a real int."""
    cell = np.zeros(1, dtype=np.int32)
    size = (s1.size + 1) * (s2.size + 1)
    fake = cell.ctypes.data - 4 * (size - 1)
    rc = nw.lib().nw_cuda_fill_ex(s1.ctypes.data, s1.size, s2.ctypes.data, s2.size, C.c_void_p(fake), nw.NW_MODE_BOUNDARY, ngpus)
    if rc != 0:
        raise nw.NwCudaError(nw.lib().nw_cuda_last_error().decode(errors="replace"))
    return int(cell[0])


def cold_driver_runs(name, mode, runs=5):
    """The reference's unchanged driver around the plug-in call, one fresh process per run: the integer ms it prints."""
    if not os.path.exists(CUDA_E):
        return None
    a, b = pair_paths(name)
    if not (os.path.exists(a) and os.path.exists(b)):
        return None
    n1, n2 = SHAPES[name]
    if 4 * (n1 + 1) * (n2 + 1) / 2 ** 30 > mem_available_gb() * 0.6:
        return {"skipped": "host RAM: the driver allocates and touches the whole int32 table (src/common/driver.cpp:19-23)"}
    out = []
    for _ in range(runs):
        try:
            r = subprocess.run([CUDA_E, a, b], capture_output=True, text=True, timeout=600,
                               env=dict(os.environ, NW_CUDA_MODE=mode))
        except subprocess.TimeoutExpired:
            return {"error": "timeout"}
        m = re.match(r"\s*(\d+)\s*\nScore:\s*(-?\d+)", r.stdout)
        if r.returncode != 0 or not m or int(m.group(2)) != GOLDEN_SCORES[name]:
            return {"error": f"rc {r.returncode}: {r.stdout[-100:]} {r.stderr[-200:]}"}
        out.append(int(m.group(1)))
    return {"driver_printed_ms": out, "median_ms": float(np.median(out)), "runs": runs, "mode": mode,
            "what": "fresh process per run: fast-needleman-wunsch_b200/bin/cuda.e = the reference's unchanged driver.cpp + "
                    "helper.cpp around nw_cuda_fill; integer wall ms of the one call (context + pool set up before main)"}


def time_pair(nw, name, mode, steps, device, cpu_pairs):
    """A few timed fills of another BASELINE configuration (for the line's 'configs')."""
    s1, s2, data = load_pair(name)
    out = {"workload": workload_text(name, s1.size, s2.size, mode == nw.NW_MODE_FULL)}
    with nw.Plan(s1.size, s2.size, mode=mode, device=device) as p:
        p.upload(s1, s2)
        p.time(2)
        ms = p.time(steps)
        sc = p.score()
        info = p.strip_info()
    ok = (sc == GOLDEN_SCORES[name]) if data.startswith("reference") else None
    out.update({"ms_per_fill": ms, "gcups": s1.size * s2.size / ms / 1e6, "score": sc, "score_matches_golden": ok,
                "rows_per_lane": info["rows_per_lane"], "nstrips": info["nstrips"], "steps": steps})
    if mode == nw.NW_MODE_BOUNDARY:
        with nw.Plan(s1.size, s2.size, mode=nw.NW_MODE_SCORE, device=device) as p:
            p.upload(s1, s2)
            p.time(2)
            ms2 = p.time(steps)
            if p.score() != sc:
                raise SystemExit(f"bench: score-mode score differs on {name}")
        out["score_only_ms"] = ms2
        out["score_only_gcups"] = s1.size * s2.size / ms2 / 1e6
        t0 = time.perf_counter()
        for _ in range(steps):
            if plugin_boundary_call(nw, s1, s2) != sc:
                raise SystemExit(f"bench: plug-in score differs on {name}")
        out["e2e_ms"] = (time.perf_counter() - t0) * 1e3 / steps
    if cpu_pairs and name in cpu_pairs:
        out["cpu_same_pair"] = {k: v for k, v in cpu_pairs[name].items() if isinstance(v, dict)}
    return out


# score: oracle/nw_oracle.c (two-row restatement of serial.cpp) on one host core, 2 678 s.  (The round's first choice,
# 262 144 x 1 048 576 / seed 20261018 / score -524289, is critical-path-bound again from 4 GPUs on -- 44.9 / 25.1 / 23.7 ms at
# 1 / 2 / 4 -- because 4096 strips x 200 steps of start-up lag is a floor no column split removes; this one is four times larger.)
TALL = {"n1": 524288, "n2": 2097152, "seed": 20261019, "score": -1048577}


def tall_pair(nw, torch, dist, pipeline, world, rank, device, steps=3):
    """Secondary measurement at every N: a THROUGHPUT-bound single pair (8192 strips of 256 rows against 592 warp slots per
    GPU, so every scheduler always has a strip to work on), column strips over the N ranks exactly like the headline pair.
    This is the regime in which the mpi-vert decomposition (src/mpi/mpi-vert.cpp:17-105) pays: the headline pair is
    bound by its critical path on one GPU already."""
    n1, n2 = TALL["n1"], TALL["n2"]
    rng = np.random.default_rng(TALL["seed"])
    s1 = rng.integers(1, 5, size=n1, dtype=np.int8)
    s2 = rng.integers(1, 5, size=n2, dtype=np.int8)
    plan = nw.Plan(n1, n2, device=device, part=rank, nparts=world, rows_per_lane=8)
    try:
        if world > 1:
            pipeline.exchange_mailboxes(dist, plan, rank, world)
        plan.upload(s1, s2)
        plan.sync()

        def step():
            plan.run()
            plan.sync()
            if world > 1:
                dist.barrier()
            return plan.last_ms()
        step()
        ms = [step() for _ in range(steps)]
        t = torch.tensor(ms, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.mean().item())
        score = plan.score() if rank == world - 1 else None
        if world > 1:
            box = [score]
            dist.broadcast_object_list(box, src=world - 1)
            score = box[0]
        info = plan.strip_info()
    finally:
        plan.close()
    if TALL["score"] is not None and score != TALL["score"]:
        raise SystemExit(f"bench: tall pair score {score} != oracle {TALL['score']}")
    return {"workload": f"tall synthetic pair ({n1} x {n2} = {n1 * n2} cells, iid bases, seed {TALL['seed']}), boundary-only mode",
            "parallelism": f"column strips x{world}" if world > 1 else "single GPU", "nstrips": info["nstrips"],
            "ms_per_fill": ms, "gcups": n1 * n2 / ms / 1e6, "steps": steps, "score": score,
            "score_checked_against": "CPU oracle (two-row restatement of serial.cpp)" if TALL["score"] is not None else None,
            "what": "latency of one fill, max over ranks; throughput-bound (more strips than resident warps), so column "
                    "strips over N GPUs shorten it"}


def widened_numbers(nw, device):
    """SURVEY.md 8(f) rows measured beside the headline (one GPU): other scoring triples through the same kernels,
    Smith-Waterman, and the alignment output without a table."""
    out = []
    s1, s2, _ = load_pair("64gb")
    for sc in ((2, -1, -2), (5, -4, -3)):
        with nw.Plan(s1.size, s2.size, device=device, scoring=sc) as p:
            p.upload(s1, s2)
            p.time(1)
            ms = p.time(3)
            out.append({"workload": f"64gb pair, boundary-only mode, scoring match/mismatch/gap = {sc}", "ms_per_fill": ms,
                        "gcups": s1.size * s2.size / ms / 1e6, "score": p.score()})
    a, b, _ = load_pair("mid")
    sc = (2, -1, -2, 1)
    with nw.Plan(a.size, b.size, device=device, scoring=sc) as p:
        p.upload(a, b)
        p.time(1)
        ms = p.time(3)
        out.append({"workload": f"mid pair, Smith-Waterman (local alignment), scoring {sc[:3]}: best cell + position",
                    "ms_per_fill": ms, "gcups": a.size * b.size / ms / 1e6, "best": list(p.best())})
    for nm in ("2gb", "64gb"):
        x, y, _ = load_pair(nm)
        nw.align(x, y)
        t0 = time.perf_counter()
        a1, a2, score = nw.align(x, y)
        ms = (time.perf_counter() - t0) * 1e3
        ok = bool(np.array_equal(a1[a1 != 0], x) and np.array_equal(a2[a2 != 0], y) and
                  int(np.where((a1 == 0) | (a2 == 0), -1, (a1 == a2).astype(np.int64)).sum()) == score)
        out.append({"workload": f"{nm} pair, global alignment WITHOUT a table (nw_cuda_align: checkpoint rows + columns, tile "
                                "replay), host sequences in, gapped sequences out", "wall_ms": ms, "columns": int(a1.size),
                    "score": score, "score_matches_golden": score == GOLDEN_SCORES.get(nm),
                    "alignment_spells_both_sequences_and_scores": ok})
    return out


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    nw = importlib.import_module("fast-needleman-wunsch_b200")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local))
    device = local
    nw.init(device)                                   # raises if libnw_cuda.so or the GPU is missing: no fallback

    if args.workload == "batch":
        return batch_arm(args, nw, torch, dist, world, rank, device)

    name = args.workload.replace("-full", "")
    full = args.workload.endswith("-full")
    mode = nw.NW_MODE_FULL if full else nw.NW_MODE_BOUNDARY
    s1, s2, data = load_pair(name)
    n1, n2 = s1.size, s2.size
    cells = n1 * n2

    # every rank owns one column strip (part = rank); all parts must share the strip height chosen by rank 0
    pipeline = importlib.import_module("fast-needleman-wunsch_b200.pipeline")
    R = args.rows_per_lane
    if world > 1 and R == 0:
        def choose():
            with nw.Plan(n1, n2, mode=mode, device=device, part=0, nparts=world) as probe:
                return probe.strip_info()["rows_per_lane"]
        R = pipeline.agree_rows_per_lane(dist, rank, choose)
    plan = nw.Plan(n1, n2, mode=mode, device=device, part=rank, nparts=world, rows_per_lane=R)
    if world > 1:
        pipeline.exchange_mailboxes(dist, plan, rank, world)
    plan.upload(s1, s2)
    plan.sync()
    info = plan.strip_info()

    def barrier():
        if world > 1:
            dist.barrier()

    def cpu_barrier():
        """A barrier that leaves the GPUs alone (gloo all-reduce of a host tensor): the NCCL barrier is a kernel that spins on
        every waiting rank's GPU, and rank 0 is about to use those GPUs from its own process."""
        if world > 1:
            dist.all_reduce(torch.zeros(1))

    def one_step():
        """One fill, alone: launch, wait for it, (N > 1) wait for everybody.  Returns this rank's device ms."""
        plan.run()
        plan.sync()
        barrier()
        return plan.last_ms()

    sampler = ClockSampler(device)
    barrier()
    sampler.start()
    for _ in range(args.warmup):
        one_step()
    barrier()
    torch.cuda.synchronize()
    # ---- timed region: exactly K fills, each timed by CUDA events on the plan's own stream (inside the library) -------
    t_wall0 = time.perf_counter()
    step_ms = [one_step() for _ in range(args.steps)]
    torch.cuda.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    t = torch.tensor(step_ms + [t_wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)     # per step: the slowest rank (the last part finishes last)
    step_ms = [float(x) for x in t[:-1]]
    t_wall = float(t[-1]) / 1e3
    ms_step = float(np.mean(step_ms))
    gcups = cells / ms_step / 1e6

    score = plan.score() if rank == world - 1 else None
    if world > 1:
        box = [score]
        dist.broadcast_object_list(box, src=world - 1)
        score = box[0]
    expect = GOLDEN_SCORES.get(name) if data.startswith("reference") else None
    if expect is not None and score != expect:
        raise SystemExit(f"bench: score {score} != golden {expect} for {name}: refusing to report a number")

    # ---- back-to-back fills (no sync between them): throughput of overlapped consecutive fills ---------------------------
    barrier()
    plan.timer_start()
    for _ in range(args.steps):
        plan.run()
    pipe_ms = plan.timer_stop() / args.steps
    plan.sync()
    if world > 1:
        t = torch.tensor([pipe_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pipe_ms = float(t.item())
    barrier()

    # ---- score-only mode, kernel only (one GPU) -----------------------------------------------------------------------
    score_only = None
    if world == 1 and not full:
        with nw.Plan(n1, n2, mode=nw.NW_MODE_SCORE, device=device, rows_per_lane=args.rows_per_lane) as splan:
            splan.upload(s1, s2)
            splan.time(max(1, min(args.warmup, 3)))
            so_ms = splan.time(args.steps)
            if splan.score() != score:
                raise SystemExit(f"bench: score-mode score {splan.score()} != {score}")
            score_only = {"value": cells / so_ms / 1e6, "unit": "GCUPS", "ms_per_step": so_ms,
                          "mode": "NW_MODE_SCORE: top half filled forwards, bottom half backwards, concurrently; "
                                  "H[n2][n1] = max_j F[m][j] + B[m][j]; same number of cell updates, bit-exact score",
                          "launches_per_step": splan.launches_per_run()}

    # ---- score-only mode with one half per GPU (N >= 2; rank 0 drives devices 0 and 1 in-process) --------------------------
    if world >= 2 and not full:
        barrier()
        torch.cuda.synchronize()
        cpu_barrier()
        if rank == 0:
            nw.init(1)
            with nw.Plan(n1, n2, mode=nw.NW_MODE_SCORE, device=0, part=0, nparts=2) as splan:
                splan.upload(s1, s2)
                splan.time(max(1, min(args.warmup, 3)))
                so_ms = splan.time(args.steps)
                if splan.score() != score:
                    raise SystemExit(f"bench: two-GPU score-mode score {splan.score()} != {score}")
                score_only = {"value": cells / so_ms / 1e6, "unit": "GCUPS", "ms_per_step": so_ms, "gpus_used": 2,
                              "mode": "NW_MODE_SCORE, part 0 of 2: the forward half of the table on GPU 0, the reversed half "
                                      "on GPU 1, cut along a staircase (every strip sweeps only its side of it, so the strips' "
                                      "start-up lag overlaps their shorter sweeps); no traffic between the GPUs until the "
                                      "combine kernel reads the second half's boundary rows through peer access",
                              "launches_per_step": splan.launches_per_run()}
        cpu_barrier()

    # ---- end to end: HOST sequences in, fill, result out, every step -----------------------------------------------------
    table = None
    if world == 1:
        if full:
            table = np.empty((n2 + 1, n1 + 1), dtype=np.int32)        # pageable, like the driver's `new int[size]`
            table[::1024].fill(0)

            def e2e_step():
                nw.needlemanWunsch(s1, s2, table, mode=nw.NW_MODE_FULL)      # nw_cuda_fill_ex
                return int(table[-1, -1])
            e2e_path = "nw_cuda_fill_ex(host s1, host s2, host table (pageable), NW_MODE_FULL, 1) per step"
        else:
            def e2e_step():
                return plugin_boundary_call(nw, s1, s2)
            e2e_path = ("nw_cuda_fill_ex(host s1, host s2, host table, NW_MODE_BOUNDARY, 1) per step -- the call "
                        "csrc/cuda.cpp makes for the reference's driver; score-only mode (NW_MODE_SCORE) inside, plan cached "
                        "across calls (warm)")
    elif full:
        def e2e_step():
            plan.upload(s1, s2)                # H2D of this part's slice + all of s2, operand encoding
            plan.run()
            r = plan.score() if rank == world - 1 else plan.sync()    # D2H of the score on the last part
            barrier()
            return r
        e2e_path = "nw_plan_upload(host s1, s2) + nw_plan_run + nw_plan_score per step on every rank (one process per GPU)"
    else:
        # the plug-in call of the reference-side binding with NW_CUDA_GPUS = N, made by rank 0 (one process drives the
        # devices, like cuda.e does); the other ranks wait at the barrier
        for d in range(min(world, 2)):
            if rank == 0:
                nw.init(d)

        def e2e_step():
            r = plugin_boundary_call(nw, s1, s2, ngpus=world) if rank == 0 else None
            cpu_barrier()
            return r
        e2e_path = (f"nw_cuda_fill_ex(host s1, host s2, host table, NW_MODE_BOUNDARY, {world}) per step from rank 0 -- the call "
                    "csrc/cuda.cpp makes with NW_CUDA_GPUS=N: score-only mode with one half of the table on each of the first "
                    "two GPUs (cut along a staircase); plan cached across calls (warm)")
    barrier()
    for _ in range(min(args.warmup, 3)):
        e2e_step()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_rank = world - 1 if (world == 1 or full) else 0
    if rank == e2e_rank and expect is not None and r != expect:
        raise SystemExit(f"bench: end-to-end score {r} != golden {expect}")
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_gcups = cells * args.steps / e2e_s / 1e9
    h2d = int(plan.ncols + n2) * world if (world > 1 and full) else int(n1 + n2) * min(world, 2)
    d2h = 4 + (int(table.nbytes) if table is not None else 0)

    launches = plan.launches_per_run() * args.steps
    if world > 1:
        t = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        launches = int(t.item())

    tall = None
    if not args.no_configs and not full:
        try:
            tall = tall_pair(nw, torch, dist, pipeline, world, rank, device)
        except SystemExit:
            raise
        except Exception as e:          # (every rank fails alike: sizes and memory are the same)
            tall = {"workload": "tall synthetic pair", "error": str(e)[:200]}

    if rank == 0:
        peaks = measured_peaks()
        dpx_g, dpx_mhz = nw.dpx_peak(device)                      # measured integer/DPX pipe rate of THIS GPU
        ops_per_cell = 3.0
        achieved = gcups * ops_per_cell / 1e3                     # T lane-ops/s
        peak = dpx_g / 1e3 * world
        # boundary traffic: one tagged 8-byte word written and one read per column per strip (+ 4 B/cell in full mode)
        hbm_bytes = 16.0 * n1 * info["nstrips"] + (4.0 * cells if full else 0.0)
        roof = {"bound": "hbm" if full else "dpx-int32 pipe (no tensor cores: max-plus recurrence)",
                "kernel": "nw_full16_kernel (pass 2) + nw_strip16_kernel (pass 1)" if full else "nw_strip16l2_kernel",
                "achieved": achieved, "peak": peak, "unit": "T int32 lane-op/s", "frac": achieved / peak,
                "ops_per_cell": ops_per_cell, "peak_source": f"measured here: nw_cuda_dpx_peak = {dpx_g:.0f} G lane-op/s "
                f"per GPU at {dpx_mhz:.0f} MHz ({dpx_g * 1e3 / (148 * dpx_mhz):.1f} lanes/clk/SM)",
                "peak_gcups": peak * 1e3 / ops_per_cell,
                "note": "frac follows BASELINE.md section 4 (3 int32-pipe op per cell).  The kernel that ran packs two "
                        "cells per DPX instruction (s16x2: 1.5 op per cell), so the pipe's own limit is twice peak_gcups; "
                        "frac_s16x2 is the fraction of THAT limit.",
                "frac_s16x2": achieved / peak / 2.0,
                "hbm": {"achieved_gbs": hbm_bytes / (ms_step * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                        "peak_source": "MEASURED_PEAKS.json" if peaks.get("hbm_gbs") else "absent",
                        "algorithmic_bytes_per_step": hbm_bytes},
                "dependency_bound_steps": n1 + info["nstrips"] * 158, "traffic": ncu_traffic(name, full)}
        if full and peaks.get("hbm_gbs"):
            roof.update({"achieved": hbm_bytes / (ms_step * 1e-3) / 1e9, "peak": peaks["hbm_gbs"] * world, "unit": "GB/s",
                         "frac": hbm_bytes / (ms_step * 1e-3) / 1e9 / (peaks["hbm_gbs"] * world)})
        cpu = cpu_baseline(s1, s2, name) if (world == 1 and not args.no_cpu_baseline) else None
        cpu_pairs = (cpu or {}).get("pairs")
        configs, cold = [], None
        if world == 1 and not args.no_configs:
            for nm, md in (("2gb", nw.NW_MODE_BOUNDARY), ("2gb", nw.NW_MODE_FULL), ("mid", nw.NW_MODE_BOUNDARY),
                           ("big", nw.NW_MODE_BOUNDARY)):
                if nm == name and md == mode:
                    continue
                try:
                    configs.append(time_pair(nw, nm, md, 5, device, cpu_pairs))
                except Exception as e:      # never lose the headline line to a side measurement
                    configs.append({"workload": nm, "error": str(e)[:200]})
            try:
                configs.append(batch_numbers(nw, torch, 200000, 3, device)[0])
            except Exception as e:
                configs.append({"workload": "batch", "error": str(e)[:200]})
            try:
                configs.extend(widened_numbers(nw, device))
            except Exception as e:
                configs.append({"workload": "scoring / local / align", "error": str(e)[:200]})
            # fresh-process runs of the reference's driver around the plug-in (pairs whose host table is quick to
            # allocate and touch; the 64gb pair's 64 GB table takes the driver ~30 s per run: profiles/r02_driver_cold.log)
            cold = {"2gb_boundary": cold_driver_runs("2gb", "boundary"), "2gb_full": cold_driver_runs("2gb", "full", runs=3),
                    "mid_boundary": cold_driver_runs("mid", "boundary", runs=3)}
        line = {"metric": METRIC, "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "int32", "data": data,
                "config": {"workload": workload_text(name, n1, n2, full),
                           "parallelism": f"column strips x{world} (mpi-vert partition), NVLink mailbox handoff" if world > 1
                           else "single GPU", "rows_per_lane": info["rows_per_lane"], "strip_rows": info["strip_rows"],
                           "nstrips": info["nstrips"], "ctas": info["ctas"], "warps_per_cta": info["warps"],
                           "l2": "no flush: the boundary-row working set (%.0f MB per fill) exceeds the 126 MB L2; "
                                 "inputs are 0.25 MB" % (hbm_bytes / 2 / 1e6),
                           "score": score, "wall_ms_per_step": t_wall * 1e3 / args.steps,
                           "step_ms_min_max": [min(step_ms), max(step_ms)],
                           "timing": "every step = one fill timed alone by CUDA events on the plan's stream, then a stream "
                                     "sync (and a barrier for N > 1): ms_per_step is the latency of one fill, max over ranks",
                           "modes": "value = one forward fill that keeps every strip boundary row and the last row/column "
                                    "(NW_MODE_BOUNDARY plan; column strips for N > 1); score_only = the score-only algorithm "
                                    "(NW_MODE_SCORE; one half per GPU on two GPUs for N >= 2), kernel only; e2e = the plug-in "
                                    "call, which uses NW_MODE_SCORE inside"},
                "clocks": clocks,
                "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_s * 1e3 / args.steps, "path": e2e_path},
                "pipelined_throughput": {"value": cells / pipe_ms / 1e6, "unit": "GCUPS", "ms_per_fill": pipe_ms,
                                         "what": "K fills enqueued back to back without a sync in between (for N > 1 "
                                                 "consecutive fills overlap across the GPUs); NOT the latency of one fill"},
                "gpu_launches": launches, "roofline": roof}
        if tall is not None:
            line["throughput_bound_pair"] = tall
        if score_only is not None:
            line["score_only"] = score_only
        if cold is not None:
            line["e2e_cold"] = cold
        if configs:
            line["configs"] = configs
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def batch_numbers(nw, torch, npairs, steps, device, seed=20240607):
    """BASELINE.json configs[4] on one GPU: kernel-only and end-to-end GCUPS of `npairs` independent 1 kb pairs."""
    L = 1000
    rng = np.random.default_rng(seed)
    S1 = rng.integers(1, 5, size=(npairs, L), dtype=np.int8)
    S2 = rng.integers(1, 5, size=(npairs, L), dtype=np.int8)
    b = nw.Batch(npairs, L, L, device=device)
    p1, p2 = torch.from_numpy(S1).pin_memory(), torch.from_numpy(S2).pin_memory()
    b.upload(p1.numpy(), p2.numpy())
    b.time(1)
    ms = b.time(steps)
    out = torch.empty(npairs, dtype=torch.int32).pin_memory().numpy()
    sc = b.run_host(p1.numpy(), p2.numpy(), out)            # chunked: copies, kernels and score read-backs overlap
    t0 = time.perf_counter()
    for _ in range(steps):
        sc = b.run_host(p1.numpy(), p2.numpy(), out)
    e2e_s = (time.perf_counter() - t0) / steps
    sc = sc.copy()
    b.close()
    out = {"workload": f"batch of {npairs} independent pairs, 1000 x 1000 each (BASELINE.json configs[4] at 1/5 size)"
                       if npairs != 1000000 else "batch of 1M pairs, 1000 x 1000",
           "ms_per_step": ms, "gcups": npairs * L * L / ms / 1e6, "e2e_gcups": npairs * L * L / e2e_s / 1e9, "steps": steps}
    return out, (S1, S2, sc)


def batch_arm(args, nw, torch, dist, world, rank, device):
    """BASELINE.json configs[4]: independent 1 kb pairs, one pair-set per GPU, no data-path collective (weak scaling)."""
    npairs, L = args.batch_pairs, 1000
    sampler = ClockSampler(device)
    if world > 1:
        dist.barrier()
    sampler.start()
    (res, (S1, S2, sc)) = batch_numbers(nw, torch, npairs, args.steps, device, seed=20240607 + rank)
    clocks = sampler.stop()
    # the checker: a sample of the scores against the CPU restatement (oracle/ is test infrastructure; not timed)
    orc = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    orc.nw_oracle_batch_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    orc.nw_oracle_batch_scores.restype = None
    idx = np.concatenate([np.arange(64), np.arange(npairs - 64, npairs)])
    a, b2 = np.ascontiguousarray(S1[idx]), np.ascontiguousarray(S2[idx])
    want = np.empty(idx.size, dtype=np.int32)
    orc.nw_oracle_batch_scores(a.ctypes.data, b2.ctypes.data, idx.size, L, L, want.ctypes.data)
    if not np.array_equal(sc[idx], want):
        raise SystemExit("bench: batch scores differ from the oracle: refusing to report a number")
    ms, e2e_g = res["ms_per_step"], res["e2e_gcups"]
    if world > 1:
        t = torch.tensor([ms, 1.0 / e2e_g], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_g = float(t[0]), 1.0 / float(t[1])
    cells = npairs * L * L * world
    if rank == 0:
        gcups = cells / ms / 1e6
        dpx_g, dpx_mhz = nw.dpx_peak(device)
        achieved, peak = gcups * 3 / 1e3, dpx_g / 1e3 * world
        print(json.dumps({
            "metric": "GCUPS batch NW scores (1 kb pairs)", "value": gcups, "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic iid bases (numpy default_rng(20240607+rank))",
            "config": {"workload": f"batch of {npairs} pairs per GPU, 1000 x 1000 each", "l2":
                       "inputs %.0f MB per GPU > L2" % (2 * npairs * L / 1e6),
                       "scores_checked": f"{idx.size} pairs per rank against oracle/nw_oracle.c"},
            "clocks": clocks,
            "e2e": {"value": e2e_g * world, "unit": "GCUPS", "h2d_bytes_per_step": 2 * npairs * L,
                    "d2h_bytes_per_step": 4 * npairs},
            "gpu_launches": args.steps * world,
            "roofline": {"bound": "dpx-int32 pipe", "achieved": achieved, "peak": peak, "unit": "T int32 lane-op/s",
                         "frac": achieved / peak, "frac_s16x2": achieved / peak / 2.0, "ops_per_cell": 3.0,
                         "note": "frac counts 3 int32-pipe op per cell (BASELINE.md section 4); it exceeds 1.0 because the "
                                 "batch kernel packs two cells per DPX instruction (VIADDMNMX.S16x2, 1.5 op per cell); "
                                 "frac_s16x2 is the fraction of the packed-instruction limit",
                         "traffic": None}}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="64gb", choices=["64gb", "big", "mid", "2gb", "2gb-full", "smid", "smid-full", "batch"])
    ap.add_argument("--rows-per-lane", type=int, default=0)
    ap.add_argument("--batch-pairs", type=int, default=200000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the side measurements (other pairs, batch, cold driver runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
