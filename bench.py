#!/usr/bin/env python3
"""bench.py -- GCUPS of the single-pair Needleman-Wunsch fill (BASELINE.json metric) on 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 64gb|big|mid|2gb|2gb-full|batch] [--impl reference]

A "step" is one complete fill of the pair's scoring table.  The default workload is the reference's 64gb-1/64gb-2
fixture pair (BASELINE.json configs[3]; 126 440 x 127 240 = 16.09 G cells) in boundary-only memory mode, which fits
one GPU, so the same job is timed at N = 1, 2, 4, 8 (strong scaling: column strips over the GPUs, NVLink handoff of
the boundary column; reference decomposition: src/mpi/mpi-vert.cpp:17-105).  For N > 1 the driver launches one
process per GPU with torchrun; torch.distributed is only plumbing (rendezvous, handle exchange, barrier, max-reduce).

JSON line (rank 0): value = kernel-only GCUPS with the sequences resident in HBM (CUDA events on the plan's own
stream, max over ranks); e2e = the same metric through the reference-facing C-ABI call with HOST buffers (pinned
sequences in, score out, copies inside the timed region); roofline = the strip kernel against the MEASURED integer/DPX
pipe rate of this GPU (3 lane-ops per cell); cpu_baseline = the reference's own serial.cpp (and its
OpenMP variants) timed on this box's host cores on a bounded sample.

--impl reference times the reference's CPU implementation (compiled unmodified into oracle/_ref/) with all host
threads, on a bounded prefix of the same pair.
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
METRIC = "GCUPS single-pair NW fill"
GOLDEN_SCORES = {"64gb": 73888, "big": 58529, "mid": 29249, "2gb": 12958, "smid": 5839}
SHAPES = {"64gb": (126440, 127240), "big": (100063, 99977), "mid": (49902, 49555), "2gb": (22541, 22116),
          "smid": (10030, 9976)}


def load_pair(name):
    """The reference's bdna fixture when it was staged next to the compiled reference; else seeded synthetic bases of
    the same lengths (iid uniform on 1..4, like the fixtures)."""
    if name.endswith("gb"):
        a, b = os.path.join(BDNA, f"{name}-1.bdna"), os.path.join(BDNA, f"{name}-2.bdna")
    else:
        a, b = os.path.join(BDNA, f"{name}1.bdna"), os.path.join(BDNA, f"{name}2.bdna")
    if os.path.exists(a) and os.path.exists(b):
        return np.fromfile(a, dtype=np.int8), np.fromfile(b, dtype=np.int8), f"reference bdna fixture {name} (1 byte per base)"
    n1, n2 = SHAPES[name]
    rng = np.random.default_rng(20240607)
    return (rng.integers(1, 5, size=n1, dtype=np.int8), rng.integers(1, 5, size=n2, dtype=np.int8),
            f"synthetic iid bases, lengths of the {name} fixture")


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
def run_ref_binary(exe, a, b, threads):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false")
    out = subprocess.run([os.path.join(REF, exe), a, b], capture_output=True, text=True, env=env, timeout=900)
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


def cpu_baseline(s1, s2, workload):
    """Reference serial.cpp on one core + its OpenMP variants on all cores, driver-printed ms, bounded sample."""
    cores = os.cpu_count() or 1
    if not have_reference_binaries():
        # restatement of the same loop (oracle/nw_oracle.c), one core
        lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        lib.nw_oracle_score.restype = C.c_int32
        lib.nw_oracle_score.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        n = 40000
        a, b = np.ascontiguousarray(s1[:n]), np.ascontiguousarray(s2[:n])
        t0 = time.perf_counter()
        lib.nw_oracle_score(a.ctypes.data, a.size, b.ctypes.data, b.size)
        dt = time.perf_counter() - t0
        return {"value": a.size * b.size / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port",
                "sample": f"two-row restatement on the {a.size}x{b.size} prefix of {workload}"}
    n = 24000
    a, b, m1, m2 = write_prefix_pair(s1, s2, n)
    cells = m1 * m2
    ms, _ = run_ref_binary("serial.e", a, b, 1)
    out = {"value": cells / max(ms, 1) / 1e6, "unit": "GCUPS", "cores": 1, "kind": "reference",
           "sample": f"reference serial.e (src/serial/serial.cpp, unmodified) on the {m1}x{m2} prefix of {workload}, "
                     f"driver-printed {ms} ms", "host_cores": cores}
    mt = {}
    for exe in ("idxarray-mod-mt.e", "sentinel-otf-blocked-mt.e"):
        if not os.path.exists(os.path.join(REF, exe)):
            continue
        best = None
        for th in sorted({min(4, cores), min(8, cores), cores}):
            ms_t, _ = run_ref_binary(exe, a, b, th)
            if ms_t and (best is None or ms_t < best[0]):
                best = (ms_t, th)
        if best:
            mt[exe[:-2]] = {"value": cells / max(best[0], 1) / 1e6, "unit": "GCUPS", "threads": best[1], "ms": best[0]}
    out["multithreaded"] = mt
    return out


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation, all host threads, bounded prefix of the same pair."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload.replace("-full", "")
    if name == "batch":
        name = "64gb"
    s1, s2, data = load_pair(name)
    cores = os.cpu_count() or 1
    n = 24000
    a, b, m1, m2 = write_prefix_pair(s1, s2, n)
    cells = m1 * m2
    if have_reference_binaries():
        kind = "reference"
        cands = [("serial.e", 1)]
        for exe in ("sentinel-otf-blocked-mt.e", "idxarray-mod-mt.e"):
            if os.path.exists(os.path.join(REF, exe)):
                cands.append((exe, cores))
        # untimed: pick the fastest variant on this host (the reference does not say which is its best)
        trial = []
        for exe, th in cands:
            ms, _ = run_ref_binary(exe, a, b, th)
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
    for _ in range(max(0, args.warmup - len(cands) if have_reference_binaries() else args.warmup)):
        step()
    secs = [step() for _ in range(args.steps)]
    total = sum(secs)
    value = cells * args.steps / total / 1e9
    sample = f"{impl} on the {m1}x{m2} prefix of {name}; time = the reference driver's own printed wall ms"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": data,
            "config": {"workload": f"{name} pair, bounded sample {m1}x{m2}", "cells_per_step": cells},
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
    mode = nw.NW_MODE_FULL if args.workload.endswith("-full") else nw.NW_MODE_BOUNDARY
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

    def one_step():
        plan.run()

    sampler = ClockSampler(device)
    barrier()
    sampler.start()
    for _ in range(args.warmup):
        one_step()
    plan.sync()
    barrier()
    torch.cuda.synchronize()
    # ---- timed region: exactly K fills, CUDA events on the plan's own stream (inside the library) -------------------
    t_wall0 = time.perf_counter()
    plan.timer_start()
    for _ in range(args.steps):
        one_step()
    ms_total = plan.timer_stop()          # CUDA events on the plan's stream; synchronises
    plan.sync()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    barrier()
    clocks = sampler.stop()
    if world > 1:
        # device time of a pipeline: every rank brackets its K fills with CUDA events (its first kernel starts right
        # after the barrier and spins on its halo until the left neighbour delivers); the job time is the max over ranks
        t = torch.tensor([ms_total, t_wall * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, t_wall = float(t[0].item()), float(t[1].item()) / 1e3
    ms_step = ms_total / args.steps
    gcups = cells / ms_step / 1e6

    score = plan.score() if rank == world - 1 else None
    if world > 1:
        box = [score]
        dist.broadcast_object_list(box, src=world - 1)
        score = box[0]
    expect = GOLDEN_SCORES.get(name) if data.startswith("reference") else None
    if expect is not None and score != expect:
        raise SystemExit(f"bench: score {score} != golden {expect} for {name}: refusing to report a number")

    # ---- end-to-end through the reference-facing call: HOST sequences in (pinned), fill, score out, every step --------
    pin1 = torch.from_numpy(s1.copy()).pin_memory()
    pin2 = torch.from_numpy(s2.copy()).pin_memory()
    h1, h2 = pin1.numpy(), pin2.numpy()
    if mode == nw.NW_MODE_FULL and world == 1:
        table = torch.empty((n2 + 1, n1 + 1), dtype=torch.int32).pin_memory().numpy()
    else:
        table = None

    # Boundary mode on one GPU delivers only the score, so the reference-facing call (nw_cuda_fill with
    # NW_CUDA_MODE=boundary, nw_cuda_score) meets in the middle: NW_MODE_SCORE, two half-length dependency chains.
    # It is timed on its own (score_only) and it is what the end-to-end number goes through.
    splan, score_only = None, None
    if world == 1 and mode == nw.NW_MODE_BOUNDARY:
        splan = nw.Plan(n1, n2, mode=nw.NW_MODE_SCORE, device=device, rows_per_lane=args.rows_per_lane)
        splan.upload(s1, s2)
        splan.time(max(1, min(args.warmup, 3)))
        so_ms = splan.time(args.steps)
        if splan.score() != score:
            raise SystemExit(f"bench: score-mode score {splan.score()} != {score}")
        score_only = {"value": cells / so_ms / 1e6, "unit": "GCUPS", "ms_per_step": so_ms,
                      "mode": "NW_MODE_SCORE: top half filled forwards, bottom half backwards, concurrently; "
                              "H[n2][n1] = max_j F[m][j] + B[m][j]; same number of cell updates, bit-exact score",
                      "launches_per_step": splan.launches_per_run()}
    eplan = splan if splan is not None else plan

    def e2e_step():
        eplan.upload(h1, h2)               # H2D of both sequences + operand encoding
        eplan.run()
        if table is not None:
            eplan.table_to_host(table)     # D2H of the whole table (what the reference driver's caller owns)
        return eplan.score() if rank == world - 1 else eplan.sync()   # D2H of the score (driver.cpp:35 reads it)

    barrier()
    for _ in range(min(args.warmup, 3)):
        e2e_step()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()        # no barrier between steps: the mailbox ack words keep a producer at most two fills ahead
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_gcups = cells * args.steps / e2e_s / 1e9
    h2d = int(plan.ncols + n2) * world if world > 1 else int(n1 + n2)
    d2h = 4 + (int(table.nbytes) if table is not None else 0)

    launches = plan.launches_per_run() * args.steps
    if world > 1:
        t = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        launches = int(t.item())

    line = None
    if rank == 0:
        peaks = measured_peaks()
        dpx_g, dpx_mhz = nw.dpx_peak(device)                      # measured integer/DPX pipe rate of THIS GPU
        ops_per_cell = 3.0
        achieved = gcups * ops_per_cell / 1e3                     # T lane-ops/s
        peak = dpx_g / 1e3 * world
        strip_rows = info["strip_rows"]
        # boundary traffic: one tagged 8-byte word written and one read per column per strip (+ 4 B/cell in full mode)
        hbm_bytes = 16.0 * n1 * info["nstrips"] + (4.0 * cells if mode == nw.NW_MODE_FULL else 0.0)
        roof = {"bound": "dpx-int32 pipe (no tensor cores: max-plus recurrence)" if mode != nw.NW_MODE_FULL else "hbm",
                "achieved": achieved, "peak": peak, "unit": "T int32 lane-op/s", "frac": achieved / peak,
                "ops_per_cell": ops_per_cell, "peak_source": f"measured here: nw_cuda_dpx_peak = {dpx_g:.0f} G lane-op/s "
                f"per GPU at {dpx_mhz:.0f} MHz ({dpx_g * 1e3 / (148 * dpx_mhz):.1f} lanes/clk/SM)",
                "peak_gcups": peak * 1e3 / ops_per_cell,
                "note": "frac follows BASELINE.md section 4 (3 int32-pipe op per cell).  The kernel that ran packs two "
                        "cells per DPX instruction (s16x2: 1.5 op per cell), so the pipe's own limit is twice peak_gcups; "
                        "frac_s16x2 is the fraction of THAT limit." if info.get("packed", True) else "",
                "frac_s16x2": achieved / peak / 2.0,
                "hbm": {"achieved_gbs": hbm_bytes / (ms_step * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                        "peak_source": "MEASURED_PEAKS.json" if peaks.get("hbm_gbs") else "absent",
                        "algorithmic_bytes_per_step": hbm_bytes},
                "dependency_bound_steps": n1 + info["nstrips"] * 32, "traffic": ncu_traffic(name, mode == nw.NW_MODE_FULL)}
        if mode == nw.NW_MODE_FULL and peaks.get("hbm_gbs"):
            roof.update({"achieved": hbm_bytes / (ms_step * 1e-3) / 1e9, "peak": peaks["hbm_gbs"] * world, "unit": "GB/s",
                         "frac": hbm_bytes / (ms_step * 1e-3) / 1e9 / (peaks["hbm_gbs"] * world)})
        cpu = cpu_baseline(s1, s2, name) if (world == 1 and not args.no_cpu_baseline) else None
        line = {"metric": METRIC, "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "int32", "data": data,
                "config": {"workload": f"{name} pair ({n1} x {n2} = {cells} cells), "
                           f"{'full-table' if mode == nw.NW_MODE_FULL else 'boundary-only'} mode",
                           "parallelism": f"column strips x{world} (mpi-vert partition), NVLink mailbox handoff" if world > 1
                           else "single GPU", "rows_per_lane": info["rows_per_lane"], "strip_rows": strip_rows,
                           "nstrips": info["nstrips"], "ctas": info["ctas"], "warps_per_cta": info["warps"],
                           "l2": "no flush: the boundary-row working set (%.0f MB per fill) exceeds the 126 MB L2; "
                                 "inputs are 0.25 MB" % (hbm_bytes / 2 / 1e6),
                           "score": score, "wall_ms_per_step": t_wall * 1e3 / args.steps,
                           "modes": "value = one forward fill that keeps every strip boundary row and the last row/column "
                                    "(NW_MODE_BOUNDARY plan, column strips over the GPUs); score_only and, on one GPU, e2e = "
                                    "the score-only path the reference-facing call takes in boundary mode (NW_MODE_SCORE)"},
                "clocks": clocks,
                "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_s * 1e3 / args.steps,
                        "path": ("NW_MODE_SCORE plan: " if splan is not None else "") +
                                "nw_plan_upload(host s1,s2) + nw_plan_run + nw_plan_score per step"},
                "gpu_launches": launches, "roofline": roof}
        if score_only is not None:
            line["score_only"] = score_only
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    plan.close()
    if splan is not None:
        splan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def batch_arm(args, nw, torch, dist, world, rank, device):
    """BASELINE.json configs[4]: independent 1 kb pairs, one pair-set per GPU, no data-path collective (weak scaling)."""
    npairs, L = args.batch_pairs, 1000
    rng = np.random.default_rng(20240607 + rank)
    S1 = rng.integers(1, 5, size=(npairs, L), dtype=np.int8)
    S2 = rng.integers(1, 5, size=(npairs, L), dtype=np.int8)
    b = nw.Batch(npairs, L, L, device=device)
    p1, p2 = torch.from_numpy(S1).pin_memory(), torch.from_numpy(S2).pin_memory()
    b.upload(p1.numpy(), p2.numpy())
    sampler = ClockSampler(device)
    if world > 1:
        dist.barrier()
    sampler.start()
    b.time(max(1, args.warmup))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = b.time(args.steps)
    clocks = sampler.stop()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b.upload(p1.numpy(), p2.numpy())
        b.run()
        b.scores()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
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
                       "inputs %.0f MB per GPU > L2" % (2 * npairs * L / 1e6)},
            "clocks": clocks,
            "e2e": {"value": cells * args.steps / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": 2 * npairs * L,
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
