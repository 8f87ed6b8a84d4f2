#!/usr/bin/env python3
"""Developer sweep: kernel-only GCUPS of the single-pair fill for several pairs / rows-per-lane, plus the DPX peak."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nw = importlib.import_module("fast-needleman-wunsch_b200")
BDNA = os.path.join(ROOT, "oracle", "_ref", "bdna")

def pair(name):
    if name.endswith("gb"):
        return (np.fromfile(f"{BDNA}/{name}-1.bdna", dtype=np.int8), np.fromfile(f"{BDNA}/{name}-2.bdna", dtype=np.int8))
    return (np.fromfile(f"{BDNA}/{name}1.bdna", dtype=np.int8), np.fromfile(f"{BDNA}/{name}2.bdna", dtype=np.int8))

def main():
    nw.init(0)
    print(json.dumps(nw.device_info(0)))
    g, mhz = nw.dpx_peak(0)
    print(f"dpx_peak: {g:.1f} G lane-ops/s at {mhz:.0f} MHz -> {g*1e9/(148*mhz*1e6):.1f} lanes/clk/SM", flush=True)
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["2gb", "mid", "64gb"]
    Rs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 4, 8]
    warps = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [8]
    for name in names:
        s1, s2 = pair(name)
        cells = s1.size * s2.size
        for R in Rs:
            for w in warps:
                with nw.Plan(s1.size, s2.size, rows_per_lane=R, warps_per_cta=w) as p:
                    p.upload(s1, s2)
                    p.time(2)
                    ms = p.time(5)
                    print(f"{name} R={R} warps={w} {p.strip_info()} score={p.score()} ms={ms:.3f} GCUPS={cells/ms/1e6:.1f}", flush=True)

if __name__ == "__main__":
    main()
