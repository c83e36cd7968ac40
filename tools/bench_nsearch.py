"""Neighbour-search-only microbenchmark (BASELINE.json config 4): hash + counting sort by cell + cell
start offsets on uniform-random positions in the box x,z in [h-8, 8-h], y in [h, 16], h chosen for a
mean occupancy of ~4 particles per cell, seeds 1024/1025/1026.

    python tools/bench_nsearch.py [--sizes 1 4 16 64 100] [--repeats 20]
Prints one JSON line per size: particles/s and algorithmic GB/s at 128 B per particle.
"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sph_b200 as S

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", type=float, nargs="+", default=[1, 4, 16, 64, 100], help="millions of particles")
ap.add_argument("--repeats", type=int, default=20)
args = ap.parse_args()
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6650.0
for m in args.sizes:
    n = int(m * 1e6)
    res = []
    for seed in (1024, 1025, 1026):
        # volume (16-2h)^2 * (16-h) ~ 4096 m^3; occupancy 4 => cells = n/4 => h = (4096*4/n)^(1/3)
        h = float((4096.0 * 4.0 / n) ** (1.0 / 3.0))
        rng = np.random.default_rng(seed)
        pos = np.empty((n, 3), np.float32)
        pos[:, 0] = rng.uniform(h - 8, 8 - h, n); pos[:, 1] = rng.uniform(h, 16, n); pos[:, 2] = rng.uniform(h - 8, 8 - h, n)
        s = S.scaled_settings(h)
        sim = S.Sim(s, capacity=n)
        sim.upload(pos, np.zeros_like(pos))
        sim.neighbor_search(3)
        sim.sync()
        stream = torch.cuda.ExternalStream(sim.stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            sim.neighbor_search(args.repeats)
            e1.record(stream)
        sim.sync()
        ms = e0.elapsed_time(e1) / args.repeats
        st = sim.stats()
        res.append(ms)
        cells = int(st.grid_cells)
        sim.close()
        del pos
    ms = float(np.median(res))
    print(json.dumps({"workload": "neighbour-search-only", "particles": n, "h": h, "grid_cells": cells, "ms": ms,
                      "particles_per_s": n / (ms * 1e-3), "algorithmic_GBs": 128.0 * n / (ms * 1e-3) / 1e9,
                      "frac_of_hbm_peak": 128.0 * n / (ms * 1e-3) / 1e9 / peak, "ms_per_seed": res}), flush=True)
